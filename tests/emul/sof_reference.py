"""numpy restatement of the chunked n-axis solve (kernel family K10, csrc/abc_*.cu) — TEST INFRASTRUCTURE ONLY.

One large AlmostBandedMatrix system A x = b (AlmostBandedMatrix.jl:84-90 semantics) is cut into P row chunks.
Chunk c owns rows [r0, r1) and the "local" columns g = r0 + u + j', j' = 0..m-1 (m = r1 - r0).  With lp = l + u:

  separator  S_c   = columns [r0 - l, r0 + u)        (lp unknowns; phantom ones outside [0, n) have zero coefficients)
  tail       T_c   = sum_{g >= r1 + u} V[:, g] x_g   (r unknowns), so row i's fill is U[i,:] . (sum_{local j' > i'} V x + T_c)
  tail rows  T_{c-1} - T_c - sum_{local j'} V[:, g(j')] x = 0       (r extra equations per chunk)

Seen in local columns the chunk is a banded matrix with lower bandwidth lp and upper bandwidth 0 plus rank-r fill.  Its
m - lp interior columns are eliminated by Householder reflectors spanning the lp+1 active band rows and the r tail rows;
every row carries [window | generator (coefficient of T_c and of far local columns) | C_L (coefficients of S_c) |
coefficient of T_{c-1} | rhs].  What is left, (lp + r) rows in (S_c, T_{c-1}, S_{c+1}, T_c), forms a block-bidiagonal
reduced system  F_c z_c + G_c z_{c+1} = h_c,  z_c = [x_{S_c}; T_{c-1}],  which is reduced pairwise by QR (cyclic
reduction), solved at the root, expanded top-down, and finally every chunk back-substitutes its interior.
Every transformation is orthogonal or a triangular solve with an R factor of a column subset of the (augmented) matrix,
so the method is backward stable like the reference's sweep; it returns the SOLUTION, not the reference's factors.
"""
from __future__ import annotations

import numpy as np


def band_entry(bands, l, u, i, j):
    """A[i, j] of the band part; bands is (n, l+u+1) with bands[j, u + i - j] (reference column-major band storage)."""
    n = bands.shape[0]
    if i < 0 or j < 0 or i >= n or j >= n or j - i > u or i - j > l:
        return 0.0
    return bands[j, u + i - j]


def householder(x):
    """LinearAlgebra.reflector! convention: returns v (v[0] = 1), tau, new diagonal."""
    nrm = np.sqrt(np.dot(x, x))
    if nrm == 0.0:
        return np.zeros_like(x), 0.0, x[0]
    nu = np.copysign(nrm, x[0])
    xi = x[0] + nu
    v = x / xi
    v[0] = 1.0
    return v, xi / nu, -nu


def chunk_bounds(n, P):
    return [(c * n) // P for c in range(P + 1)]


def regular_layout(n, lp, P, same=4):
    """The CUDA path's chunk layout (abc_layout in csrc/abc_kernels.cu): every chunk length is == lp (mod lp+1), so each chunk
    eliminates a multiple of lp+1 columns; the system is padded with identity rows to n' = bounds[-1].  `same` consecutive chunks
    share a length: the 4 chunks one stage-1 warp sweeps, or the 32 chunks of a stage-3 tile (many chunks / tiled records).
    Returns (P, bounds)."""
    pw = lp + 1
    P = max(P, 1)
    while P > 1 and (n // P + 1) // pw < 2:
        P -= 1
    q = max((n // P + 1) // pw, 2)
    m0 = pw * q - 1
    rem = n - P * m0
    jx = (rem + pw - 1) // pw if rem > 0 else 0
    jx = min(P, -(-jx // same) * same)      # chunk lengths change only at multiples of `same`
    return P, [c * m0 + pw * min(c, jx) for c in range(P + 1)]


def pad_system(bands, U, V, b, l, u, npad):
    """append identity rows/columns so that the system has npad >= n rows (x_pad = 0)"""
    n = bands.shape[0]
    if npad == n:
        return bands, U, V, b
    e = npad - n
    bp = np.zeros((npad, l + u + 1)); bp[:n] = bands
    # entries of real columns that referred to rows >= n were zero (outside the matrix); pad columns get a unit diagonal
    for j in range(max(0, n - l), n):                 # stored corner entries below the original matrix are not part of it
        for d in range(l + u + 1):
            if j + d - u >= n:
                bp[j, d] = 0.0
    bp[n:, u] = 1.0
    Up = np.zeros((U.shape[0], npad)); Up[:, :n] = U
    Vp = np.zeros((npad, V.shape[1])); Vp[:n] = V
    return bp, Up, Vp, np.concatenate([b, np.zeros(e)])


def chunk_reduce(bands, U, V, b, l, u, r0, r1):
    """Stage 1 for one chunk.  Returns (Rrows, left) where Rrows[k] = record of the R row of local column k and
    left = the (lp + r) x (lp + r + lp + r + 1) leftover block [win(S_{c+1}) | gen | C_L | Tprev | rhs]."""
    n = bands.shape[0]
    r = U.shape[0]
    lp = l + u
    m = r1 - r0
    assert m >= lp
    nc = (lp + 1) + r + lp + r + 1
    W0, G0, C0, T0, B0 = 0, lp + 1, lp + 1 + r, lp + 1 + r + lp, nc - 1

    def vcol(jp):                     # V[:, g(jp)] (0 outside the matrix or the chunk)
        g = r0 + u + jp
        return V[g, :] if (0 <= jp < m and g < n) else np.zeros(r)

    # active rows: list of dicts {win: {jp: val}, gen, cl, tp, rhs}
    def band_row(ip):
        i = r0 + ip
        row = np.zeros(nc)
        gen = U[:, i].copy()
        return {"ip": ip, "gen": gen, "cl": np.array([band_entry(bands, l, u, i, r0 - l + t) for t in range(lp)]),
                "tp": np.zeros(r), "rhs": b[i], "win": {}}

    def win_entry(row, jp):
        """true current value of a band row at local column jp the first time it is needed"""
        i = r0 + row["ip"]
        g = r0 + u + jp
        if jp <= row["ip"]:
            return band_entry(bands, l, u, i, g)
        return float(np.dot(row["gen"], vcol(jp)))

    rows = [band_row(ip) for ip in range(lp)]
    for row in rows:
        for jp in range(0, lp + 1):
            row["win"][jp] = win_entry(row, jp)
    trows = []
    for t in range(r):
        e = np.zeros(r); e[t] = 1.0
        tr = {"ip": None, "gen": -e, "cl": np.zeros(lp), "tp": e.copy(), "rhs": 0.0, "win": {jp: -vcol(jp)[t] for jp in range(0, lp + 1)}}
        trows.append(tr)

    Rrows = []
    for k in range(m - lp):
        new = band_row(k + lp)
        for jp in range(k, k + lp + 1):
            new["win"][jp] = win_entry(new, jp)
        rows.append(new)
        act = rows + trows
        x = np.array([a["win"][k] for a in act])
        v, tau, diag = householder(x)
        # apply to every carried column
        def apply(get, put):
            col = np.array([get(a) for a in act])
            s = tau * np.dot(v, col)
            col = col - v * s
            for a, val in zip(act, col):
                put(a, val)
        for jp in range(k + 1, k + lp + 1):
            apply(lambda a: a["win"][jp], lambda a, val: a["win"].__setitem__(jp, val))
        for q in range(r):
            apply(lambda a: a["gen"][q], lambda a, val: a["gen"].__setitem__(q, val))
            apply(lambda a: a["tp"][q], lambda a, val: a["tp"].__setitem__(q, val))
        for t in range(lp):
            apply(lambda a: a["cl"][t], lambda a, val: a["cl"].__setitem__(t, val))
        apply(lambda a: a["rhs"], lambda a, val: a.__setitem__("rhs", val))
        piv = rows.pop(0)
        rec = np.zeros(nc)
        rec[W0] = diag
        for t in range(1, lp + 1):
            rec[W0 + t] = piv["win"][k + t]
        rec[G0:G0 + r] = piv["gen"]; rec[C0:C0 + lp] = piv["cl"]; rec[T0:T0 + r] = piv["tp"]; rec[B0] = piv["rhs"]
        Rrows.append(rec)
        # slide: column k + lp + 1 enters for the surviving rows (fill region of every one of them)
        jn = k + lp + 1
        for a in rows + trows:
            a["win"][jn] = float(np.dot(a["gen"], vcol(jn)))
    left = np.zeros((lp + r, nc))
    for a_i, a in enumerate(rows + trows):
        for t in range(lp):
            left[a_i, W0 + t] = a["win"][m - lp + t]
        left[a_i, G0:G0 + r] = a["gen"]; left[a_i, C0:C0 + lp] = a["cl"]; left[a_i, T0:T0 + r] = a["tp"]; left[a_i, B0] = a["rhs"]
    return Rrows, left


def qr_rows(M, ncols):
    """Householder-triangularise the first ncols columns of M (in place copy); returns the transformed matrix."""
    M = M.copy()
    for j in range(ncols):
        v, tau, diag = householder(M[j:, j].copy())
        if tau != 0.0:
            s = tau * (v @ M[j:, :])
            M[j:, :] -= np.outer(v, s)
        M[j, j] = diag
        M[j + 1:, j] = 0.0
    return M


def solve_chunked(bands, U, V, b, l, u, P, regular=False):
    """x = A \\ b through P chunks.  bands (n, l+u+1), U (r, n), V (n, r), b (n).  regular=True uses the CUDA path's padded
    chunk layout, otherwise near-equal chunks of the unpadded system."""
    n_in = bands.shape[0]
    if regular:
        P, bnd = regular_layout(n_in, l + u, P)
        bands, U, V, b = pad_system(bands, U, V, b, l, u, bnd[-1])
        return _solve_chunked(bands, U, V, b, l, u, P, bnd)[:n_in]
    return _solve_chunked(bands, U, V, b, l, u, P, chunk_bounds(n_in, P))


def _solve_chunked(bands, U, V, b, l, u, P, bnd):
    n = bands.shape[0]
    r = U.shape[0]
    lp = l + u
    w = lp + r
    R, F, G, h = [], [], [], []
    for c in range(P):
        Rr, left = chunk_reduce(bands, U, V, b, l, u, bnd[c], bnd[c + 1])
        R.append(Rr)
        W0, G0, C0, T0, B0 = 0, lp + 1, lp + 1 + r, lp + 1 + r + lp, left.shape[1] - 1
        F.append(np.hstack([left[:, C0:C0 + lp], left[:, T0:T0 + r]]))
        G.append(np.hstack([left[:, W0:W0 + lp], left[:, G0:G0 + r]]))
        h.append(left[:, B0].copy())
    # cyclic reduction of F_c z_c + G_c z_{c+1} = h_c
    levels = []
    cur = [(F[c], G[c], h[c], c, c + 1) for c in range(P)]       # (F, G, h, left index, right index)
    while len(cur) > 1:
        nxt, saved = [], []
        for i in range(0, len(cur) - 1, 2):
            Fa, Ga, ha, iL, iM = cur[i]
            Fb, Gb, hb, _, iR = cur[i + 1]
            M = np.zeros((2 * w, 3 * w + 1))
            M[:w, 0:w] = Ga; M[w:, 0:w] = Fb                      # eliminated unknown z_M first
            M[:w, w:2 * w] = Fa; M[w:, 2 * w:3 * w] = Gb
            M[:w, 3 * w] = ha; M[w:, 3 * w] = hb
            M = qr_rows(M, w)
            saved.append((iL, iM, iR, M[:w].copy()))
            nxt.append((M[w:, w:2 * w].copy(), M[w:, 2 * w:3 * w].copy(), M[w:, 3 * w].copy(), iL, iR))
        if len(cur) % 2:
            nxt.append(cur[-1])
        levels.append(saved)
        cur = nxt
    Fr, Gr, hr, _, _ = cur[0]
    # root: unknowns = real entries of z_0 (S_0 columns >= 0, T_{-1}) and of z_P (S_P columns < n; T_{P-1} = 0)
    z = [None] * (P + 1)
    real0 = [t for t in range(lp) if bnd[0] - l + t >= 0] + list(range(lp, w))
    realP = [t for t in range(lp) if bnd[P] - l + t < n]
    Mroot = np.hstack([Fr[:, real0], Gr[:, realP]])
    sol = np.linalg.solve(Mroot, hr)
    z0 = np.zeros(w); z0[real0] = sol[:len(real0)]
    zP = np.zeros(w); zP[realP] = sol[len(real0):]
    z[0], z[P] = z0, zP
    for saved in reversed(levels):
        for iL, iM, iR, top in saved:
            rhs = top[:, 3 * w] - top[:, w:2 * w] @ z[iL] - top[:, 2 * w:3 * w] @ z[iR]
            z[iM] = np.linalg.solve(np.triu(top[:, 0:w]), rhs)
    # stage 3: interior back-substitution
    x = np.zeros(n)
    for c in range(P + 1):
        for t in range(lp):
            g = bnd[c] - l + t
            if 0 <= g < n:
                x[g] = z[c][t]
    for c in range(P):
        r0, r1 = bnd[c], bnd[c + 1]
        m = r1 - r0
        G0, C0, T0 = lp + 1, lp + 1 + r, lp + 1 + r + lp
        xl = np.zeros(m)
        xl[m - lp:] = z[c + 1][:lp]
        s = z[c + 1][lp:].copy()
        for k in range(m - lp - 1, -1, -1):
            jn = k + lp + 1
            if jn <= m - 1:
                g = r0 + u + jn
                if g < n:
                    s = s + V[g, :] * xl[jn]
            rec = R[c][k]
            acc = rec[-1] - rec[G0:G0 + r] @ s - rec[C0:C0 + lp] @ z[c][:lp] - rec[T0:T0 + r] @ z[c][lp:]
            for t in range(1, lp + 1):
                acc -= rec[t] * xl[k + t]
            xl[k] = acc / rec[0]
        for jp in range(m - lp):
            x[r0 + u + jp] = xl[jp]
    return x


def dense(bands, U, V, l, u):
    n = bands.shape[0]
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if j > i + u:
                A[i, j] = U[:, i] @ V[j, :]
            elif i - j <= l:
                A[i, j] = bands[j, u + i - j]
    return A
