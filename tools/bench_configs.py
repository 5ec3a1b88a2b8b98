#!/usr/bin/env python
"""Secondary measurements (not the driver's bench contract): the other BASELINE.json configs on ONE GPU.

  python tools/bench_configs.py c3            BPS QR, n=4096, l=m=3, r=p=2, 16,384 systems  (configs[2])
  python tools/bench_configs.py c4            one large system n=2^24, l=u=4, r=2, chunked n-axis path (configs[3])
  python tools/bench_configs.py c5            almost-banded sweep n=256..8192, 1M systems total (configs[4]), this GPU's share
  torchrun --nproc-per-node N tools/bench_configs.py c5dist   the same sweep on N GPUs, pieces assigned by sharder.plan_groups
  python tools/bench_configs.py c2 [--variant V --stages S]   per-kernel timing of config 2 (tuning aid)

Prints one JSON line per measurement; timing = CUDA events on the library's stream, 3 warm-ups, mean of `--steps`.
"""
from __future__ import annotations

import argparse
import os
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    sys.path.insert(0, str(p))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ssm_b200 as ssm  # noqa: E402
from ssm_b200 import bps  # noqa: E402

PEAK = 6544.3
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timeit(fn, steps, warm=3):
    stream = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def c3(args, ctx):
    n, l, m, r, p, nb, seed = 4096, 3, 3, 2, 2, args.nb or 16384, 20261019 + 2
    A = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, seed, ctx=ctx)
    b = ssm.Vec(ctx, n, nb).generate(seed, array_id=5)
    x = ssm.Vec(ctx, n, nb)
    F = bps.qr(A)
    t_qr = timeit(lambda: F.refactor(A), args.steps)
    t_sv = timeit(lambda: F.solve(b, out=x), args.steps)
    rr = A.residual(x, b)
    qr_bytes = 8 * n * (3 * l + 2 * m + 5 * r + 3 * p + 3) * nb
    sv_bytes = 8 * n * (2 * l + m + 4 * r + 2 * p + 4) * nb
    print(json.dumps({"config": "c3 BPS qr n=4096 l=m=3 r=p=2", "systems": nb, "qr_ms": t_qr, "solve_ms": t_sv,
                      "qr_systems_per_s": nb / (t_qr * 1e-3), "qr_GBps_algorithmic": qr_bytes / t_qr / 1e6, "qr_frac_of_measured_peak": qr_bytes / t_qr / 1e6 / PEAK,
                      "solve_GBps_algorithmic": sv_bytes / t_sv / 1e6, "solve_frac_of_measured_peak": sv_bytes / t_sv / 1e6 / PEAK,
                      "relres_max": float(rr.max()), "relres_median": float(np.median(rr))}))


def c2(args, ctx):
    n, l, u, r, nb, seed = args.n or 1024, 2, 2, 2, args.nb or 65536, 20261019 + 1
    A = ssm.AlmostBandedMatrix.generate(n, l, u, r, nb, seed, ctx=ctx)
    b = ssm.Vec(ctx, n, nb).generate(seed)
    x = ssm.Vec(ctx, n, nb)
    F = ssm.qr(A)
    t_qr = timeit(lambda: F.refactor(A), args.steps)
    t_sv = timeit(lambda: F.solve(b, out=x), args.steps)
    t_fused = timeit(lambda: A.solve(b, out=x), args.steps)
    rr = A.residual(x, b)
    qr_bytes = 8 * n * (3 * l + 2 * u + 3 * r + 3) * nb
    sv_bytes = 8 * n * (2 * l + u + 2 * r + 4) * nb
    fused_bytes = 8 * n * (l + u + 2 * r + 3) * nb
    print(json.dumps({"config": f"almost-banded n={n} l=u=2 r=2", "systems": nb, "variant": args.variant, "stages": args.stages,
                      "qr_ms": t_qr, "ldiv_ms": t_sv, "fused_solve_ms": t_fused, "qr_plus_ldiv_systems_per_s": nb / ((t_qr + t_sv) * 1e-3),
                      "qr_frac": qr_bytes / t_qr / 1e6 / PEAK, "ldiv_frac": sv_bytes / t_sv / 1e6 / PEAK,
                      "qr_plus_ldiv_frac": (qr_bytes + sv_bytes) / (t_qr + t_sv) / 1e6 / PEAK,
                      "fused_strict_frac": fused_bytes / t_fused / 1e6 / PEAK, "fused_systems_per_s": nb / (t_fused * 1e-3),
                      "relres_max": float(rr.max())}))


def c5(args, ctx):
    """configs[4]: n = 256 .. 8192, 2^20 systems in total; this GPU's share = total / world (default world=1)."""
    groups = [(256, 524288), (512, 262144), (1024, 131072), (2048, 65536), (4096, 32768), (8192, 32768)]
    tot_sys, tot_ms = 0, 0.0
    for n, nbt in groups:
        nb = nbt // args.world
        args.n, args.nb = n, nb
        A = ssm.AlmostBandedMatrix.generate(n, 2, 2, 2, nb, 20261019 + 4, ctx=ctx)
        b = ssm.Vec(ctx, n, nb).generate(20261019 + 4)
        x = ssm.Vec(ctx, n, nb)
        F = ssm.qr(A)
        t = timeit(lambda: (F.refactor(A), F.solve(b, out=x)), args.steps)
        rr = A.residual(x, b)
        byts = 8 * n * 33 * nb
        print(json.dumps({"config": "c5", "n": n, "systems": nb, "ms": t, "systems_per_s": nb / (t * 1e-3), "frac_of_measured_peak": byts / t / 1e6 / PEAK,
                          "relres_max": float(rr.max())}))
        tot_sys += nb; tot_ms += t
        del A, b, x, F
    print(json.dumps({"config": "c5 total", "systems": tot_sys, "ms": tot_ms, "systems_per_s": tot_sys / (tot_ms * 1e-3)}))


def timeit_streams(fns, streams, steps, warm=3):
    """fns[i] enqueues one step of piece i on streams[i]; all pieces run concurrently.  Device time from a common start event to
    the last piece's end (CUDA events; every stream waits for the start event, the main stream waits for every end event)."""
    main = torch.cuda.current_stream()

    def run(k):
        start = torch.cuda.Event(enable_timing=True); stop = torch.cuda.Event(enable_timing=True)
        start.record(main)
        for st in streams:
            st.wait_event(start)
        for _ in range(k):
            for fn in fns:
                fn()
        for st in streams:
            e = torch.cuda.Event(); e.record(st); main.wait_event(e)
        stop.record(main)
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / k
    run(warm)
    return run(steps)


def c5c(args, ctx):
    """configs[4], this GPU's share, with the groups running CONCURRENTLY (one library context = one CUDA stream per group): small
    shares of the large-n groups are latency bound on their own (a few warps per SM) and fill the GPU only together."""
    groups = [(256, 524288), (512, 262144), (1024, 131072), (2048, 65536), (4096, 32768), (8192, 32768)]
    dev = torch.cuda.current_device()
    objs, streams = [], []
    for n, nbt in groups:
        nb = nbt // args.world
        st = torch.cuda.Stream()
        cx = ssm.Context(dev, stream=st.cuda_stream)
        cx.set_option("group_smem_kb", 96)          # several contexts share the SMs (scan 24..200 KB: 96 is best, 5.85 ms at 1/8 share)
        for kv in args.opt:
            k, v = kv.split("="); cx.set_option(k, int(v))
        A = ssm.AlmostBandedMatrix.generate(n, 2, 2, 2, nb, 20261019 + 4, ctx=cx)
        b = ssm.Vec(cx, n, nb).generate(20261019 + 4)
        x = ssm.Vec(cx, n, nb)
        objs.append((A, b, x, ssm.qr(A), cx)); streams.append(st)
    torch.cuda.synchronize()
    fns = [(lambda A=A, b=b, x=x, F=F: (F.refactor(A), F.solve(b, out=x))) for A, b, x, F, _ in objs]
    order = sorted(range(len(fns)), key=lambda i: -groups[i][0])          # longest serial chains first
    if os.environ.get("C5C_ORDER") == "asc":
        order = order[::-1]
    t = timeit_streams([fns[i] for i in order], streams, args.steps)
    rrs = [float(A.residual(x, b).max()) for A, b, x, F, _ in objs]
    rr = max(rrs)
    if rr > 1e-12:
        print("per-group relres", rrs, file=sys.stderr)
    tot = sum(nbt // args.world for _, nbt in groups)
    byts = sum(8 * n * 33 * (nbt // args.world) for n, nbt in groups)
    print(json.dumps({"config": "c5 concurrent groups", "share": f"1/{args.world}", "systems": tot, "ms": t, "systems_per_s": tot / (t * 1e-3),
                      "frac_of_measured_peak": byts / t / 1e6 / PEAK, "relres_max": rr}))


def c4(args, ctx):
    """configs[3]: one large almost-banded system, n = 2^24, l = u = 4, r = 2, chunked n-axis path (K10)."""
    from ssm_b200 import chunked
    n, l, u, r, seed = args.n or (1 << 24), 4, 4, 2, 20261019 + 3
    A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, seed, dist=1, ctx=ctx)
    b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), seed)
    x = torch.empty_like(b)
    for P in ([args.chunks] if args.chunks else [0, 2048, 4096, 16384, 32768]):
        t = timeit(lambda: A.solve(b, out=x, nchunks=P, sync=False), args.steps)
        torch.cuda.synchronize()
        rr = A.residual(x, b)
        qrldiv = 8 * n * (5 * l + 3 * u + 5 * r + 7)         # qr + ldiv! algorithmic bytes
        fused = 8 * n * (l + u + 2 * r + 3)                  # fused A\\b (no factor output)
        print(json.dumps({"config": f"c4 single system n={n} l=u=4 r=2 (bc)", "chunks": A.chunks, "ms": t, "systems_per_s": 1e3 / t,
                          "columns_per_s": n / (t * 1e-3), "GBps_qr_plus_ldiv_algorithmic": qrldiv / t / 1e6,
                          "frac_of_measured_peak_qr_plus_ldiv": qrldiv / t / 1e6 / PEAK, "frac_of_measured_peak_fused_strict": fused / t / 1e6 / PEAK,
                          "relres": rr}))


def c2m(args, ctx):
    """multi-RHS ldiv!(F, B::Matrix) on the config-2 shape: nrhs columns per system in one pass over the factors per group."""
    import ctypes as C
    from ssm_b200 import _lib
    n, l, u, r, nb, seed = 1024, 2, 2, 2, args.nb or 65536, 20261019 + 1
    A = ssm.AlmostBandedMatrix.generate(n, l, u, r, nb, seed, ctx=ctx)
    F = ssm.qr(A)
    b = ssm.Vec(ctx, n, nb).generate(seed)
    t1 = timeit(lambda: F.ldiv(b), args.steps)
    L = _lib.lib()
    for nrhs in (4, 8):
        h = C.c_void_p()
        _lib.check(L.ssm_mat_create(ctx._h, n, nrhs, nb, C.byref(h)), "ssm_mat_create")
        t = timeit(lambda: _lib.check(L.ssm_abqr_ldiv_mat(F._h, h), "ldiv_mat"), args.steps)
        byts = 8 * n * (2 * l + u + 2 * r + 1 + 3 * nrhs) * nb        # factors once + (read b, write x) per column; Q'B round trip not counted
        print(json.dumps({"config": f"ldiv!(F, B) n={n} (2,2,2) nrhs={nrhs}", "systems": nb, "ms": t, "ms_single_vector_ldiv": t1,
                          "speedup_vs_columnwise": nrhs * t1 / t, "frac_of_measured_peak": byts / t / 1e6 / PEAK}))
        L.ssm_mat_destroy(h)


def c3rq(args, ctx):
    """one QR-iteration step of a symmetric BPS batch: qr(A) then rq_mul(F) (R*Q), n=4096, l=m=3, r=p=2 (host-generated symmetric operands)."""
    n, l, r, nb = args.n or 4096, 3, 2, args.nb or 4096
    rng = np.random.default_rng(7)
    B = rng.uniform(-1, 1, (nb, n, 2 * l + 1))
    for t in range(1, l + 1):                       # symmetric band: A[j+t, j] (slot l+t of column j) = A[j, j+t] (slot l-t of column j+t)
        B[:, t:, l - t] = B[:, :n - t, l + t]
        B[:, :t, l - t] = 0.0; B[:, n - t:, l + t] = 0.0
    B[:, :, l] = 2 * l + 2 + rng.uniform(0, 1, (nb, n))
    U = rng.uniform(-1, 1, (nb, r, n)) / np.sqrt(r * n); V = rng.uniform(-1, 1, (nb, r, n)) / np.sqrt(r * n)
    A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (V.copy(), U.copy()), (l, l), ctx=ctx)
    F = bps.qr(A)
    t_qr = timeit(lambda: F.refactor(A), args.steps)
    import ctypes as C
    from ssm_b200 import _lib
    RQ = F.rq_mul()
    t_rq = timeit(lambda: _lib.check(_lib.lib().ssm_bpsqr_rq_mul(F._h, C.byref(RQ._h)), "rq_mul"), args.steps)     # into the existing batch
    gB, gU, gV, _, _ = RQ.get()
    # similarity transform: trace(R*Q) == trace(A)
    tr_err = float(np.abs(gB[:, :, l].sum(1) - B[:, :, l].sum(1)).max() / np.abs(B[:, :, l].sum(1)).max())
    rd = 8 * n * ((2 * l + l + 1) + l + r + 1 + 4 * r + 2 * r) * nb      # factor records read once (R rows, tails, k-bar, tau, [alpha beta], S, A'U, U)
    wr = 8 * n * ((2 * l + 1) + 2 * r) * nb
    print(json.dumps({"config": f"symmetric BPS QR-iteration step n={n} l={l} r={r}", "systems": nb, "qr_ms": t_qr, "rq_mul_ms": t_rq,
                      "rq_systems_per_s": nb / (t_rq * 1e-3), "rq_GBps_algorithmic": (rd + wr) / t_rq / 1e6, "trace_relerr": tr_err}))


def c5dist(args, ctx):
    """configs[4] on N GPUs, one process per GPU (torchrun).  --plan even (default): every group is cut into N contiguous shards
    (ssm_shard_range) and a rank runs its six shards CONCURRENTLY, one library context (= CUDA stream) per shard, so that the small
    shares of the large-n groups — latency bound on their own — fill the GPU together.  --plan lpt: whole groups as units
    (sharder.plan_groups), run back to back.  Time = max over ranks of the device time for the rank's whole list."""
    from ssm_b200 import sharder
    dev = torch.cuda.current_device()
    job = sharder.Job(backend="nccl", device=torch.device("cuda", dev))
    groups = [(256, 524288), (512, 262144), (1024, 131072), (2048, 65536), (4096, 32768), (8192, 32768)]
    if args.plan == "lpt":
        mine = sharder.plan_groups(groups, job.world)[job.rank]
    else:
        mine = [(gi,) + sharder.shard_range(nb, job.rank, job.world) for gi, (n, nb) in enumerate(groups)]
        mine = [p for p in mine if p[2] > p[1]]
    objs, streams = [], []
    for gi, s0, s1 in mine:
        n, nb = groups[gi][0], s1 - s0
        if args.plan == "lpt":
            cx = ctx
        else:
            st = torch.cuda.Stream(); streams.append(st)
            cx = ssm.Context(dev, stream=st.cuda_stream)
            cx.set_option("group_smem_kb", 96)      # several contexts share the SMs
        A = ssm.AlmostBandedMatrix.generate(n, 2, 2, 2, nb, 20261019 + 4 + 1000 * gi + s0, ctx=cx)
        b = ssm.Vec(cx, n, nb).generate(20261019 + 4 + 1000 * gi + s0)
        objs.append((A, b, ssm.Vec(cx, n, nb), ssm.qr(A), cx))
    torch.cuda.synchronize()

    def sweep():
        for A, b, x, F, _ in objs:
            F.refactor(A); F.solve(b, out=x)
    job.barrier()
    if not objs:
        t = 0.0
    elif args.plan == "lpt":
        t = timeit(sweep, args.steps)
    else:
        order = sorted(range(len(objs)), key=lambda i: -objs[i][0].n)          # longest serial chains first
        fns = [(lambda A=A, b=b, x=x, F=F: (F.refactor(A), F.solve(b, out=x))) for A, b, x, F, _ in objs]
        t = timeit_streams([fns[i] for i in order], streams, args.steps)
    tmax = job.max_over_ranks(t)
    rr = max([float(A.residual(x, b).max()) for A, b, x, F, _ in objs] or [0.0])
    rrmax = job.max_over_ranks(rr)
    per_rank = job.gather({"rank": job.rank, "ms": t, "pieces": [(groups[gi][0], s1 - s0) for gi, s0, s1 in mine]})
    if job.rank == 0:
        tot = sum(nb for _, nb in groups)
        print(json.dumps({"config": "c5 sharded sweep n=256..8192, 2^20 systems, qr+ldiv!", "plan": args.plan, "n_gpus": job.world, "ms": tmax,
                          "systems_per_s": tot / (tmax * 1e-3), "relres_max": rrmax, "per_rank": per_rank}))
    job.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["c2", "c2m", "c3", "c3rq", "c4", "c5", "c5c", "c5dist"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="extra context option key=value (tuning aid)")
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--plan", choices=["even", "lpt"], default="even")
    args = ap.parse_args()
    import os
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    ctx = ssm.Context(dev, stream=torch.cuda.current_stream().cuda_stream)
    if args.variant:
        ctx.set_option("ab_variant", args.variant)
    if args.stages:
        ctx.set_option("pipeline_stages", args.stages)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    {"c2": c2, "c2m": c2m, "c3": c3, "c3rq": c3rq, "c4": c4, "c5": c5, "c5c": c5c, "c5dist": c5dist}[args.which](args, ctx)
