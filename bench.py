#!/usr/bin/env python
"""bench.py — almost-banded QR+solve systems/sec (BASELINE.json metric) on N B200s, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs all|none|c2f,e2e,cpu,c3,c4,c1,c5]

Headline workload = BASELINE.json configs[1]: batched AlmostBandedMatrix qr + ldiv!, fp64, n=1024, l=u=2, r=2,
65,536 systems PER GPU (independent systems are sharded across ranks: weak scaling, no collective on the
data path; torch.distributed is used only for the barrier and the max-over-ranks time).  One "step" =
qr(A) into a preallocated factor object, then F \\ b -> x, over the whole device-resident batch.

JSON keys beyond the base contract: `roofline` (dominant kernel = the QR sweep, timed live with CUDA events),
`cpu_baseline` (the oracle port on the host cores, bounded sample), `e2e` (same metric through the host-buffer
C-ABI call ssm_ab_solve_host, H2D/D2H inside the timed region, with the box's concurrent pinned-copy ceiling beside it),
`clocks`, `gpu_launches`, `parity` (structured residual of every system + relative error against the oracle on a sample)
and `configs`: bounded legs for the other BASELINE.json configs, each with its own ms, roofline, cpu_baseline and parity —
  c2_fused  the fused `A \\ b` of config 2 (no factor output; SURVEY 8d "stricter line")
  c3        BandedPlusSemiseparable QR + solve, n=4096, l=m=3, r=p=2, 16,384 systems           (N = 1 only)
  c4        one system of n = 2^24, l=u=4, r=2, chunked n-axis path                             (N = 1 only)
  c1        README exp(z) system, n = 20 and n = 10^6, GPU next to the CPU oracle               (N = 1 only)
  c5        the n = 256..8192 sweep, 2^20 systems in total, STRONG scaling over the N GPUs (speed-up against rank 0 alone).
The oracle (oracle/) is executed only by the cpu_baseline / parity / --impl reference legs — never inside a timed GPU region.
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import socket
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

N_SYS, L_BW, U_BW, RANK_R = 1024, 2, 2, 2
NB_PER_GPU = 65536
SEED0 = 20261019                         # seed = 20261019 + config index
SEED = SEED0 + 1
METRIC = "almost-banded QR+solve systems/sec"
UNIT = "systems/s"
C5_GROUPS = [(256, 524288), (512, 262144), (1024, 131072), (2048, 65536), (4096, 32768), (8192, 32768)]
ALL_LEGS = ("c2f", "e2e", "cpu", "c3", "c4", "c1", "c5")


def ab_bytes(n, l, u, r):
    """inputs read once + outputs written once, per system: (qr, ldiv!, fused A\\b)."""
    return 8 * n * (3 * l + 2 * u + 3 * r + 3), 8 * n * (2 * l + u + 2 * r + 4), 8 * n * (l + u + 2 * r + 3)


def bps_bytes(n, l, m, r, p):
    return 8 * n * (3 * l + 2 * m + 5 * r + 3 * p + 3), 8 * n * (2 * l + m + 4 * r + 2 * p + 4)


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_of(kernel_key):
    """dram__bytes_read+write per launch of one kernel, from the committed ncu capture (profiles/traffic.json)."""
    tf = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(tf.read_text()).get(kernel_key)
    except Exception:
        return None


def workload_config(gpus: int) -> dict:
    """`config` of the JSON line — the same object for both arms (`--impl ours` / `--impl reference`)."""
    n, l, u, r = N_SYS, L_BW, U_BW, RANK_R
    return {"workload": f"BASELINE.json configs[1]: batched AlmostBandedMatrix qr + ldiv! fp64, n={n}, l=u={l}, r={r}, "
                        f"{NB_PER_GPU} systems per GPU", "n": n, "l": l, "u": u, "r": r, "systems_per_gpu": NB_PER_GPU,
            "parallelism": f"batch-sharded x{gpus} (independent systems, no collective)",
            "l2": "inputs larger than L2: 17.7 GB streamed per step vs 126 MB L2", "seed": SEED}


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML during the timed region (recipe's clocks line)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self._stop, self.max_mhz, self.err, self.power = [], set(), threading.Event(), None, None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v not in (0,):
                    names.setdefault(v, k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for v, k in names.items():
                    if bits & v and "All" not in k and "None" not in k:
                        self.reasons.add(k)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.nv and self.t.is_alive():
            self.t.join(timeout=1)
        out = {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples), "power_w_max": (max(self.power) if self.power else None)}
        if self.err:
            out["error"] = self.err
        return out


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# =================================================================================================================
# CPU legs (the oracle port: C restatement of the reference algorithm, OpenMP over systems; Julia is not in this image)
# =================================================================================================================
def oracle_all_cores():
    """Load the oracle and give its OpenMP team every host core, whatever OMP_NUM_THREADS the launcher exported
    (torch.distributed.run sets it to 1 for nproc > 1).  Returns (module, threads really used)."""
    os.environ.pop("OMP_NUM_THREADS", None)
    import oracle
    return oracle, oracle.threads(oracle.host_cores())


def relerr_rows(xg, xo):
    import numpy as np
    return float(max(np.linalg.norm(xg[s] - xo[s]) / np.linalg.norm(xo[s]) for s in range(xo.shape[0])))


def cpu_c2(sample: int, repeats: int):
    """qr + ldiv! of `sample` systems of the headline workload on all host cores, best of `repeats`."""
    oracle, threads = oracle_all_cores()
    n, l, u, r = N_SYS, L_BW, U_BW, RANK_R
    bands, U, V, b = oracle.ab_generate(n, l, u, r, sample, SEED)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        F, Uo, tau = oracle.ab_qr_batch(bands, U, V, l, u)
        oracle.ab_ldiv_batch(F, Uo, V, tau, b, l, l + u)
        best = min(best, time.perf_counter() - t0)
    return {"value": sample / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample} systems of the same workload (n={n}, l=u={l}, r={r}, fp64), qr+ldiv!, best of {repeats}; "
                      "C restatement of the reference algorithm (oracle/ab_oracle.c, gcc -O3, OpenMP over systems); "
                      "the Julia reference cannot run in this image"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; Julia is unavailable) on all host
    threads, rank 0 only; each step a bounded sample of the workload, EXACTLY --steps timed steps after --warmup steps."""
    if rank != 0:
        return
    oracle, threads = oracle_all_cores()
    n, l, u, r = N_SYS, L_BW, U_BW, RANK_R
    sample = 8192

    def step(bands, U, V, b):
        F, Uo, tau = oracle.ab_qr_batch(bands, U, V, l, u)
        oracle.ab_ldiv_batch(F, Uo, V, tau, b, l, l + u)

    ops = oracle.ab_generate(n, l, u, r, sample, SEED)
    t0 = time.perf_counter(); step(*ops); t1 = time.perf_counter() - t0            # untimed probe: sizes the sample
    budget = 150.0
    while sample > 256 and t1 * (args.steps + args.warmup) > budget:
        sample //= 2; t1 /= 2
    if sample != 8192:
        ops = oracle.ab_generate(n, l, u, r, sample, SEED)
    for _ in range(args.warmup):
        step(*ops)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(*ops)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    val = sample / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "systems_per_step": sample,
        "note": f"CPU reference arm: rank 0 only, {threads} OpenMP threads on {oracle.host_cores()} host cores; one step = a bounded "
                f"sample of the workload ({sample} of its {NB_PER_GPU} systems per GPU)",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} systems per step; oracle port (Julia unavailable in this image)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# =================================================================================================================
# GPU helpers
# =================================================================================================================
def timeit(fn, steps, stream, warm=3, warm_seconds=0.25):
    """mean device time of `steps` calls (CUDA events on the launching stream) after at least `warm` calls AND `warm_seconds` of work: a leg
    that follows a CPU phase (oracle baseline, parity sample) finds the GPU at idle clocks, and three short calls do not bring them back."""
    import torch
    t0 = time.perf_counter(); k = 0
    while k < warm or time.perf_counter() - t0 < warm_seconds:
        fn(); k += 1
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def timeit_streams(fns, streams, steps, warm=3):
    """fns[i] enqueues one step of piece i on streams[i]; all pieces run concurrently.  Device time from a common start event to
    the last piece's end (every stream waits for the start event, the main stream waits for every end event)."""
    import torch
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
    t0 = time.perf_counter()
    run(warm)
    while time.perf_counter() - t0 < 0.25:          # idle clocks after a CPU phase: see timeit
        run(1)
    return run(steps)


def timed_leg(fn, steps, stream, gpu_index):
    """timeit + the SM clock / throttle reasons sampled while it ran (each leg reports its own `clocks`)"""
    sampler = ClockSampler(gpu_index)
    sampler.start()
    t = timeit(fn, steps, stream)
    return t, sampler.stop()


def release():
    import torch
    import ssm_b200 as ssm
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ssm.pool_trim(torch.cuda.current_device())


# =================================================================================================================
# legs
# =================================================================================================================
def leg_c3(args, ctx, stream, peak, with_cpu):
    """BASELINE configs[2]: BandedPlusSemiseparableMatrix QR (semiseparableqr.jl), n=4096, l=m=3, r=p=2, 16,384 systems."""
    import numpy as np
    import ssm_b200 as ssm
    from ssm_b200 import bps
    n, l, m, r, p, nb, seed = 4096, 3, 3, 2, 2, 16384, SEED0 + 2
    A = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, seed, ctx=ctx)
    b = ssm.Vec(ctx, n, nb).generate(seed, array_id=5)
    x = ssm.Vec(ctx, n, nb)
    F = bps.qr(A)
    l0 = ctx.launches
    F.refactor(A)
    launches_per_qr = ctx.launches - l0
    t_qr, clk = timed_leg(lambda: F.refactor(A), args.steps, stream, args.gpu_index)
    t_sv = timeit(lambda: F.solve(b, out=x), args.steps, stream)
    rr = A.residual(x, b)
    qb, sb = bps_bytes(n, l, m, r, p)
    out = {"workload": f"BASELINE.json configs[2]: BandedPlusSemiseparableMatrix qr + solve fp64, n={n}, l=m={l}, r=p={r}, {nb} systems",
           "value": nb / (t_qr * 1e-3), "unit": "systems/s (qr)", "qr_ms": t_qr, "solve_ms": t_sv, "steps": args.steps,
           "qr_plus_solve_systems_per_s": nb / ((t_qr + t_sv) * 1e-3), "launches_per_qr": int(launches_per_qr), "clocks": clk,
           "roofline": {"bound": "hbm", "kernel": "bps_qr_kernel (K6) incl. its reduction pre-pass", "achieved": qb * nb / t_qr / 1e6, "peak": peak,
                        "unit": "GB/s", "frac": qb * nb / t_qr / 1e6 / peak, "traffic": traffic_of("bps_qr_dram_bytes_per_qr"),
                        "algorithmic_bytes_per_launch": qb * nb,
                        "solve": {"achieved": sb * nb / t_sv / 1e6, "frac": sb * nb / t_sv / 1e6 / peak, "algorithmic_bytes": sb * nb}},
           "parity": {"relres_max": float(rr.max()), "relres_median": float(np.median(rr)), "tolerance_relerr": 1e-12}}
    if with_cpu:
        oracle, threads = oracle_all_cores()
        ns = 32
        xg = x.get()
        errs, tq = [], 0.0
        for s0 in (0, nb - ns):
            B, U, V, W, S, bb = oracle.bps_generate(n, l, m, r, p, ns, seed, s0=s0)
            oB, oV, oW, oS, ot = oracle.bps_qr_batch(B, U, V, W, S, l, m)
            ox = oracle.bps_solve_batch(oB, U, oV, oW, oS, ot, bb, l, l + m)
            errs.append(relerr_rows(xg[s0:s0 + ns], ox))
        out["parity"]["relerr_sample"] = max(errs)
        out["parity"]["relerr_sample_systems"] = f"first {ns} + last {ns} systems against oracle/bps_oracle.c"
        del xg
        sample = 32 * threads
        B, U, V, W, S, bb = oracle.bps_generate(n, l, m, r, p, sample, seed)
        best_q, best_s = float("inf"), float("inf")
        for _ in range(3):
            t0 = time.perf_counter(); o = oracle.bps_qr_batch(B, U, V, W, S, l, m); t1 = time.perf_counter()
            oracle.bps_solve_batch(o[0], U, o[1], o[2], o[3], o[4], bb, l, l + m); t2 = time.perf_counter()
            best_q, best_s = min(best_q, t1 - t0), min(best_s, t2 - t1)
        out["cpu_baseline"] = {"value": sample / best_q, "unit": "systems/s (qr)", "qr_plus_solve_systems_per_s": sample / (best_q + best_s),
                               "cores": threads, "kind": "port", "sample": f"{sample} systems of the same workload, best of 3 (oracle/bps_oracle.c)"}
    del A, b, x, F
    release()
    return out


def leg_c4(args, ctx, stream, peak, with_cpu):
    """BASELINE configs[3]: one almost-banded system, n = 2^24, l=u=4, r=2 ("bc" boundary-row form), chunked n-axis path (K10).
    Roofline on the FUSED byte count 8n(l+u+2r+3): the path returns the solution, not the factors."""
    import numpy as np
    import torch
    from ssm_b200 import chunked
    n, l, u, r, seed = 1 << 24, 4, 4, 2, SEED0 + 3
    A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, seed, dist=1, ctx=ctx)
    b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), seed)
    x = torch.empty_like(b)
    l0 = ctx.launches
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A.solve(b, out=x, sync=True)          # first solve of these operands: stage 1 with lanes = columns (nothing to build)
    t_first = (time.perf_counter() - t0) * 1e3
    launches = ctx.launches - l0
    t0 = time.perf_counter()
    A.solve(b, out=x, sync=True)          # second solve: builds the chunk-interleaved copy the lane-per-chunk stage 1 reads, then uses it
    t_second = (time.perf_counter() - t0) * 1e3
    t, clk = timed_leg(lambda: A.solve(b, out=x, sync=False), args.steps, stream, args.gpu_index)
    torch.cuda.synchronize()
    rr = A.residual(x, b)
    fused = ab_bytes(n, l, u, r)[2]
    qrldiv = sum(ab_bytes(n, l, u, r)[:2])
    out = {"workload": f"BASELINE.json configs[3]: one AlmostBandedMatrix system A \\ b fp64, n={n}, l=u={l}, r={r}, chunked n-axis sweep, 1 GPU",
           "value": 1e3 / t, "unit": "systems/s", "ms": t, "columns_per_s": n / (t * 1e-3), "chunks": A.chunks, "steps": args.steps,
           "launches_per_solve": int(launches), "clocks": clk,
           "first_solve_ms_wall": t_first, "second_solve_ms_wall": t_second,
           "note": "ms / value: solves of resident operands from the third on (stage 1 with one lane per chunk over the interleaved copy the second solve built)",
           "roofline": {"bound": "hbm", "kernel": "abc_* (K10: chunk QR + separator tree + back-substitution)", "achieved": fused / t / 1e6,
                        "peak": peak, "unit": "GB/s", "frac": fused / t / 1e6 / peak, "traffic": traffic_of("abc_solve_dram_bytes_per_solve"),
                        "algorithmic_bytes_per_launch": fused, "byte_count": "fused A\\b, 8n(l+u+2r+3) (no factor output)",
                        "frac_on_qr_plus_ldiv_bytes": qrldiv / t / 1e6 / peak},
           "parity": {"relres": float(rr), "tolerance_relerr": 1e-12}}
    if with_cpu:
        oracle, _ = oracle_all_cores()
        bands, U, V, bb = oracle.ab_generate(n, l, u, r, 1, seed, dist=1)
        t0 = time.perf_counter()
        ox = oracle.ab_solve(bands, U, V, bb, l, u)[0]
        dt = time.perf_counter() - t0
        xg = x.cpu().numpy()
        out["parity"]["relerr"] = float(np.linalg.norm(xg - ox) / np.linalg.norm(ox))
        out["parity"]["relerr_against"] = "the oracle's sequential sweep on the whole system"
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "systems/s", "seconds": dt, "cores": 1, "kind": "port",
                               "sample": "the whole system once (the reference's sweep is strictly sequential: one core)"}
        del bands, U, V, bb, ox, xg
    del A, b, x
    release()
    return out


def leg_c1(args, ctx, stream, with_cpu):
    """BASELINE configs[0]: README exp(z) Taylor system (l=1, u=0, rank-1 boundary row), n = 20 and n = 10^6 (README.md:50,73-97:
    "just over a second" for n = 10^6 in Julia).  GPU: batched sweep (n = 20), chunked path (n = 10^6); CPU: the oracle."""
    import numpy as np
    import torch
    import ssm_b200 as ssm
    from ssm_b200 import chunked
    out = {"workload": "BASELINE.json configs[0]: README exp(z) AlmostBandedMatrix (l=1,u=0, rank-1 boundary row) qr + \\ fp64"}

    def readme(n):
        bands = np.zeros((n, 2)); bands[0, 0] = 1.0; bands[1:, 0] = np.arange(1, n); bands[: n - 1, 1] = -1.0
        U = np.zeros((1, n)); U[0, 0] = 1.0
        V = np.ones((n, 1)); b = np.zeros(n); b[0] = math.e
        return bands, U, V, b

    fact = np.array([1.0 / math.factorial(k) for k in range(20)])
    for n, key in ((20, "n20"), (1_000_000, "n1e6")):
        bands, U, V, b = readme(n)
        rec = {}
        if n == 20:
            A = ssm.AlmostBandedMatrix(bands, (U, V), (1, 0), ctx=ctx)
            F = ssm.qr(A)
            bv = ssm.Vec(ctx, n, 1, b[None]); xv = ssm.Vec(ctx, n, 1)
            rec["gpu_ms_qr_plus_ldiv"] = timeit(lambda: (F.refactor(A), F.solve(bv, out=xv)), args.steps, stream)
            xg = xv.get()[0]
            del A, F, bv, xv
        else:
            A = chunked.LargeAlmostBandedMatrix(bands, (U, V), (1, 0), ctx=ctx)
            bd = torch.from_numpy(b).cuda(); xd = torch.empty_like(bd)
            rec["gpu_ms_solve"] = timeit(lambda: A.solve(bd, out=xd, sync=False), args.steps, stream)
            rec["chunks"] = A.chunks
            t0 = time.perf_counter()
            xg = chunked.LargeAlmostBandedMatrix(bands, (U, V), (1, 0), ctx=ctx).solve(b)
            rec["gpu_ms_host_arrays_end_to_end"] = (time.perf_counter() - t0) * 1e3
            del A, bd, xd
        rec["err_vs_1_over_k_factorial_first20"] = float(np.abs(xg[:20] - fact).max())
        if with_cpu:
            oracle, _ = oracle_all_cores()
            best = float("inf")
            for _ in range(3 if n > 1000 else 50):
                t0 = time.perf_counter(); ox = oracle.ab_solve(bands, U, V, b, 1, 0); best = min(best, time.perf_counter() - t0)
            rec["cpu_ms"] = best * 1e3
            rec["cpu"] = "oracle port, 1 core (the sweep is sequential), best of 3"
            rec["relerr_vs_oracle"] = float(np.linalg.norm(xg - ox) / np.linalg.norm(ox))
        out[key] = rec
    release()
    return out


def leg_c5(args, job, dev, peak, with_cpu):
    """BASELINE configs[4]: almost-banded qr + ldiv! sweep, (l,u,r) = (2,2,2), n = 256..8192, 2^20 systems in total — STRONG scaling:
    the fixed job is cut evenly over the N ranks (ssm_shard_range per group) and a rank runs its six shards concurrently, one library
    context (= CUDA stream) per shard.  At N > 1 rank 0 first runs the whole job alone, so the speed-up is measured in this run."""
    import numpy as np
    import torch
    import ssm_b200 as ssm
    from ssm_b200 import sharder
    seed = SEED0 + 4
    steps = max(3, min(args.steps, 10))
    tot = sum(nb for _, nb in C5_GROUPS)
    tot_bytes = sum(sum(ab_bytes(n, 2, 2, 2)[:2]) * nb for n, nb in C5_GROUPS)

    def build(pieces):
        objs, streams = [], []
        for gi, s0, s1 in pieces:
            n, nb = C5_GROUPS[gi][0], s1 - s0
            st = torch.cuda.Stream(); streams.append(st)
            cx = ssm.Context(dev, stream=st.cuda_stream)
            cx.set_option("group_smem_kb", 96)          # several contexts share the SMs
            sd = seed + 1000 * gi + s0
            A = ssm.AlmostBandedMatrix.generate(n, 2, 2, 2, nb, sd, ctx=cx)
            b = ssm.Vec(cx, n, nb).generate(sd)
            objs.append((A, b, ssm.Vec(cx, n, nb), ssm.qr(A), cx, (gi, s0, s1, sd)))
        torch.cuda.synchronize()
        return objs, streams

    def run(objs, streams):
        order = sorted(range(len(objs)), key=lambda i: -objs[i][0].n)          # longest serial chains first
        fns = [(lambda A=A, b=b, x=x, F=F: (F.refactor(A), F.solve(b, out=x))) for A, b, x, F, _, _ in objs]
        return timeit_streams([fns[i] for i in order], [streams[i] for i in order], steps)

    def need_bytes(pieces):
        return sum((C5_GROUPS[gi][0] + 48) * 21 * 8 * (s1 - s0) for gi, s0, s1 in pieces)

    out = {"workload": "BASELINE.json configs[4]: batch-sharded almost-banded qr + ldiv! sweep fp64, l=u=2, r=2, n=256..8192, 2^20 systems in total",
           "scaling": "strong", "n_gpus": job.world, "systems": tot, "plan": "every (n, batch) group cut evenly over the ranks (ssm_shard_range); "
           "a rank's six shards run concurrently, one context/stream each", "steps": steps}
    whole = [(gi, 0, nb) for gi, (_, nb) in enumerate(C5_GROUPS)]
    ms_1gpu, mode_1gpu = None, None
    rr1 = None
    if job.rank == 0:
        # rank 0 must reach the barrier below whatever happens here (the other ranks are waiting in it)
        try:
            free, _ = torch.cuda.mem_get_info()
            objs = streams = None
            if need_bytes(whole) + (3 << 30) < free:
                try:
                    objs, streams = build(whole)
                    ms_1gpu, mode_1gpu = run(objs, streams), "six groups concurrently (one stream each)"
                    rr1 = max(float(A.residual(x, b).max()) for A, b, x, F, _, _ in objs)
                except Exception:                           # e.g. fragmentation: fall through to group after group
                    ms_1gpu = None
                objs = streams = None
                release()
            if ms_1gpu is None:                             # not enough HBM for all six at once: group after group
                ms_1gpu, mode_1gpu, rr1 = 0.0, "six groups back to back (HBM too small to hold them together)", 0.0
                for piece in whole:
                    objs, streams = build([piece])
                    ms_1gpu += run(objs, streams)
                    rr1 = max(rr1, float(objs[0][0].residual(objs[0][2], objs[0][1]).max()))
                    objs = streams = None
                    release()
        except Exception as e:  # noqa: BLE001
            ms_1gpu, mode_1gpu = None, f"failed: {type(e).__name__}: {e}"[:200]
            if job.world == 1:
                raise
        release()
    job.barrier()
    if job.world == 1:
        ms, rr = ms_1gpu, rr1
        relerr = None
    else:
        mine = [(gi,) + sharder.shard_range(nb, job.rank, job.world) for gi, (_, nb) in enumerate(C5_GROUPS)]
        mine = [p for p in mine if p[2] > p[1]]
        objs, streams = build(mine)
        job.barrier()
        t = run(objs, streams)
        ms = job.max_over_ranks(t)
        rr = job.max_over_ranks(max(float(A.residual(x, b).max()) for A, b, x, F, _, _ in objs))
        del objs, streams
        release()
    out.update({"value": tot / (ms * 1e-3), "unit": UNIT, "ms": ms, "ms_1gpu": ms_1gpu, "mode_1gpu": mode_1gpu,
                "speedup_vs_1gpu": (ms_1gpu / ms) if (ms_1gpu and job.rank == 0) else None,
                "roofline": {"bound": "hbm", "achieved_per_gpu": tot_bytes / ms / 1e6 / job.world, "peak": peak, "unit": "GB/s",
                             "frac": tot_bytes / ms / 1e6 / job.world / peak, "algorithmic_bytes": tot_bytes},
                "parity": {"relres_max": rr, "tolerance_relerr": 1e-12}})
    if job.rank == 0 and with_cpu:
        # parity sample + CPU baseline on a bounded sample of every group (first systems of each group, seeds of the 1-GPU run)
        oracle, threads = oracle_all_cores()
        errs, cpu_cols, cpu_s = [], 0, 0.0
        cx = ssm.Context(dev)
        for gi, (n, nb) in enumerate(C5_GROUPS):
            ns = max(threads, (1 << 19) // n)
            bands, U, V, b = oracle.ab_generate(n, 2, 2, 2, ns, seed + 1000 * gi)
            t0 = time.perf_counter()
            Fo, Uo, tau = oracle.ab_qr_batch(bands, U, V, 2, 2)
            ox = oracle.ab_ldiv_batch(Fo, Uo, V, tau, b, 2, 4)
            cpu_s += time.perf_counter() - t0; cpu_cols += n * ns
            A = ssm.AlmostBandedMatrix.generate(n, 2, 2, 2, ns, seed + 1000 * gi, ctx=cx)
            xg = ssm.qr(A).ldiv(b)
            errs.append(relerr_rows(xg, ox))
            del A
        cols = sum(n * nb for n, nb in C5_GROUPS)
        out["parity"]["relerr_sample"] = max(errs)
        out["parity"]["relerr_sample_systems"] = "first 2^19/n systems of every group against oracle/ab_oracle.c"
        out["cpu_baseline"] = {"value": tot / (cpu_s * cols / cpu_cols), "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"2^19 columns of every group ({cpu_cols} of the job's {cols} columns), qr+ldiv!, scaled to the whole job by columns"}
        cx.close()
        release()
    return out


def copy_ceiling(pins, xout, job, reps=3, chunk=8192, nstreams=3):
    """What the box gives for this step's transfers with no kernel in the way: every rank moves the same bytes as one e2e step in the same
    pattern — chunks of `chunk` systems round-robin over `nstreams` streams, the four input slices H2D then the x slice D2H per chunk (what
    ssm_ab_solve_host's pipeline does between its kernels).  Returns aggregate GB/s over all ranks (H2D, D2H), best of `reps`."""
    import torch
    nb = xout.shape[0]
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    dev_in = [[torch.empty((chunk,) + tuple(p.shape[1:]), dtype=p.dtype, device="cuda") for p in pins] for _ in range(nstreams)]
    dev_out = [torch.empty((chunk,) + tuple(xout.shape[1:]), dtype=xout.dtype, device="cuda") for _ in range(nstreams)]
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize(); job.barrier()
        t0 = time.perf_counter()
        for k, s0 in enumerate(range(0, nb, chunk)):
            cnt = min(chunk, nb - s0)
            sl = k % nstreams
            with torch.cuda.stream(streams[sl]):
                for p, d in zip(pins, dev_in[sl]):
                    d[:cnt].copy_(p[s0:s0 + cnt], non_blocking=True)
                xout[s0:s0 + cnt].copy_(dev_out[sl][:cnt], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        job.barrier()
        best = min(best, job.max_over_ranks(dt))
    h2d = sum(p.numel() * 8 for p in pins); d2h = xout.numel() * 8
    del dev_in, dev_out
    return job.world * h2d / best / 1e9, job.world * d2h / best / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nb", type=int, default=NB_PER_GPU, help="systems per GPU (default = BASELINE config)")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 direct-load, 2 bulk-async ring")
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--legs", default="all", help="comma list of " + ",".join(ALL_LEGS) + " | all | none")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-seconds", type=float, default=20.0, help="time cap of the e2e leg (it runs --steps steps unless that exceeds the cap)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", str(port), str(Path(__file__).resolve())] + sys.argv[1:])

    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    legs = set(ALL_LEGS) if args.legs == "all" else set(x for x in args.legs.split(",") if x and x != "none")
    if args.no_e2e:
        legs.discard("e2e")
    if args.no_cpu:
        legs.discard("cpu")
    with_cpu = "cpu" in legs

    import numpy as np
    import torch
    import ssm_b200 as ssm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    args.gpu_index = physical_gpu_index(local_rank)
    from ssm_b200 import sharder
    job = sharder.Job(backend="nccl", device=torch.device("cuda", local_rank))   # barrier + max-over-ranks only: no data-path collective
    barrier = job.barrier

    n, l, u, r, nb = N_SYS, L_BW, U_BW, RANK_R, args.nb
    stream = torch.cuda.current_stream()
    ctx = ssm.Context(local_rank, stream=stream.cuda_stream)
    if args.variant:
        ctx.set_option("ab_variant", args.variant)
    if args.stages:
        ctx.set_option("pipeline_stages", args.stages)

    A = ssm.AlmostBandedMatrix.generate(n, l, u, r, nb, SEED + 1000 * rank, ctx=ctx)
    b = ssm.Vec(ctx, n, nb).generate(SEED + 1000 * rank)
    x = ssm.Vec(ctx, n, nb)
    F = ssm.qr(A)

    def step():
        F.refactor(A)            # qr(A)      : K1
        F.solve(b, out=x)        # F \ b -> x : K2 (Q'b) + K3 (back-substitution)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(physical_gpu_index(local_rank))
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    for k in range(args.steps):
        ev[k][0].record(stream)
        F.refactor(A)
        ev[k][1].record(stream)
        F.solve(b, out=x)
        ev[k][2].record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    launches = ctx.launches - launches0
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    qr_ms = [e[0].elapsed_time(e[1]) for e in ev]
    ld_ms = [e[1].elapsed_time(e[2]) for e in ev]
    total_ms_max = job.max_over_ranks(total_ms)
    ms_per_step = total_ms_max / args.steps
    value = world * nb / (ms_per_step * 1e-3)

    # parity record (outside the timed region): structured residual of every system of this rank's batch, and the relative
    # error against the oracle on the first and last 64 systems of rank 0's batch (north_star states the bar on relerr)
    relres = A.residual(x, b)
    parity = {"relres_max": float(relres.max()), "relres_median": float(np.median(relres)), "tolerance_relerr": 1e-12}
    x_host = x.get() if (rank == 0 and (with_cpu or "e2e" in legs)) else None
    if rank == 0 and with_cpu:
        oracle, _ = oracle_all_cores()
        ns, errs = min(64, nb), []
        for s0 in sorted({0, nb - ns}):
            ob, oU, oV, orhs = oracle.ab_generate(n, l, u, r, ns, SEED, s0=s0)
            Fo, Uo, tau = oracle.ab_qr_batch(ob, oU, oV, l, u)
            errs.append(relerr_rows(x_host[s0:s0 + ns], oracle.ab_ldiv_batch(Fo, Uo, oV, tau, orhs, l, l + u)))
        parity["relerr_sample"] = max(errs)
        parity["relerr_sample_systems"] = f"first {ns} + last {ns} systems of rank 0 against oracle/ab_oracle.c (qr + ldiv!)"

    qr_bytes, ldiv_bytes, fused_bytes = ab_bytes(n, l, u, r)
    peak, peak_src = measured_peak_gbs()
    qr_avg_ms = statistics.mean(qr_ms)
    achieved = qr_bytes * nb / (qr_avg_ms * 1e-3) / 1e9
    step_gbs = (qr_bytes + ldiv_bytes) * nb / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "ab_qr_kernel (K1, Householder sweep with tracked fill)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_of("ab_qr_kernel_dram_bytes_per_launch"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": qr_bytes * nb, "avg_launch_ms": qr_avg_ms, "share_of_step": qr_avg_ms / (total_ms / args.steps),
                "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak, "frac_of_8TBs": step_gbs / 8000.0,
                               "algorithmic_bytes_per_step": (qr_bytes + ldiv_bytes) * nb, "ldiv_avg_ms": statistics.mean(ld_ms),
                               "ldiv_traffic": traffic_of("ab_ldiv_dram_bytes_per_ldiv")}}
    configs = {}

    # ---- c2_fused: the fused A \ b (no factor output) — SURVEY 8(d)'s stricter line -----------------------------------------
    if "c2f" in legs:
        t_f = job.max_over_ranks(timeit(lambda: A.solve(b, out=x), args.steps, stream))
        rrf = A.residual(x, b)
        configs["c2_fused"] = {"workload": "configs[1] shape, fused A \\ b (ssm_ab_solve: factor + solve, factors not kept)", "value": world * nb / (t_f * 1e-3),
                               "unit": UNIT, "ms": t_f, "steps": args.steps,
                               "roofline": {"bound": "hbm", "achieved": fused_bytes * nb / t_f / 1e6, "peak": peak, "unit": "GB/s",
                                            "frac": fused_bytes * nb / t_f / 1e6 / peak, "algorithmic_bytes_per_launch": fused_bytes * nb,
                                            "traffic": traffic_of("ab_fused_dram_bytes_per_solve"),
                                            "note": "ceiling ~0.38: R must round-trip HBM because the back-substitution runs the other way"},
                               "parity": {"relres_max": float(rrf.max())}}

    # ---- e2e: the same metric through the host-buffer C-ABI call (H2D + pack + qr + solve + D2H per step) ----
    e2e = None
    if "e2e" in legs:
        nb_e = nb
        try:
            import psutil
            need = nb_e * (l + u + 1 + 2 * r + 2) * n * 8
            if psutil.virtual_memory().available < 2.5 * need * max(1, world):
                nb_e = max(2048, nb // 8)
        except Exception:
            pass
        hb, hU, hV = A._get() if nb_e == nb else ssm.AlmostBandedMatrix.generate(n, l, u, r, nb_e, SEED + 1000 * rank, ctx=ctx)._get()
        hrhs = b.get() if nb_e == nb else ssm.Vec(ctx, n, nb_e).generate(SEED + 1000 * rank).get()
        pins = [torch.from_numpy(a).pin_memory() for a in (hb, hU, hV, hrhs)]
        del hb, hU, hV, hrhs
        xout = torch.empty((nb_e, n), dtype=torch.float64).pin_memory()
        ectx = ssm.Context(local_rank)
        if args.variant:
            ectx.set_option("ab_variant", args.variant)
        t0 = time.perf_counter()
        ssm.solve_host(n, l, u, r, *pins, x=xout, ctx=ectx)       # warm-up (allocates the chunk pipeline)
        ssm.solve_host(n, l, u, r, *pins, x=xout, ctx=ectx)
        t_probe = job.max_over_ranks((time.perf_counter() - t0) / 2)
        e_steps = int(max(3, min(args.steps, args.e2e_seconds / max(t_probe, 1e-3))))
        torch.cuda.synchronize()
        barrier()
        e_l0 = ectx.launches
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ssm.solve_host(n, l, u, r, *pins, x=xout, ctx=ectx)   # synchronises before returning
        dt = time.perf_counter() - t0
        barrier()
        dt = job.max_over_ranks(dt) / e_steps
        h2d = int(nb_e * (l + u + 1 + 2 * r + 1) * n * 8); d2h = int(nb_e * n * 8)
        e2e = {"value": world * nb_e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3, "steps": e_steps, "systems_per_step_per_gpu": nb_e,
               "api": "ssm_ab_solve_host (pinned host buffers in the reference layout -> x in host memory)",
               "gpu_launches": int(ectx.launches - e_l0),
               "h2d_GBps_aggregate": world * h2d / dt / 1e9, "d2h_GBps_aggregate": world * d2h / dt / 1e9,
               "matches_device_path": (bool(np.array_equal(xout.numpy(), x_host)) if (x_host is not None and nb_e == nb) else None)}
        ceil_h2d, ceil_d2h = copy_ceiling(pins, xout, job)
        e2e["copy_ceiling"] = {"h2d_GBps_aggregate": ceil_h2d, "d2h_GBps_aggregate": ceil_d2h,
                               "how": "every rank moves one step's bytes in the pipeline's own pattern (8192-system chunks round-robin over 3 streams, inputs H2D then x D2H), no kernels; best of 3, max over ranks",
                               "e2e_frac_of_ceiling": (world * h2d / dt / 1e9) / ceil_h2d}
        # the single-process sharder north_star describes (one process, one host thread + context per GPU; the host buffers are the gather)
        barrier()
        if rank == 0:
            devs = list(range(min(world, torch.cuda.device_count())))
            sharder.solve_host_sharded(devs, n, l, u, r, *pins, x=xout)      # warm-up: contexts + chunk pipelines of every device
            reps = max(2, min(5, e_steps))
            t0 = time.perf_counter()
            for _ in range(reps):
                sharder.solve_host_sharded(devs, n, l, u, r, *pins, x=xout)
            dts = (time.perf_counter() - t0) / reps
            e2e["sharded_single_process"] = {"api": "ssm_ab_solve_host_sharded", "devices": devs, "systems": nb_e, "scaling": "strong (one host batch split over the devices)",
                                             "value": nb_e / dts, "unit": UNIT, "ms": dts * 1e3, "h2d_GBps": h2d / dts / 1e9,
                                             "matches_device_path": (bool(np.array_equal(xout.numpy(), x_host)) if (x_host is not None and nb_e == nb) else None)}
        barrier()
        ectx.close()
        del pins, xout
    del x_host

    cpu = cpu_c2(sample=16384, repeats=6) if (rank == 0 and world == 1 and with_cpu) else None

    # ---- the other BASELINE.json configs (bounded legs) ------------------------------------------------------------------
    del A, b, x, F
    release()
    def guarded(name, fn, collective=False):
        """a leg that fails must not take the headline line with it: its entry becomes {"error": ...}.  Legs every rank takes part in
        (barriers inside) are not guarded against rank-local failures — a rank that dropped out would hang the others."""
        try:
            configs[name] = fn()
        except Exception as e:  # noqa: BLE001
            if collective and world > 1:
                raise
            configs[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            try:
                release()
            except Exception:
                pass

    if world == 1:
        if "c3" in legs:
            guarded("c3", lambda: leg_c3(args, ctx, stream, peak, with_cpu))
        if "c4" in legs:
            guarded("c4", lambda: leg_c4(args, ctx, stream, peak, with_cpu))
        if "c1" in legs:
            guarded("c1", lambda: leg_c1(args, ctx, stream, with_cpu))
    if "c5" in legs:
        guarded("c5", lambda: leg_c5(args, job, local_rank, peak, with_cpu), collective=True)

    if rank == 0:
        config = workload_config(world)
        if nb != NB_PER_GPU:
            config["systems_per_gpu"] = nb
            config["workload"] += f" (run with --nb {nb})"
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "kernel_variant": ctx.get_option("ab_variant"),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches), "parity": parity,
            "configs": configs,
        }
        print(json.dumps(out))
    job.close()


if __name__ == "__main__":
    main()
