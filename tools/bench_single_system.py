"""ONE system: the chunked path (ssm_abc_*) against the sequential sweep kernels run with a batch of one — the measurement behind the
Julia glue's CHUNKED_MIN_N (profiles/r1_measurements.md).  Run on the GPU box from the repo root: python tools/bench_single_system.py"""
import sys, json, time
sys.path.insert(0, "semiseparablematrices.jl_b200"); sys.path.insert(0, ".")
import torch, ssm_b200 as ssm
from ssm_b200 import chunked
ctx = ssm.default_context(0)
def timeit(f, k=20):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(ctx.stream) if hasattr(ctx, "stream") else None
    t0 = time.perf_counter()
    for _ in range(k): f()
    ctx.sync(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3
for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    for (l, u, r) in ((4, 4, 2), (1, 0, 1)):
        A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, 7, dist=1, ctx=ctx)
        b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), 7)
        x = torch.empty_like(b)
        tc = timeit(lambda: A.solve(b, out=x, sync=False))
        B = ssm.AlmostBandedMatrix.generate(n, l, u, r, 1, 7, dist=1, ctx=ctx)
        bb = ssm.Vec(ctx, n, 1).generate(7); xx = ssm.Vec(ctx, n, 1)
        ts = timeit(lambda: B.solve(bb, out=xx))
        print(json.dumps({"n": n, "lur": [l, u, r], "chunked_ms": round(tc, 4), "chunks": A.chunks, "sequential_sweep_ms": round(ts, 4)}))
