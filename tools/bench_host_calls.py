"""What one `A \\ b` costs through the single-system path with HOST arrays (create + set + solve + destroy, as the Julia glue does per call)
against the solve alone — allocation and staging overheads (profiles/r1_measurements.md).  Run from the repo root on the GPU box."""
import sys, json, time
sys.path.insert(0, "semiseparablematrices.jl_b200"); sys.path.insert(0, ".")
import numpy as np, torch, ssm_b200 as ssm, oracle
from ssm_b200 import chunked
ctx = ssm.default_context(0)
for n in (1024, 16384, 262144, 1000000):
    l, u, r = 4, 4, 2
    bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=3, dist=1)
    bands, U, V, b = bands[0], U[0], V[0], b[0]
    def call():
        A = chunked.LargeAlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
        x = A.solve(b)
        del A
        return x
    call(); call()
    t0 = time.perf_counter()
    k = 10
    for _ in range(k): x = call()
    dt = (time.perf_counter() - t0) / k * 1e3
    A = chunked.LargeAlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    A.solve(b)
    t0 = time.perf_counter()
    for _ in range(k): x = A.solve(b)
    dt2 = (time.perf_counter() - t0) / k * 1e3
    print(json.dumps({"n": n, "create_set_solve_destroy_ms": round(dt, 3), "solve_only_host_b_ms": round(dt2, 3), "relres": A.residual(x, b)}))
