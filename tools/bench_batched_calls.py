"""What one `qr(A)` + `ldiv!(F, b)` on ONE matrix costs through the batched objects with HOST arrays — upload + qr + get + ldiv! + destroy, as
the Julia glue's `qr(A)` / `F \\ b` do per call — with and without the per-device free list of device blocks (SSM_POOL=0 disables it;
run the script once each way).  Run from the repo root on the GPU box; numbers go to profiles/r1_measurements.md."""
import sys, json, time, os
sys.path.insert(0, "semiseparablematrices.jl_b200"); sys.path.insert(0, ".")
import numpy as np, torch, ssm_b200 as ssm, oracle
ctx = ssm.default_context(0)
for n, nb in ((1024, 1), (16384, 1), (1024, 32), (1024, 1024)):
    l, u, r = 2, 2, 2
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=3)
    def call():
        A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
        F = ssm.qr(A)
        fac = F.get()
        x = F.ldiv(b.copy())
        del F, A
        return x
    call(); call()
    k = 20
    t0 = time.perf_counter()
    for _ in range(k): x = call()
    dt = (time.perf_counter() - t0) / k * 1e3
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    F = ssm.qr(A)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): F.refactor(A)
    ctx.sync()
    dt2 = (time.perf_counter() - t0) / k * 1e3
    print(json.dumps({"n": n, "systems": nb, "pool": os.environ.get("SSM_POOL", "1"), "upload_qr_get_ldiv_destroy_ms": round(dt, 3),
                      "qr_kernel_only_ms": round(dt2, 3), "pool_stats": ssm.pool_stats(0)}))
