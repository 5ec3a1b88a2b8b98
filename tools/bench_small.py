#!/usr/bin/env python
"""qr / ldiv! timing of small batches and single matrices per kernel variant (tuning aid for the small-batch selection rules)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    sys.path.insert(0, str(p))
import numpy as np, torch
import ssm_b200 as ssm

def timeit(fn, steps=10, warm=3):
    st = torch.cuda.current_stream()
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

ctx = ssm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
shapes = [(2, 2, 2)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (l, u, r) in shapes:
    for n, nb in [(1024, 1), (16384, 1), (8192, 32), (8192, 1024), (8192, 4096), (8192, 8192), (8192, 16384), (2048, 8192), (1024, 16384), (256, 4096)]:
        A = ssm.AlmostBandedMatrix.generate(n, l, u, r, nb, 7, ctx=ctx)
        b = ssm.Vec(ctx, n, nb).generate(7); x = ssm.Vec(ctx, n, nb)
        row = {"shape": (l, u, r), "n": n, "nb": nb}
        ref = None
        for v in (0, 1, 4):
            ctx.set_option("ab_variant", v)
            F = ssm.qr(A)
            row[f"qr_v{v}_ms"] = round(timeit(lambda: F.refactor(A)), 4)
            g = F.get()
            if ref is None: ref = g
            else: row[f"v{v}_bit_identical"] = all(np.array_equal(p, q) for p, q in zip(g, ref))
            if v == 0: row["ldiv_ms"] = round(timeit(lambda: F.solve(b, out=x)), 4)
            del F
        ctx.set_option("ab_variant", 0)
        print(json.dumps(row))
        del A, b, x
