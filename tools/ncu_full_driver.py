#!/usr/bin/env python
"""The three dominant kernels, a few launches each, for ONE `ncu --set full` capture per round (K1 ab_qr_kernel at config 2,
K6 bps_qr_kernel at config 3, K10 stage 1 abc_qr_kernel at config 4).  usage: python tools/ncu_full_driver.py [k1|k6|k10]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402

import ssm_b200 as ssm  # noqa: E402
from ssm_b200 import bps, chunked  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "k1"
ctx = ssm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
if which == "k1":
    A = ssm.AlmostBandedMatrix.generate(1024, 2, 2, 2, 65536, 20261020, ctx=ctx)
    F = ssm.qr(A)
    for _ in range(3):
        F.refactor(A)
elif which == "k6":
    A = bps.BandedPlusSemiseparableMatrix.generate(4096, 3, 3, 2, 2, 16384, 20261021, ctx=ctx)
    F = bps.qr(A)
    for _ in range(3):
        F.refactor(A)
else:
    n = 1 << 24
    A = chunked.LargeAlmostBandedMatrix.generate(n, 4, 4, 2, 20261022, dist=1, ctx=ctx)
    b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), 20261022)
    x = torch.empty_like(b)
    for _ in range(3):
        A.solve(b, out=x, sync=True)
torch.cuda.synchronize()
print("ok", which, ctx.launches)
