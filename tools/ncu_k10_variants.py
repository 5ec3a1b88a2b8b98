#!/usr/bin/env python
"""Stage 1 of the chunked path under its tuning options, one solve each (for a light `ncu -k regex:abc_qr --metrics ...` pass).
usage: python tools/ncu_k10_variants.py "abc_gs=8,abc_sb=0" "abc_gs=4,abc_sb=0" ..."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402

import ssm_b200 as ssm  # noqa: E402
from ssm_b200 import chunked  # noqa: E402

ctx = ssm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
n = 1 << 24
A = chunked.LargeAlmostBandedMatrix.generate(n, 4, 4, 2, 20261022, dist=1, ctx=ctx)
b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), 20261022)
x = torch.empty_like(b)
for spec in sys.argv[1:] or ["abc_gs=8"]:
    for kv in spec.split(","):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    for _ in range(2):
        A.solve(b, out=x, sync=True)
    torch.cuda.synchronize()
    print("ok", spec, A.residual(x, b))
