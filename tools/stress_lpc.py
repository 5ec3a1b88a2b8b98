#!/usr/bin/env python
"""Stress of the lane-per-chunk stage 1 (four mbarrier-coupled warps per CTA): many solves of one system, alone and beside a second
stream that keeps HBM and the SMs busy; every solution must equal the first steady-state one bit for bit.
usage: python tools/stress_lpc.py [n] [solves]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402

import ssm_b200 as ssm  # noqa: E402
from ssm_b200 import chunked  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
st0 = torch.cuda.Stream()
ctx = ssm.Context(0, stream=st0.cuda_stream)
A = chunked.LargeAlmostBandedMatrix.generate(n, 4, 4, 2, 20261022, dist=1, ctx=ctx)
with torch.cuda.stream(st0):
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    x = torch.empty_like(b)
A.generate_rhs(b, 7)
for _ in range(3):
    A.solve(b, out=x, sync=True)
ref = x.clone()
print("chunks", A.chunks, "relres", A.residual(x, b))
st1 = torch.cuda.Stream()
t1 = torch.empty(1 << 26, dtype=torch.float64, device="cuda"); t2 = torch.empty_like(t1)
m1 = torch.randn(4096, 4096, device="cuda"); m2 = torch.empty_like(m1)
bad = 0
for it in range(solves):
    mode = it % 3
    if mode == 1:
        with torch.cuda.stream(st1):
            t2.copy_(t1)
    elif mode == 2:
        with torch.cuda.stream(st1):
            torch.matmul(m1, m1, out=m2)
    with torch.cuda.stream(st0):
        x.zero_()
    A.solve(b, out=x, sync=False)
    if it % 25 == 24:
        torch.cuda.synchronize()
        if not torch.equal(x, ref):
            bad += 1
            print("MISMATCH at", it, float((x - ref).abs().max()))
torch.cuda.synchronize()
print("solves", solves, "mismatches", bad)
sys.exit(1 if bad else 0)
