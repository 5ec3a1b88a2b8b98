#!/usr/bin/env python
"""One pass of every timed operation of bench.py's configs, for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
(the numbers behind `roofline.traffic`; parsed here by tools/ncu_traffic_parse.py into profiles/traffic.json).  Prints the ordered list of
phases with the library's launch counter before / after each, so that the parser can attribute launches to operations."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402

import ssm_b200 as ssm  # noqa: E402
from ssm_b200 import bps, chunked  # noqa: E402

ctx = ssm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
phases = []


def phase(name, fn):
    l0 = ctx.launches
    fn()
    torch.cuda.synchronize()
    phases.append({"phase": name, "first_launch": l0, "launches": ctx.launches - l0})


# config 2
n, l, u, r, nb = 1024, 2, 2, 2, 65536
A = b = x = F = None
def mk2():
    global A, b, x, F
    A = ssm.AlmostBandedMatrix.generate(n, l, u, r, nb, 20261020, ctx=ctx); b = ssm.Vec(ctx, n, nb).generate(20261020); x = ssm.Vec(ctx, n, nb); F = ssm.qr(A)
phase("c2 setup (generate, first qr)", mk2)
phase("c2 qr", lambda: F.refactor(A))
phase("c2 ldiv", lambda: F.solve(b, out=x))
phase("c2 fused", lambda: A.solve(b, out=x))
del A, b, x, F
# config 3
n, l, m, r, p, nb = 4096, 3, 3, 2, 2, 16384
def mk3():
    global A, b, x, F
    A = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, 20261021, ctx=ctx); b = ssm.Vec(ctx, n, nb).generate(20261021, array_id=5); x = ssm.Vec(ctx, n, nb); F = bps.qr(A)
phase("c3 setup (generate + U'U, first qr)", mk3)
phase("c3 qr", lambda: F.refactor(A))
phase("c3 solve", lambda: F.solve(b, out=x))
del A, b, x, F
# config 4
n, l, u, r = 1 << 24, 4, 4, 2
def mk4():
    global A, b, x
    A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, 20261022, dist=1, ctx=ctx)
    b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), 20261022); x = torch.empty_like(b)
    A.solve(b, out=x, sync=True)
    A.solve(b, out=x, sync=True)       # the second solve of the same operands builds the interleaved copy of the lane-per-chunk stage 1
phase("c4 setup (generate, first two solves)", mk4)
phase("c4 solve", lambda: A.solve(b, out=x, sync=True))
print(json.dumps(phases))
