"""CPU tier: the N>1 host logic of the batch sharder with world_size-2 gloo (no GPU): shard cuts, the barrier + max-over-ranks
timing reduction bench.py uses, and the host gather — with the CPU oracle standing in for the device so the gathered
per-shard solutions can be compared with the unsharded solve."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "semiseparablematrices.jl_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def test_shard_ranges_partition_and_match_c_abi():
    from ssm_b200 import sharder
    for nb in (1, 31, 32, 33, 1000, 65536, 1 << 20, 12345):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharder.shard_range(nb, g, world) for g in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nb
            for g in range(world):
                assert cuts[g] == sharder.shard_range_c(nb, g, world)
                assert cuts[g][0] % 32 == 0 and cuts[g][0] <= cuts[g][1]
                if g: assert cuts[g][0] == cuts[g - 1][1]
            if nb >= 64 * world:        # balanced to within one tile
                sizes = [b - a for a, b in cuts]
                assert max(sizes) - min(sizes) <= 32 + nb % 32


def test_group_plan_covers_every_system_once_and_balances():
    from ssm_b200 import sharder
    groups = [(256, 524288), (512, 262144), (1024, 131072), (2048, 65536), (4096, 32768), (8192, 32768)]   # BASELINE configs[4]
    total = sum(n * nb for n, nb in groups)
    for world in (1, 2, 3, 4, 8):
        plan = sharder.plan_groups(groups, world)
        assert len(plan) == world
        cover = {}
        for pieces in plan:
            for gi, s0, s1 in pieces:
                assert s0 % 32 == 0 and s1 > s0
                cover.setdefault(gi, []).append((s0, s1))
        for gi, (n, nb) in enumerate(groups):
            c = sorted(cover[gi])
            assert c[0][0] == 0 and c[-1][1] == nb and all(c[i][1] == c[i + 1][0] for i in range(len(c) - 1))
            assert all(b - a >= 16384 for a, b in c)                  # pieces stay large enough to fill a GPU
        loads = [sum(groups[gi][0] * (b - a) for gi, a, b in pieces) for pieces in plan]
        assert max(loads) <= 1.15 * total / world


def _worker(rank, world, port, nb, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import oracle
    from ssm_b200 import sharder
    job = sharder.Job(backend="gloo")
    s0, s1 = job.shard(nb)
    n, l, u, r = 48, 2, 2, 2
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=20261020)
    x = oracle.ab_solve(bands[s0:s1], U[s0:s1], V[s0:s1], b[s0:s1], l, u)     # the oracle stands in for the device here
    job.barrier()
    t = job.max_over_ranks(1.0 + rank)                                         # "device time" of this rank
    tot = job.sum_over_ranks(float(s1 - s0))
    parts = job.gather((s0, s1, x))
    if rank == 0:
        full = np.empty((nb, n))
        for a, e, xs in parts:
            full[a:e] = xs
        ref = oracle.ab_solve(bands, U, V, b, l, u)
        q.put((t, tot, bool(np.array_equal(full, ref))))
    job.close()


def test_world2_gloo_shards_gather_and_max_time():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, nb = 2, 200
    procs = [ctx.Process(target=_worker, args=(rk, world, port, nb, q)) for rk in range(world)]
    for p in procs: p.start()
    t, tot, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0 and tot == nb and same
