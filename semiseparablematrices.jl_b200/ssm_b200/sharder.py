"""Batch sharder: independent systems are split across the GPUs of one box (no collective on the data path).

Two ways to drive N GPUs:
* one process, N host threads inside the library — ``solve_host_sharded`` (``ssm_ab_solve_host_sharded``);
* one process per GPU (``torchrun``) — every rank owns ``shard_range(nb, rank, world)`` and only the timing /
  parity summaries travel through ``torch.distributed`` (``Job``), NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check, c_vp

TILE = 32


def shard_range(nb: int, rank: int, world: int) -> tuple[int, int]:
    """systems [s0, s1) owned by shard ``rank`` of ``world``; cuts fall on multiples of the 32-system tile.
    Pure-Python twin of ``ssm_shard_range`` (tests check they agree)."""
    if world < 1 or not (0 <= rank < world) or nb < 0:
        raise ValueError("bad shard arguments")

    def cut(k):
        if k >= world:
            return nb
        c = (nb * k) // world
        return c - c % TILE
    return cut(rank), cut(rank + 1)


def shard_range_c(nb: int, rank: int, world: int) -> tuple[int, int]:
    s0, s1 = C.c_int64(), C.c_int64()
    check(_lib.lib().ssm_shard_range(nb, rank, world, C.byref(s0), C.byref(s1)), "ssm_shard_range")
    return s0.value, s1.value


def solve_host_sharded(devices, n, l, u, r, bands, U, V, b, x=None):
    """``A \\ b`` for a host-resident batch on several GPUs of this box (one host thread per GPU inside the library)."""
    from .almostbanded import _as_ptr
    nb = int(b.shape[0])
    if x is None:
        x = np.empty((nb, n))
    dev = (C.c_int * len(devices))(*devices)
    pb, kb = _as_ptr(bands, (nb, n, l + u + 1), "bands")
    pu, ku = _as_ptr(U, (nb, r, n), "U")
    pv, kv = _as_ptr(V, (nb, n, r), "V")
    pr, kr = _as_ptr(b, (nb, n), "b")
    px, kx = _as_ptr(x, (nb, n), "x")
    check(_lib.lib().ssm_ab_solve_host_sharded(len(devices), dev, n, l, u, r, nb, c_vp(pb), c_vp(pu), c_vp(pv), c_vp(pr), c_vp(px)),
          "ssm_ab_solve_host_sharded")
    return x


class Job:
    """The one-process-per-GPU plumbing of bench.py: rank / world from the launcher's environment, a barrier, the
    max-over-ranks reduction of a device-timed duration and a host gather of per-shard summaries.  ``backend`` is
    "nccl" on the GPU box and "gloo" in the CPU tests; with world == 1 nothing is initialised."""

    def __init__(self, backend: str = "nccl", device=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend
        self._dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            if not dist.is_initialized():
                dist.init_process_group(backend, **kw)
            self._dist = dist

    def shard(self, nb_total: int) -> tuple[int, int]:
        return shard_range(nb_total, self.rank, self.world)

    def barrier(self):
        if self._dist is not None:
            self._dist.barrier()

    def max_over_ranks(self, value: float) -> float:
        if self._dist is None:
            return float(value)
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value: float) -> float:
        if self._dist is None:
            return float(value)
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, obj):
        """host gather of small per-shard results (rank 0 gets the list, others None)"""
        if self._dist is None:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self._dist.gather_object(obj, out, dst=0)
        return out

    def close(self):
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
            self._dist = None


def plan_groups(groups, world: int, min_piece_systems: int = 16384):
    """Work plan for a sweep over several batch shapes (BASELINE config 5: n = 256 .. 8192, 1M systems).

    ``groups`` = [(n, nb), ...]; cost of a piece = n * systems (the sweeps are O(n) per system and HBM bound).  Splitting EVERY
    group across all ranks leaves each rank with batches too small to fill a B200 at large n (a sweep is a serial chain per
    warp: 4,096 systems are 128 warps for 148 SMs), so whole groups are the unit: longest-processing-time-first assignment, and
    only the largest piece of the most loaded rank is halved (never below ``min_piece_systems``, cuts on 32-system tiles) while
    that lowers the makespan.  Returns ``plan[rank] = [(group_index, s0, s1), ...]`` — every system exactly once."""
    pieces = [(gi, 0, nb) for gi, (n, nb) in enumerate(groups)]

    def cost(p):
        return groups[p[0]][0] * (p[2] - p[1])

    def lpt(ps):
        loads, plan = [0] * world, [[] for _ in range(world)]
        for p in sorted(ps, key=cost, reverse=True):
            r = min(range(world), key=lambda i: loads[i])
            plan[r].append(p); loads[r] += cost(p)
        return plan, loads

    plan, loads = lpt(pieces)
    if world == 1:
        return [sorted(plan[0])]
    neutral = 0
    for _ in range(8 * world + len(groups)):
        worst = max(range(world), key=lambda i: loads[i])
        best = None
        for cand in sorted(plan[worst], key=cost, reverse=True):
            gi, s0, s1 = cand
            mid = s0 + (((s1 - s0) // 2) // TILE) * TILE
            if mid - s0 < min_piece_systems or s1 - mid < min_piece_systems:
                continue
            trial = [p for p in pieces if p != cand] + [(gi, s0, mid), (gi, mid, s1)]
            tplan, tloads = lpt(trial)
            if best is None or max(tloads) < max(best[2]):
                best = (trial, tplan, tloads)
            if max(tloads) < max(loads):
                break
        if best is None or max(best[2]) > max(loads):
            break
        if max(best[2]) == max(loads):
            neutral += 1
            if neutral > 3:
                break
        else:
            neutral = 0
        pieces, plan, loads = best
    return [sorted(p) for p in plan]
