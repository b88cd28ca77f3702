"""Sharding of a problem array across ranks (one process per GPU) and the host-side gather.

The gap-fill problems are independent, so the data path has no collective: rank r solves its contiguous slice and
rank 0 gathers results and Pair records in input order.  torch.distributed is used for the gather only (gloo on
CPU in the tests, any backend that supports gather_object otherwise); the bench uses NCCL only for timing."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n problems owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def solve_sharded(solver, problems, rank, world, dist=None):
    """Every rank solves its slice with `solver.solve`; rank 0 returns (results, pairs, pair_off) of the whole
    array in input order, the other ranks return None."""
    lo, hi = shard_range(len(problems), rank, world)
    res, pairs, off = solver.solve(problems[lo:hi])
    if world == 1:
        return res, pairs, off
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((res, pairs, off), gathered, dst=0)
    if rank != 0:
        return None
    all_res = np.concatenate([g[0] for g in gathered])
    all_pairs = np.concatenate([g[1] for g in gathered])
    offs, base = [], 0
    for g in gathered:
        offs.append(g[2][:-1] + base)
        base += int(g[2][-1])
    all_off = np.concatenate(offs + [np.array([base], dtype=np.int64)])
    return all_res, all_pairs, all_off
