"""Multi-GPU sharding of the 3c2e tensor over auxiliary shells.

Every (ab|c) block is independent, so there is NO collective on the data path: rank r owns a
contiguous range of auxiliary shells chosen by prefix sums of the per-shell cost model and writes
its own (rows, naux_local) slab.  Only when a caller asks for the whole tensor on every device is
an all-gather issued (NCCL over NVLink via torch.distributed; gloo on CPU in the tests).
"""
import numpy as np


def partition_aux_shells(costs, nranks):
    """Contiguous shell ranges [(begin, end), ...] with balanced cumulative cost.

    Cut k is placed where the prefix sum crosses k/nranks of the total (nearest shell boundary);
    every rank gets at least one shell when len(costs) >= nranks."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if nranks <= 0 or n < nranks:
        raise ValueError("need at least one auxiliary shell per rank")
    pref = np.concatenate([[0.0], np.cumsum(costs)])
    cuts = [0]
    for k in range(1, nranks):
        target = pref[-1] * k / nranks
        i = int(np.searchsorted(pref, target))
        if i > 0 and abs(pref[i - 1] - target) <= abs(pref[min(i, n)] - target):
            i -= 1
        i = max(i, cuts[-1] + 1)
        i = min(i, n - (nranks - k))
        cuts.append(i)
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(nranks)]


def shard_columns(aux_sizes, ranges):
    """Column (auxiliary function) ranges of the shards."""
    off = np.concatenate([[0], np.cumsum(aux_sizes)])
    return [(int(off[b]), int(off[e])) for (b, e) in ranges]


def allgather_slabs(local_slab, col_ranges, group=None):
    """Assemble the full (rows, naux) tensor on every rank from the per-rank (rows, ncol_r) slabs.

    Slabs have unequal widths, so each rank's slab is broadcast from its owner into the matching
    column range (one collective per rank; with NCCL these are NVLink broadcasts).  Used only on
    request -- the scaling benchmark never calls it."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = local_slab.shape[0]
    ncol = col_ranges[-1][1]
    full = torch.empty((rows, ncol), dtype=local_slab.dtype, device=local_slab.device)
    for r in range(world):
        c0, c1 = col_ranges[r]
        buf = local_slab.contiguous() if r == rank else torch.empty((rows, c1 - c0), dtype=local_slab.dtype, device=local_slab.device)
        dist.broadcast(buf, src=r, group=group)
        full[:, c0:c1] = buf
    return full
