"""Multi-GPU sharding of the 3c2e tensor over auxiliary shells.

Every (ab|c) block is independent, so there is NO collective on the data path: rank r owns a
contiguous range of the class-major ordered auxiliary shells (plan_aux_shards) chosen by prefix sums of
the per-shell cost model and writes its own (rows, naux_local) slab.  Only when a caller asks for the whole tensor on every device is
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


def plan_aux_shards(orb_shells, aux_shells, nranks, class_major=True):
    """Sharding plan of the (ab|c) tensor: ``{"order", "aux", "ranges", "columns", "costs"}``.

    ``order`` permutes the input auxiliary shells (class-major by default, shells.class_major_order: better store
    locality and 1-2 classes per shard), ``aux`` is the permuted list every rank builds its basis from, ``ranges`` the
    contiguous shell range of every rank (balanced by the measured-rate cost model, workmodel.aux_shell_costs) and
    ``columns`` the matching column ranges.  Columns of the distributed tensor follow ``aux``;
    shells.function_permutation(aux_shells, order) maps them back to the input order."""
    from . import workmodel
    from .shells import class_major_order

    order = class_major_order(aux_shells) if class_major else np.arange(len(aux_shells))
    aux = [aux_shells[i] for i in order]
    costs = workmodel.aux_shell_costs(orb_shells, aux)
    ranges = partition_aux_shells(costs, nranks)
    return {"order": order, "aux": aux, "ranges": ranges, "columns": shard_columns([2 * s.L + 1 for s in aux], ranges),
            "costs": costs}


def shard_columns(aux_sizes, ranges):
    """Column (auxiliary function) ranges of the shards."""
    off = np.concatenate([[0], np.cumsum(aux_sizes)])
    return [(int(off[b]), int(off[e])) for (b, e) in ranges]


def allgather_slabs(local_slab, col_ranges, group=None, out=None, row_block=None):
    """Assemble the full (rows, naux) tensor on every rank from the per-rank (rows, ncol_r) slabs.

    The slabs are column blocks of different widths of a row-major matrix, so they are not contiguous in the
    result.  Each rank's slab is padded to the widest one and ONE ``all_gather_into_tensor`` (NCCL all-gather over
    NVLink / NVSwitch; gloo in the CPU tests) moves a block of rows from all ranks at once; the blocks are then
    copied into their column ranges.  ``row_block`` bounds the staging buffer (world x row_block x wmax doubles;
    default: at most ~2 GiB).  Used only on request -- the scaling benchmark times it separately, outside the
    scaling figure."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = local_slab.shape[0]
    ncol = col_ranges[-1][1]
    wmax = max(c1 - c0 for (c0, c1) in col_ranges)
    assert local_slab.shape[1] == col_ranges[rank][1] - col_ranges[rank][0]
    full = out if out is not None else torch.empty((rows, ncol), dtype=local_slab.dtype, device=local_slab.device)
    if row_block is None:
        row_block = max(1, min(rows, (1 << 28) // max(1, world * wmax)))
    send = torch.zeros((row_block, wmax), dtype=local_slab.dtype, device=local_slab.device)
    recv = torch.empty((world, row_block, wmax), dtype=local_slab.dtype, device=local_slab.device)
    w_me = local_slab.shape[1]
    for r0 in range(0, rows, row_block):
        n = min(row_block, rows - r0)
        send[:n, :w_me] = local_slab[r0:r0 + n]
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
        for r in range(world):
            c0, c1 = col_ranges[r]
            full[r0:r0 + n, c0:c1] = recv[r, :n, :c1 - c0]
    return full
