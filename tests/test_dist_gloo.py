"""Multi-rank host logic on CPU: cost-balanced aux-shell partition and the on-request all-gather
(world_size 2, gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sympleints_b200 import dist as sdist
from sympleints_b200 import synthetic, workmodel


def test_partition_is_contiguous_balanced_and_complete():
    shells, sites = synthetic.orbital_basis(natoms=10, seed=3)
    aux = synthetic.aux_basis(sites, seed=4)
    costs = workmodel.aux_shell_costs(shells, aux)
    for n in (1, 2, 3, 4, 8):
        r = sdist.partition_aux_shells(costs, n)
        assert r[0][0] == 0 and r[-1][1] == len(aux)
        assert all(r[i][1] == r[i + 1][0] for i in range(n - 1)) and all(b < e for b, e in r)
        loads = [sum(costs[b:e]) for b, e in r]
        assert max(loads) / (sum(loads) / n) < 1.06
    with pytest.raises(ValueError):
        sdist.partition_aux_shells(costs[:3], 4)
    # shard work adds up to the whole
    full = workmodel.tensor_3c2e_work(shells, aux)
    parts = [workmodel.tensor_3c2e_work(shells, aux, rg) for rg in sdist.partition_aux_shells(costs, 4)]
    assert sum(p["integrals"] for p in parts) == full["integrals"]
    assert sum(p["flops"] for p in parts) == full["flops"]


def test_class_major_shard_plan():
    from sympleints_b200.shells import class_major_order, function_permutation

    shells, sites = synthetic.orbital_basis(natoms=10, seed=3)
    aux = synthetic.aux_basis(sites, seed=4)
    plan = sdist.plan_aux_shards(shells, aux, 4)
    order = plan["order"]
    assert sorted(order) == list(range(len(aux))) and np.array_equal(order, class_major_order(aux))
    Ls = [s.L for s in plan["aux"]]
    assert Ls == sorted(Ls)
    # stable: shells of one class keep their input order
    for lc in set(Ls):
        ks = [k for k in order if aux[k].L == lc]
        assert ks == sorted(ks)
    assert plan["ranges"][0][0] == 0 and plan["ranges"][-1][1] == len(aux)
    assert plan["columns"][-1][1] == sum(2 * s.L + 1 for s in aux)
    loads = [sum(plan["costs"][b:e]) for b, e in plan["ranges"]]
    assert max(loads) / (sum(loads) / 4) < 1.06
    # the measured-rate weights only rescale per class
    plain = workmodel.aux_shell_costs(shells, plan["aux"], weights=None)
    assert all(abs(c / p - workmodel.LC_TIME_WEIGHT[s.L]) < 1e-12 for c, p, s in zip(plan["costs"], plain, plan["aux"]))
    # column map back to the input order
    idx = function_permutation(aux, order)
    off = np.concatenate([[0], np.cumsum([2 * s.L + 1 for s in aux])])
    assert sorted(idx) == list(range(off[-1]))
    k = order[5]
    c0 = sum(2 * aux[i].L + 1 for i in order[:5])
    assert list(idx[c0:c0 + 2 * aux[k].L + 1]) == list(range(off[k], off[k + 1]))
    same = sdist.plan_aux_shards(shells, aux, 4, class_major=False)
    assert list(same["order"]) == list(range(len(aux)))


def test_benchmark_system_totals_match_survey():
    shells, sites = synthetic.orbital_basis()
    aux = synthetic.aux_basis(sites)
    assert (len(shells), sum(2 * s.L + 1 for s in shells), len(aux), sum(2 * s.L + 1 for s in aux)) == (300, 1300, 1200, 5000)
    w = workmodel.tensor_3c2e_work(shells, aux)
    assert w["integrals"] == 4245750000 and w["triples"] == 54180000 and w["prim_triples"] == 295560000
    assert abs(w["flops"] / 3.748e11 - 1) < 1e-3 and w["bytes"] == 33966000000


def _worker(rank, world, port, rows, sizes, ranges, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cols = sdist.shard_columns(sizes, ranges)
    c0, c1 = cols[rank]
    g = torch.Generator().manual_seed(7)
    ref = torch.rand((rows, cols[-1][1]), generator=g, dtype=torch.float64)   # same on both ranks
    full = sdist.allgather_slabs(ref[:, c0:c1].clone(), cols)
    blocked = sdist.allgather_slabs(ref[:, c0:c1].clone(), cols, row_block=5)   # several all-gathers + a ragged tail
    q.put((rank, bool(torch.equal(full, ref)) and bool(torch.equal(blocked, ref))))
    dist.destroy_process_group()


def test_allgather_of_unequal_slabs_gloo_world2():
    sizes = [1, 3, 5, 7, 9, 11, 1, 3]
    ranges = sdist.partition_aux_shells([s * 1.0 for s in sizes], 2)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 13, sizes, ranges, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
