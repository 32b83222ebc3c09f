"""Time every shard of an N-way aux partition on one GPU, for the atom-major aux order and for an L-sorted one."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sympleints_b200
from sympleints_b200 import dist as sdist
from sympleints_b200 import synthetic, workmodel

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = sympleints_b200.get_engine(0, 4, 5)
shells, sites = synthetic.orbital_basis(natoms=50)
aux0 = synthetic.aux_basis(sites)
ob = eng.basis(shells)
res = {}
for name, aux in (("atom_major", aux0), ("l_sorted", sorted(aux0, key=lambda s: s.L)), ("l_sorted_desc", sorted(aux0, key=lambda s: -s.L))):
    ab = eng.basis(aux)
    costs = workmodel.aux_shell_costs(shells, aux)
    ranges = sdist.partition_aux_shells(costs, N)
    times = []
    for my in ranges:
        ncol = eng.aux_cols(ab, *my)
        out = torch.empty((ob.packed_rows, ncol), dtype=torch.float64, device=eng.tdev)
        for _ in range(3):
            eng.int3c2e(ob, ab, shell_range=my, layout="packed", out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record()
        for _ in range(10):
            eng.int3c2e(ob, ab, shell_range=my, layout="packed", out=out)
        e1.record()
        torch.cuda.synchronize()
        times.append((e0.elapsed_time(e1) / 10, (eng.launch_count() - l0) // 10, sorted(set(s.L for s in aux[my[0]:my[1]]))))
        del out
    res[name] = times
    print(name, "max %.3f sum %.3f" % (max(t[0] for t in times), sum(t[0] for t in times)))
    for t in times:
        print("   %.3f ms  launches %d  Lc %s" % t)
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/shard_study_%d.json" % N, "w"), indent=1)
eng.close()
