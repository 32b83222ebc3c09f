"""Measured time per model flop of the 3c2e tensor by auxiliary angular momentum (class-major aux order), and the
shard times of the N-way partitions balanced with these weights.  Writes gpurun_out/lc_weights.json."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sympleints_b200
from sympleints_b200 import dist as sdist
from sympleints_b200 import synthetic, workmodel

eng = sympleints_b200.get_engine(0, 4, 5)
shells, sites = synthetic.orbital_basis(natoms=50)
aux = sorted(synthetic.aux_basis(sites), key=lambda s: s.L)
ob, ab = eng.basis(shells), eng.basis(aux)
flops = np.array(workmodel.aux_shell_costs(shells, aux, weights=None), dtype=float)
Ls = np.array([s.L for s in aux])


def time_range(my, reps=10):
    ncol = eng.aux_cols(ab, *my)
    out = torch.empty((ob.packed_rows, ncol), dtype=torch.float64, device=eng.tdev)
    for _ in range(3):
        eng.int3c2e(ob, ab, shell_range=my, layout="packed", out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.int3c2e(ob, ab, shell_range=my, layout="packed", out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"per_lc": {}}
w = np.ones(6)
for lc in range(6):
    idx = np.nonzero(Ls == lc)[0]
    my = (int(idx[0]), int(idx[-1]) + 1)
    t = time_range(my)
    res["per_lc"][lc] = {"ms": t, "model_flops": float(flops[idx].sum())}
    w[lc] = t / flops[idx].sum()
w /= (w * np.array([flops[Ls == lc].sum() for lc in range(6)])).sum() / flops.sum()
res["weights"] = w.tolist()
print("weights", np.round(w, 4).tolist())
for N in (1, 2, 4, 8):
    ranges = sdist.partition_aux_shells(flops * w[Ls], N)
    ts = [time_range(my) for my in ranges]
    res["N%d" % N] = ts
    print(N, "max %.3f sum %.3f" % (max(ts), sum(ts)), np.round(ts, 3).tolist())
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/lc_weights.json", "w"), indent=1)
eng.close()
