"""Schwarz bounds on the benchmark system: time of si_schwarz and the fraction of shell triples (and of the model
flops) whose bound Q_ab * Q_c exceeds a threshold.  Writes gpurun_out/schwarz_study.json."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sympleints_b200
from sympleints_b200 import synthetic, workmodel

eng = sympleints_b200.get_engine(0, 4, 5)
shells, sites = synthetic.orbital_basis(natoms=50)
aux = synthetic.aux_basis(sites)
ob, ab = eng.basis(shells), eng.basis(aux)
eng.schwarz(ob)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
D, Q = eng.schwarz(ob)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
Q = Q.cpu().numpy()
M = eng.int2c2e(ab).cpu().numpy()
offa = np.concatenate([[0], np.cumsum([2 * s.L + 1 for s in aux])])
Qc = np.array([np.sqrt(np.abs(np.diag(M)[offa[k]:offa[k + 1]]).max()) for k in range(len(aux))])
iu = np.tril_indices(len(shells))
qab = Q[iu]                                   # unique pairs i >= j
La = np.array([max(shells[i].L, shells[j].L) for i, j in zip(*iu)])
Lb = np.array([min(shells[i].L, shells[j].L) for i, j in zip(*iu)])
Kab = np.array([len(shells[i].exps) * len(shells[j].exps) for i, j in zip(*iu)])
Lc = np.array([s.L for s in aux])
# model flops per (pair, aux) tuple
cost = np.zeros((len(qab), len(aux)))
for la in range(5):
    for lb in range(la + 1):
        for lc in range(6):
            fp, fc = workmodel.flops_3c2e(la, lb, lc)
            sel = (La == la) & (Lb == lb)
            cost[np.ix_(sel, Lc == lc)] = (Kab[sel] * fp + fc)[:, None]
bound = qab[:, None] * Qc[None, :]
rows = []
for eps in (1e-14, 1e-12, 1e-10, 1e-8, 1e-6):
    keep = bound > eps
    rows.append({"eps": eps, "surviving_tuple_fraction": float(keep.mean()), "surviving_flop_fraction": float((cost * keep).sum() / cost.sum())})
    print(rows[-1])
out = {"schwarz_ms": ms, "pairs": int(len(qab)), "aux_shells": int(len(aux)), "q_ab_min": float(qab.min()), "q_ab_max": float(qab.max()),
       "thresholds": rows}
print(out["schwarz_ms"], out["q_ab_min"], out["q_ab_max"])
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(out, open("gpurun_out/schwarz_study.json", "w"), indent=1)
eng.close()
