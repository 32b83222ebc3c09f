import sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sympleints_b200, torch
from sympleints_b200 import synthetic
from oracle import ints, boys as oboys
eng = sympleints_b200.get_engine(0, 4, 5, boys_table=oboys.table())
shells, sites = synthetic.orbital_basis(natoms=3, seed=5)
aux = synthetic.aux_basis(sites, seed=6)
ob, ab = eng.basis(shells), eng.basis(aux)
T = torch.full((ob.nbf_sph, ob.nbf_sph, ab.nbf_sph), float('nan'), dtype=torch.float64, device='cuda')
eng.int3c2e(ob, ab, out=T); T = T.cpu().numpy()
print("nan count", np.isnan(T).sum(), "of", T.size)
offo = np.concatenate([[0], np.cumsum([2*s.L+1 for s in shells])]); offa = np.concatenate([[0], np.cumsum([2*s.L+1 for s in aux])])
bad = {}
r2 = np.random.default_rng(1)
for _ in range(400):
    i, j, k = r2.integers(len(shells)), r2.integers(len(shells)), r2.integers(len(aux))
    a, b, c = shells[i], shells[j], aux[k]
    o = ints.int3c2e_sph(a.L, b.L, c.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center, c.exps, c.coeffs, c.center)
    blk = T[offo[i]:offo[i+1], offo[j]:offo[j+1], offa[k]:offa[k+1]].ravel()
    err = np.abs(blk - o).max() / max(np.abs(o).max(), 1e-300)
    key = (max(a.L,b.L), min(a.L,b.L), c.L)
    if not (err < 1e-11):
        bad.setdefault(key, []).append((i, j, k, len(a.exps), len(b.exps), float(np.abs(o).max()), float(np.nanmax(np.abs(blk))), float(err)))
for k, v in sorted(bad.items()):
    print(k, len(v), v[:2])
print("bad classes", len(bad))
