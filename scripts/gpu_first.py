"""First contact with the GPU: per-block and small batched parity against the oracle."""
import sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sympleints_b200
from sympleints_b200 import synthetic
from oracle import ints, boys as oboys

eng = sympleints_b200.get_engine(0, 4, 5)
tab = eng.boys_table()
ref = oboys.table()
print("boys table own vs oracle(quad): max rel", np.abs(tab[:ref.shape[0]] / ref - 1).max() if tab.shape == ref.shape else (tab.shape, ref.shape))
xs = np.concatenate([[0.0], np.random.default_rng(0).uniform(0, 40, 2000)])
for n in (0, 5, 13):
    print("boys n=%d gpu vs oracle: %.2e" % (n, np.abs(eng.boys(n, xs) / oboys.boys_vec(n, xs) - 1).max()))
eng.close()
eng = sympleints_b200.get_engine(0, 4, 5, boys_table=ref)

rng = np.random.default_rng(7)
def shell(L, K):
    ex = 10 ** rng.uniform(-1, 1, K); co = rng.uniform(0.2, 1, K)
    return ex, ints.norm_cgto_lmn(co, ex, L), rng.uniform(-1, 1, 3)
R = np.array([0.3, -0.2, 0.5])
worst = 0
for key in ["ovlp", "kin", "dpm", "qpm", "dqpm", "coul", "2c2e"]:
    f = ints.KEY_FUNCS[key][0]
    lm = 5 if key == "2c2e" else 4
    for La, Lb in [(0, 0), (1, 0), (1, 1), (2, 1), (3, 2), (lm, lm), (1, 3)]:
        sa, sb = shell(La, 3), shell(Lb, 2)
        for sph in (True, False):
            o = f(La, Lb, *sa, *sb, R=R, sph=sph)
            g = eng.eval_block(key, (La, Lb), *sa, *sb, R=R, sph=sph)
            err = np.abs(o - g).max() / np.abs(o).max()
            worst = max(worst, err)
            if err > 1e-12: print("FAIL", key, La, Lb, sph, err)
print("2-index blocks worst rel err %.2e" % worst)
worst = 0
for La, Lb, Lc in [(0,0,0),(0,0,1),(1,0,2),(1,1,1),(2,1,2),(2,2,2),(3,2,3),(4,4,5),(4,3,2),(1,3,2),(0,2,5),(4,0,5),(4,4,4)]:
    sa, sb, sc = shell(La, 2), shell(Lb, 2), shell(Lc, 2)
    o = ints.int3c2e_sph(La, Lb, Lc, *sa, *sb, *sc)
    g = eng.eval_block("3c2e", (La, Lb, Lc), *sa, *sb, *sc)
    err = np.abs(o - g).max() / np.abs(o).max()
    worst = max(worst, err)
    if err > 1e-12: print("FAIL 3c2e", La, Lb, Lc, err)
print("3c2e blocks worst rel err %.2e" % worst)

# small batched system
shells, sites = synthetic.orbital_basis(natoms=3, seed=5)
aux = synthetic.aux_basis(sites, seed=6)
ob, ab = eng.basis(shells), eng.basis(aux)
import torch
def ref_matrix(key, shells, sph=True, **kw):
    f, ncomp = ints.KEY_FUNCS[key]
    size = (lambda s: 2*s.L+1) if sph else (lambda s: (s.L+1)*(s.L+2)//2)
    off = np.concatenate([[0], np.cumsum([size(s) for s in shells])])
    M = np.zeros((ncomp, off[-1], off[-1]))
    for i, a in enumerate(shells):
        for j, b in enumerate(shells):
            blk = f(a.L, b.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center, sph=sph, **kw)
            M[:, off[i]:off[i+1], off[j]:off[j+1]] = blk.reshape(ncomp, size(a), size(b))
    return M
for key in ["ovlp", "kin", "dpm", "qpm"]:
    t0 = time.time(); g = eng.int1e(key, ob, R=R).cpu().numpy(); t1 = time.time()
    o = ref_matrix(key, shells, R=R)
    print(key, "matrix max abs err %.2e (max %.2e)  gpu %.3fs" % (np.abs(g.reshape(o.shape) - o).max(), np.abs(o).max(), t1 - t0))
g = eng.int2c2e(ab).cpu().numpy(); o = ref_matrix("2c2e", aux)[0]
print("2c2e matrix err %.2e (max %.2e)" % (np.abs(g - o).max(), np.abs(o).max()))
xyz, q = synthetic.point_charges(sites, n=16, seed=3)
g = eng.coul(ob, xyz, q).cpu().numpy()
o = sum(qq * ref_matrix("coul", shells, R=rr)[0] for rr, qq in zip(xyz, q))
print("coul matrix err %.2e (max %.2e)" % (np.abs(g - o).max(), np.abs(o).max()))
t0 = time.time(); T = eng.int3c2e(ob, ab); torch.cuda.synchronize(); t1 = time.time()
T = T.cpu().numpy()
Tp = eng.int3c2e(ob, ab, layout="packed").cpu().numpy()
print("3c2e tensor", T.shape, "time %.3fs" % (t1 - t0), "launches", eng.launch_count())
# reference on a sample of triples
offo = np.concatenate([[0], np.cumsum([2*s.L+1 for s in shells])]); offa = np.concatenate([[0], np.cumsum([2*s.L+1 for s in aux])])
worst = 0; r2 = np.random.default_rng(1)
for _ in range(150):
    i, j, k = r2.integers(len(shells)), r2.integers(len(shells)), r2.integers(len(aux))
    a, b, c = shells[i], shells[j], aux[k]
    o = ints.int3c2e_sph(a.L, b.L, c.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center, c.exps, c.coeffs, c.center)
    blk = T[offo[i]:offo[i+1], offo[j]:offo[j+1], offa[k]:offa[k+1]].ravel()
    worst = max(worst, np.abs(blk - o).max() / max(np.abs(o).max(), 1e-300))
print("3c2e tensor sampled worst rel err %.2e" % worst)
# packed vs full
row = 0; bad = 0
for i in range(len(shells)):
    for j in range(i + 1):
        ni, nj = offo[i+1]-offo[i], offo[j+1]-offo[j]
        blk = Tp[row:row + ni*nj].reshape(ni, nj, -1)
        bad = max(bad, np.abs(blk - T[offo[i]:offo[i+1], offo[j]:offo[j+1]]).max())
        row += ni * nj
print("packed vs full max diff %.2e" % bad, "symmetry %.2e" % np.abs(T - T.transpose(1, 0, 2)).max())
