"""Per-class cost of the fused one-electron kernel: time it on sub-bases holding one or two shell types."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sympleints_b200
from sympleints_b200 import synthetic

eng = sympleints_b200.get_engine(0, 4, 5)
shells, sites = synthetic.orbital_basis(natoms=50)
types = {"s6": [s for s in shells if s.L == 0 and len(s.exps) == 6], "s1": [s for s in shells if s.L == 0 and len(s.exps) == 1],
         "p3": [s for s in shells if s.L == 1], "d2": [s for s in shells if s.L == 2], "f1": [s for s in shells if s.L == 3],
         "g1": [s for s in shells if s.L == 4]}


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


for n in (1, 2, 8, 50):
    sub = types["s1"][:n]
    b = eng.basis(sub)
    buf = torch.empty((11, n, n), dtype=torch.float64, device=eng.tdev)
    print("s1 x", n, "ovlp %.1f us  set %.1f us" % (timeit(lambda: eng.int1e("ovlp", b, out=buf[:1])), timeit(lambda: eng.int1e_set(b, out=buf))))
for names in (["s6"], ["s1"], ["p3"], ["d2"], ["f1"], ["g1"], ["s6", "g1"], ["p3", "d2"], ["f1", "g1"], ["s6", "s1", "p3", "d2", "f1", "g1"]):
    sub = [s for n in names for s in types[n]]
    b = eng.basis(sub)
    nbf = b.nbf_sph
    buf = torch.empty((11, nbf, nbf), dtype=torch.float64, device=eng.tdev)
    npairs = len(sub) * (len(sub) + 1) // 2
    t1 = timeit(lambda: eng.int1e("ovlp", b, out=buf[:1]))
    t11 = timeit(lambda: eng.int1e_set(b, out=buf))
    print("%-22s shells %4d pairs %6d nbf %5d  ovlp %7.1f us  set %7.1f us" % ("+".join(names), len(sub), npairs, nbf, t1, t11))
eng.close()
