"""Time the fused one-electron kernel (si_int1e_set) and the per-key calls on the benchmark basis."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sympleints_b200
from sympleints_b200 import synthetic

eng = sympleints_b200.get_engine(0, 4, 5)
shells, sites = synthetic.orbital_basis(natoms=50)
ob = eng.basis(shells)
nbf = ob.nbf_sph
buf = torch.empty((11, nbf, nbf), dtype=torch.float64, device=eng.tdev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.tdev)


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


print("set (11 comps): median %.1f us, min %.1f us  -> %.0f GB/s" % (*timeit(lambda: eng.int1e_set(ob, out=buf)), 8 * 11 * nbf * nbf / timeit(lambda: eng.int1e_set(ob, out=buf))[0] / 1e3))
for key, n in (("ovlp", 1), ("kin", 1), ("dpm", 3), ("qpm", 6)):
    print(key, "median %.1f us, min %.1f us" % timeit(lambda: eng.int1e(key, ob, out=buf[:n])))
eng.close()
