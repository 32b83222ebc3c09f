"""Time the 2c2e metric (5000 aux functions) and check it against the ungraphed path."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sympleints_b200
from sympleints_b200 import synthetic

eng = sympleints_b200.get_engine(0, 4, 5)
shells, sites = synthetic.orbital_basis(natoms=50)
aux = synthetic.aux_basis(sites)
ab = eng.basis(aux)
n = ab.nbf_sph
buf = torch.empty((n, n), dtype=torch.float64, device=eng.tdev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.tdev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    first = eng.int2c2e(ab, out=buf).clone()       # ordinary launches (fills the caches)
    second = eng.int2c2e(ab, out=buf).clone()      # captured + replayed
    third = eng.int2c2e(ab, out=buf).clone()       # replayed
    s.synchronize()
    print("graph == direct:", bool((first == second).all()), bool((first == third).all()))
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.int2c2e(ab, out=buf)
        e1.record()
        s.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("2c2e on a side stream (graph replay): median %.1f us  min %.1f us  -> %.0f GB/s" % (np.median(ts), np.min(ts), 8 * n * n / np.median(ts) / 1e3))
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.int2c2e(ab, out=buf)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("2c2e on the default stream (direct launches): median %.1f us" % np.median(ts))
eng.close()
