"""Time the nuclear-attraction matrix of the benchmark system (1024 point charges)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sympleints_b200 import synthetic
from sympleints_b200.engine import Engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = Engine()
shells, sites = synthetic.orbital_basis()
ob = eng.basis(shells)
xyz, q = synthetic.point_charges(sites, n=1024)
buf = torch.empty((ob.nbf_sph, ob.nbf_sph), dtype=torch.float64, device=eng.tdev)
eng.coul(ob, xyz, q, out=buf)
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.coul(ob, xyz, q, out=buf)
    e1.record()
    torch.cuda.synchronize()
    print("coul 1024 charges: %.3f ms" % e0.elapsed_time(e1), flush=True)
print("checksum", float(buf.sum()), float(buf.abs().max()))
