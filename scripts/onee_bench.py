"""Time the 1e matrices of the benchmark system."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sympleints_b200 import synthetic
from sympleints_b200.engine import Engine

keys = sys.argv[1].split(",") if len(sys.argv) > 1 else ["ovlp", "kin", "dpm", "qpm"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = Engine()
shells, sites = synthetic.orbital_basis()
ob = eng.basis(shells)
for key in keys:
    ncomp = {"ovlp": 1, "kin": 1, "dpm": 3, "qpm": 6, "dqpm": 3}[key]
    buf = torch.empty((ncomp, ob.nbf_sph, ob.nbf_sph), dtype=torch.float64, device=eng.tdev)
    eng.int1e(key, ob, out=buf)
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.int1e(key, ob, out=buf)
        e1.record()
        torch.cuda.synchronize()
        print("%s: %.3f ms" % (key, e0.elapsed_time(e1)), flush=True)
    print(key, "checksum", float(buf.sum()), float(buf.abs().max()))
