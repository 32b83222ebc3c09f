import sys
from pathlib import Path
import torch
sys.path.insert(0, '/root/repo')
from sympleints_b200 import synthetic
from sympleints_b200.engine import Engine
eng = Engine()
shells, sites = synthetic.orbital_basis()
aux = synthetic.aux_basis(sites)
ab = eng.basis(aux)
buf = torch.empty((ab.nbf_sph, ab.nbf_sph), dtype=torch.float64, device=eng.tdev)
eng.int2c2e(ab, out=buf); torch.cuda.synchronize()
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.int2c2e(ab, out=buf); e1.record(); torch.cuda.synchronize()
    print("2c2e: %.3f ms" % e0.elapsed_time(e1))
print(float(buf.sum()))
