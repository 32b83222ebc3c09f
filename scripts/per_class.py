"""Per-class table from an ncu launch list (csv) of a full-size bench step."""
import csv, collections, sys, json
sys.path.insert(0, '.')
from sympleints_b200 import synthetic, workmodel as wm
from sympleints_b200.codegen import gen3c2e as g
fn = sys.argv[1]; natoms = int(sys.argv[2]) if len(sys.argv) > 2 else 50
shells, sites = synthetic.orbital_basis(natoms=natoms); aux = synthetic.aux_basis(sites)
w = wm.tensor_3c2e_work(shells, aux)
rows = [r for r in csv.reader(open(fn)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); mi = hdr.index('Metric Name'); ii = hdr.index('ID')
data = collections.defaultdict(dict)
for r in rows[1:]:
    k = r[ki].split('(')[0]
    if k.startswith('g3_kernel_') or k.startswith('g3_tkernel_'):
        data[(r[ii], k[-3:])][r[mi]] = float(r[vi].replace(',', ''))
per = {}
for (id_, k), m in data.items():
    per[k] = m
out = []; tot = 0
F = 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'; Wm = 'sm__warps_active.avg.pct_of_peak_sustained_active'; I = 'smsp__issue_active.avg.pct_of_peak_sustained_active'
for k, m in per.items():
    cls = tuple(int(c) for c in k); t = m['gpu__time_duration.sum'] / 1e3; pc = w['per_class'][cls]; cg = g.ClassGen(*cls)
    out.append((t, cls, pc['flops'] / t / 1e6, 100 * pc['flops'] / w['flops'], m.get(F, 0), m.get(Wm, 0), m.get(I, 0), cg.EPT, cg.WSZ, pc['prim_triples'] / pc['triples'])); tot += t
out.sort(reverse=True)
print('classes', len(out), 'sum of kernel times (ms): %.1f' % (tot / 1e3), ' serialized model TF/s %.2f' % (w['flops'] / tot / 1e6))
print('   time_us  class     TF/s  %flops  fp64%  warps%  issue%  EPT WSZ  K')
for o in out[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print('%9.0f %s %6.2f  %5.2f  %5.1f  %5.1f  %5.1f  %2d %4d %4.1f' % o)
