"""Per-step DRAM traffic and shared-memory pipe summary from an ncu launch list (csv) of one full-size bench step.

    python scripts/traffic_json.py gpurun_out/launches.csv "source text" > profiles/rNN_traffic.json
"""
import collections
import csv
import json
import sys

sys.path.insert(0, ".")
from sympleints_b200 import synthetic, workmodel

fn, source = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(fn)) if len(r) > 10]
h = rows[0]
ki, vi, mi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
per = collections.defaultdict(dict)
for r in rows[1:]:
    if r[ki].startswith("g3_"):
        per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
shells, sites = synthetic.orbital_basis(natoms=50)
work = workmodel.tensor_3c2e_work(shells, synthetic.aux_basis(sites))
S = lambda m: sum(v.get(m, 0.0) for v in per.values())
cyc = S("sm__cycles_elapsed.max")
out = {
    "source": source, "launches": len(per), "sum_kernel_time_ms": S("gpu__time_duration.sum") / 1e6,
    "dram_read_GB_per_step": S("dram__bytes_read.sum") / 1e9, "dram_write_GB_per_step": S("dram__bytes_write.sum") / 1e9,
    "algorithmic_GB_per_step": work["bytes"] / 1e9,
    "lsu_data_pipe_wavefronts_per_step": S("l1tex__data_pipe_lsu_wavefronts.sum"),
    "shared_wavefronts_per_step": S("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "sm_cycles_sum_of_max": cyc, "sms": 148,
}
if cyc:
    out["lsu_data_pipe_busy_frac"] = out["lsu_data_pipe_wavefronts_per_step"] / (cyc * 148)
    out["shared_pipe_busy_frac"] = out["shared_wavefronts_per_step"] / (cyc * 148)
out["traffic_over_algorithmic"] = (out["dram_read_GB_per_step"] + out["dram_write_GB_per_step"]) / out["algorithmic_GB_per_step"]
print(json.dumps(out, indent=1))
