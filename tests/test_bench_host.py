"""CPU tests of bench.py's host side: the synthetic system in both auxiliary orders, the config block and the
reference arm (`--impl reference`: the reference's generated numpy code on the host cores, same JSON contract)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def test_class_major_system_is_a_permutation_of_the_generated_one():
    import bench
    from sympleints_b200 import workmodel

    shells, sites, aux_c = bench.build_system(50, "class")
    shells2, _, aux_a = bench.build_system(50, "atom")
    assert len(shells) == len(shells2) == 300 and len(aux_c) == len(aux_a) == 1200
    Ls = [s.L for s in aux_c]
    assert Ls == sorted(Ls) and [s.L for s in aux_a] != Ls
    key = lambda s: (s.L, tuple(np.round(s.center, 12)), tuple(s.exps), tuple(s.coeffs))
    assert sorted(map(key, aux_c)) == sorted(map(key, aux_a))
    # same work either way
    wc, wa = workmodel.tensor_3c2e_work(shells, aux_c), workmodel.tensor_3c2e_work(shells, aux_a)
    assert wc["integrals"] == wa["integrals"] == 4245750000 and wc["flops"] == wa["flops"]
    cfg = bench.config_dict(shells, aux_c, wc, 1)
    assert "class-major" in cfg["aux_order"] and cfg["integrals_per_step"] == 4245750000
    assert "model" not in cfg and "workload" in cfg


def test_reference_arm_prints_the_contract_line():
    from oracle import refmod

    if not refmod.available("3c2e"):
        pytest.skip("oracle/_ref was not generated")
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-per-class", "1"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "integrals/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "class-major" in d["config"]["aux_order"] and "workload" in d["config"]
