"""Per-class micro-benchmark -- the CUDA counterpart of sympleints/benchmark.py.

The reference renders a Fortran driver, compiles it with gfortran and times ``niters`` calls of every
``(La, Lb[, Lc])`` class on random primitives (``nprims = 6``, benchmark.py:65-142,
templates/fortran_bench.tpl:49-106), reporting ``La Lb [Lc] tot iter`` rows.  Here every class is timed
on the GPU over a batch of random same-class shell tuples (CUDA events, warm-up + median of ``reps``),
and reported with the reference-graph flop model (workmodel.py) as integrals/s and TFLOP/s.

    python -m sympleints_b200.benchmark --key 3c2e_sph --lmax 4 --lauxmax 5 --nprims 1 --ntuples 4096
"""
import argparse
import json

import numpy as np

from . import workmodel
from .shells import Shell


def _rand_shells(rng, L, n, K):
    return [Shell(L, rng.uniform(-4, 4, 3), rng.uniform(0.2, 1, K), np.sort(10 ** rng.uniform(-1, 1, K))[::-1].copy())
            for _ in range(n)]


def run_key(engine, key, lmax, lauxmax, nprims=6, ntuples=4096, reps=5, seed=0):
    import torch

    rng = np.random.default_rng(seed)
    rows = []
    three = key in ("3c2e", "3c2e_sph")
    lm = lauxmax if key == "2c2e" else lmax
    for La in range(lm + 1):
        for Lb in range(La + 1):
            for Lc in (range(lauxmax + 1) if three else [None]):
                # a small orbital set whose pairs are dominated by the (La, Lb) class + n aux shells of Lc
                npair = max(2, int(np.sqrt(ntuples)) if not three else 16)
                orb = _rand_shells(rng, La, npair, nprims) + (_rand_shells(rng, Lb, npair, nprims) if Lb != La else [])
                ob = engine.basis(orb)
                if three:
                    aux = _rand_shells(rng, Lc, max(1, ntuples // (npair * npair)), 1)
                    ab = engine.basis(aux)
                    fn = lambda: engine.int3c2e(ob, ab, layout="packed")
                    fp, fc = workmodel.flops_3c2e(La, Lb, Lc)
                elif key == "coul":
                    xyz, q = rng.uniform(-4, 4, (64, 3)), rng.uniform(-1, 1, 64)
                    fn = lambda: engine.coul(ob, xyz, q)
                    fp, fc = workmodel.flops_coul(La, Lb)
                else:
                    fn = lambda: engine.int1e(key, ob)
                    fp, fc = (workmodel.flops_2c2e(La, Lb) if key == "2c2e" else (0, 0))
                fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e-3)
                t = float(np.median(ts))
                rows.append({"La": La, "Lb": Lb, "Lc": Lc, "tot": t, "integrals": int(out.numel()),
                             "integrals_per_s": out.numel() / t, "note": "whole mixed-class batch"})
                ob.close()
    return rows


def run(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--key", default="3c2e_sph")
    ap.add_argument("--lmax", type=int, default=4)
    ap.add_argument("--lauxmax", type=int, default=5)
    ap.add_argument("--nprims", type=int, default=6)
    ap.add_argument("--ntuples", type=int, default=4096)
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    import sympleints_b200

    eng = sympleints_b200.get_engine(0, a.lmax, a.lauxmax)
    rows = run_key(eng, a.key, a.lmax, a.lauxmax, a.nprims, a.ntuples)
    print("La Lb Lc        tot/s   integrals/s")
    for r in rows:
        print(f"{r['La']:2d} {r['Lb']:2d} {'-' if r['Lc'] is None else r['Lc']:>2} {r['tot']:12.6f} {r['integrals_per_s']:13.4e}")
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)
    eng.close()


if __name__ == "__main__":
    run()
