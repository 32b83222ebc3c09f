"""Per-class micro-benchmark -- the CUDA counterpart of sympleints/benchmark.py.

The reference renders a Fortran driver, compiles it with gfortran and times ``niters`` calls of every
``(La, Lb[, Lc])`` class on random primitives (``nprims = 6``; benchmark.py:65-142,
templates/fortran_bench.tpl:49-106), printing ``La Lb [Lc] tot iter`` rows.  Here every class is timed on the
GPU on a PURE same-class batch: one shell of ``La`` after ``npairs`` shells of ``Lb``, restricted to the pairs
(La-shell, Lb-shell) with ``si_basis_select_rows`` (no diagonal pair), times ``naux`` auxiliary shells of ``Lc``
for the 3-centre key.  Reported per class: ``tot`` (seconds per batch, median of ``reps`` CUDA-event timings after
a warm-up), ``iter`` (seconds per shell tuple), integrals/s and -- for the Boys-type keys -- TFLOP/s in the
reference-graph flop model (workmodel.py).

    python -m sympleints_b200.benchmark --keys 3c2e_sph coul ovlp --lmax 4 --lauxmax 5 --nprims 6
"""
import argparse
import json

import numpy as np

from . import workmodel
from .lib import NCOMP
from .shells import Shell

ALL_KEYS = ("ovlp", "kin", "dpm", "qpm", "dqpm", "coul", "2c2e", "3c2e_sph")


def _rand_shells(rng, L, n, K, box=6.0):
    return [Shell(L, rng.uniform(-box, box, 3), rng.uniform(0.2, 1, K), np.sort(10 ** rng.uniform(-1, 1, K))[::-1].copy())
            for _ in range(n)]


def _time(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def run_key(engine, key, lmax, lauxmax, nprims=6, npairs=256, naux=64, ncharges=64, reps=5, seed=0, classes=None):
    """Rows ``{La, Lb, Lc, tot, iter, tuples, integrals_per_s, tflops}`` for every class of ``key``."""
    torch = engine.torch
    rng = np.random.default_rng(seed)
    three = key in ("3c2e", "3c2e_sph")
    lm = lauxmax if key == "2c2e" else lmax
    rows = []
    if classes is None:
        classes = [(La, Lb, Lc) for La in range(lm + 1) for Lb in range(La + 1)
                   for Lc in (range(lauxmax + 1) if three else [None])]
    for (La, Lb, Lc) in classes:
        ka = nprims
        orb = _rand_shells(rng, Lb, npairs, nprims) + _rand_shells(rng, La, 1, ka)
        ob = engine.basis(orb)
        ob.select_rows(npairs, npairs + 1, skip_diagonal=True)   # the npairs pairs (La-shell, Lb-shell) only
        nprim_tuple = ka * nprims
        na, nb = 2 * La + 1, 2 * Lb + 1
        if three:
            ab = engine.basis(_rand_shells(rng, Lc, naux, 1))
            out = torch.empty((ob.packed_rows, ab.nbf_sph), dtype=torch.float64, device=engine.tdev)
            fn = lambda: engine.int3c2e(ob, ab, layout="packed", out=out)
            fp, fc = workmodel.flops_3c2e(La, Lb, Lc)
            ntup, nint = npairs * naux, npairs * naux * na * nb * (2 * Lc + 1)
            flops = ntup * (nprim_tuple * fp + fc)
        elif key == "coul":
            xyz, q = rng.uniform(-6, 6, (ncharges, 3)), rng.uniform(-1, 1, ncharges)
            out = torch.zeros((ob.nbf_sph, ob.nbf_sph), dtype=torch.float64, device=engine.tdev)
            fn = lambda: engine.coul(ob, xyz, q, out=out)
            fp, fc = workmodel.flops_coul(La, Lb)
            ntup, nint = npairs, npairs * na * nb
            flops = ntup * (nprim_tuple * ncharges * fp + fc)
        else:
            nc = NCOMP[key]
            out = torch.zeros((nc, ob.nbf_sph, ob.nbf_sph), dtype=torch.float64, device=engine.tdev)
            fn = (lambda: engine.int2c2e(ob, out=out[0])) if key == "2c2e" else (lambda: engine.int1e(key, ob, out=out))
            ntup, nint = npairs, npairs * na * nb * nc
            flops = None
            if key == "2c2e":
                fp, fc = workmodel.flops_2c2e(La, Lb)
                flops = ntup * (nprim_tuple * fp + fc)
        t = _time(torch, fn, reps)
        rows.append({"La": La, "Lb": Lb, "Lc": Lc, "tot": t, "iter": t / ntup, "tuples": ntup, "integrals": nint,
                     "integrals_per_s": nint / t, "tflops": None if flops is None else flops / t / 1e12})
        ob.close()
        del out
    return rows


def format_rows(key, rows):
    lines = [f"# {key}: La Lb Lc        tot/s       iter/s   integrals/s  TFLOP/s(model)"]
    for r in rows:
        lc = "-" if r["Lc"] is None else r["Lc"]
        tf = "      -" if r["tflops"] is None else f"{r['tflops']:7.3f}"
        lines.append(f"{r['La']:2d} {r['Lb']:2d} {lc:>2} {r['tot']:12.6e} {r['iter']:12.6e} {r['integrals_per_s']:13.4e} {tf}")
    return "\n".join(lines)


def run(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", nargs="+", default=["3c2e_sph"], help="any of: " + " ".join(ALL_KEYS))
    ap.add_argument("--lmax", type=int, default=2)       # the reference's defaults (benchmark.py:208-227)
    ap.add_argument("--lauxmax", type=int, default=2)
    ap.add_argument("--nprims", type=int, default=6)
    ap.add_argument("--npairs", type=int, default=256)
    ap.add_argument("--naux", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None, help="write the rows as JSON")
    a = ap.parse_args(argv)
    import sympleints_b200

    eng = sympleints_b200.get_engine(0, max(a.lmax, 0), max(a.lauxmax, a.lmax))
    res = {}
    for key in a.keys:
        res[key] = run_key(eng, key, a.lmax, a.lauxmax, a.nprims, a.npairs, a.naux, reps=a.reps)
        print(format_rows(key, res[key]))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    eng.close()
    return res


if __name__ == "__main__":
    run()
