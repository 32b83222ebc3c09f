#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE's own code.

Runs in the build container only (needs /root/reference and oracle/_ref, see
oracle/refgen/generate_ref.py).  Two kinds of fixtures:

* boys_golden.npz  -- table and sample values of the reference's
  ``sympleints.testing.boys.get_boys_func(nmax=14)`` (imported from /root/reference).
* blocks_<variant>.npz -- seeded random shell pairs/triples evaluated with the
  reference's generated numpy modules (float64 AND numpy.longdouble inputs; the
  latter measures the reference's own FP64 rounding noise).

    python tests/golden/make_golden.py [--only boys|blocks] [--variants ...]
"""
import argparse
import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
REFDIR = REPO / "oracle" / "_ref"

MODULES = {  # key -> (module file, dict name, ncomp)
    "ovlp": ("ovlp3d", 1), "kin": ("kinetic3d", 1), "dpm": ("dipole3d", 3), "qpm": ("quadrupole3d", 6),
    "dqpm": ("diag_quadrupole3d", 3), "coul": ("coulomb3d", 1), "2c2e": ("int2c2e3d", 1),
    "3c2e": ("int3c2e3d_sph", 1),
    "gto": ("sph_gto3d", 1), "prefactor": ("prefactors", 1), "multi_sph": ("multipole3d_sph", 9),
}


def load(path):
    spec = importlib.util.spec_from_file_location(path.stem + "_" + path.parent.name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def norm_cgto(coeffs, exps, L):
    from oracle.ints import norm_cgto_lmn
    return norm_cgto_lmn(coeffs, exps, L)


def make_boys():
    sys.path.insert(0, str(REPO / "oracle" / "refgen" / "stubs"))
    sys.path.insert(0, "/root/reference")
    from sympleints.testing.boys import get_boys_func
    boys = get_boys_func(nmax=14)
    table = [c.cell_contents for c in boys.__closure__ if isinstance(c.cell_contents, np.ndarray)]
    # the scalar evaluator closes over table_full; dig it out of the nested closures
    def find_table(fn, depth=0):
        for c in (fn.__closure__ or ()):
            v = c.cell_contents
            if isinstance(v, np.ndarray) and v.ndim == 2 and v.shape[1] == 301:
                return v
            if callable(v) and depth < 3:
                t = find_table(v, depth + 1)
                if t is not None:
                    return t
        return None
    table = find_table(boys)
    assert table is not None and table.shape == (21, 301)
    rng = np.random.default_rng(14)
    xs = np.concatenate([[0.0, 1e-12, 0.1, 0.3, 29.9999, 30.0, 30.0001, 35.0, 200.0],
                         rng.uniform(0, 30, 600), rng.uniform(30, 120, 100),
                         np.arange(1, 300) * 0.1])
    vals = np.array([[boys(n, float(x)) for x in xs] for n in range(15)])
    np.savez_compressed(HERE / "boys_golden.npz", table=table, xs=xs, vals=vals)
    print("boys_golden.npz", table.shape, vals.shape)


def rand_shell(rng, L, K, normalize):
    exps = np.sort(10 ** rng.uniform(-1, 1.5, K))[::-1].copy()
    coeffs = rng.uniform(0.2, 1.0, K)
    if normalize == "cgto":
        coeffs = norm_cgto(coeffs, exps, L)
    center = rng.uniform(-1.5, 1.5, 3)
    return exps, coeffs, center


def make_blocks(variant, keys, lmax, lauxmax, sph, normalize, seed, classes3=None, refdir=None):
    d = REFDIR / f"ref_{refdir or variant}_python"
    if not d.exists():
        print("skip", variant, "(not generated)")
        return
    rng = np.random.default_rng(seed)
    out = {}
    meta = []
    n = 0
    for key in keys:
        modname, ncomp = MODULES[key]
        f = d / f"{modname}.py"
        if not f.exists():
            print("  skip", key, "(module missing)")
            continue
        mod = load(f)
        fd = getattr(mod, modname)
        three = key == "3c2e"
        lm = lauxmax if key == "2c2e" else lmax
        if key == "gto":
            # one-index functions f(ax, da, A, R, result): the values of a shell's functions at the point R
            for L in range(lmax + 1):
                for K in (3, 1):
                    sa = rand_shell(rng, L, K, normalize)
                    R = sa[2] + rng.uniform(-1.2, 1.2, 3)
                    res = {}
                    for tag, dt in (("f64", np.float64), ("f80", np.longdouble)):
                        c = lambda x: np.asarray(x, dtype=dt)
                        result = np.zeros(2 * L + 1 if sph else (L + 1) * (L + 2) // 2, dtype=dt)
                        fd[(L,)](c(sa[0]), c(sa[1]), c(sa[2]), c(R), result)
                        res[tag] = result
                    tagname = f"gto_{L}" if K == 3 else f"gto_{L}0"   # (second digit only makes the tag unique)
                    out[tagname + "_a"] = np.concatenate([sa[0], sa[1], sa[2]])
                    out[tagname + "_R"] = R
                    out[tagname + "_ref64"] = res["f64"]
                    hi = res["f80"].astype(np.float64)
                    out[tagname + "_ref80hi"] = hi
                    out[tagname + "_ref80lo"] = (res["f80"] - hi.astype(np.longdouble)).astype(np.float64)
                    meta.append(tagname)
                    n += 1
            continue
        if key == "prefactor":
            # the generated functions read the module-level names rP / rP2 (they are not arguments)
            rP = rng.uniform(-1.5, 1.5, 3)
        if three:
            cls = classes3 or [(a, b, c) for a in range(lmax + 1) for b in range(a + 1) for c in range(lauxmax + 1)]
        else:
            cls = [(a, b) for a in range(lm + 1) for b in range(a + 1)]
            # a few La < Lb classes: exercises the reference's "equivalent" wrappers (2-index only;
            # the reference's 3-index wrappers raise TypeError)
            cls += [(a, b) for (a, b) in [(0, 1), (1, 2), (1, 3), (2, 4), (0, lm)] if b <= lm]
        for Ls in cls:
            if Ls not in fd:
                continue
            Ks = [(3, 2, 2), (1, 1, 1)][n % 2] if key not in ("prefactor", "multi_sph") else (1, 1, 1)   # primitive functions
            sa = rand_shell(rng, Ls[0], Ks[0], normalize)
            sb = rand_shell(rng, Ls[1], Ks[1], normalize)
            sc = rand_shell(rng, Ls[2], Ks[2], normalize) if three else None
            R = rng.uniform(-1, 1, 3)
            size = (lambda l: 2 * l + 1) if sph else (lambda l: (l + 1) * (l + 2) // 2)
            nres = ncomp * int(np.prod([size(l) for l in Ls]))
            res = {}
            for tag, dt in (("f64", np.float64), ("f80", np.longdouble)):
                c = lambda x: np.asarray(x, dtype=dt)
                result = np.zeros(nres, dtype=dt)
                if key == "multi_sph":   # primitive functions: scalar exponents / coefficients, plain assignments
                    result = np.zeros(nres, dtype=dt)
                    fd[Ls](c(sa[0])[0], c(sa[1])[0], c(sa[2]), c(sb[0])[0], c(sb[1])[0], c(sb[2]), c(R), result)
                elif key == "prefactor":
                    mod.rP, mod.rP2 = c(rP), c(rP) ** 2
                    result = np.zeros(nres, dtype=dt)
                    fd[Ls](c(sa[0])[0], c(sa[1])[0], c(sa[2]), c(sb[0])[0], c(sb[1])[0], c(sb[2]), result)
                    R = rP
                elif three:
                    fd[Ls](c(sa[0])[:, None, None], c(sa[1])[:, None, None], c(sa[2]),
                           c(sb[0])[None, :, None], c(sb[1])[None, :, None], c(sb[2]),
                           c(sc[0])[None, None, :], c(sc[1])[None, None, :], c(sc[2]), c(R), result)
                else:
                    fd[Ls](c(sa[0])[:, None], c(sa[1])[:, None], c(sa[2]),
                           c(sb[0])[None, :], c(sb[1])[None, :], c(sb[2]), c(R), result)
                res[tag] = result
            tagname = f"{key}_{''.join(map(str, Ls))}"
            out[tagname + "_a"] = np.concatenate([sa[0], sa[1], sa[2]])
            out[tagname + "_b"] = np.concatenate([sb[0], sb[1], sb[2]])
            if three:
                out[tagname + "_c"] = np.concatenate([sc[0], sc[1], sc[2]])
            out[tagname + "_R"] = R
            out[tagname + "_ref64"] = res["f64"]
            # store the longdouble result as float64 + residual (exactly representable pair)
            hi = res["f80"].astype(np.float64)
            out[tagname + "_ref80hi"] = hi
            out[tagname + "_ref80lo"] = (res["f80"] - hi.astype(np.longdouble)).astype(np.float64)
            meta.append(tagname)
            n += 1
    out["_meta"] = np.array(meta)
    out["_sph"] = np.array(int(sph))
    out["_normalize"] = np.array(normalize)
    np.savez_compressed(HERE / f"blocks_{variant}.npz", **out)
    print(f"blocks_{variant}.npz: {n} blocks")


VARIANTS = {
    "sph_cgto_l4_laux5": dict(keys=["ovlp", "kin", "dpm", "qpm", "dqpm", "coul", "2c2e", "3c2e"], lmax=4, lauxmax=5, sph=True, normalize="cgto", seed=101),
    "sph_cgto_l3_laux3": dict(keys=["3c2e"], lmax=3, lauxmax=3, sph=True, normalize="cgto", seed=102),
    "cart_cgto_l2_laux2": dict(keys=["ovlp", "kin", "dpm", "qpm", "dqpm", "coul", "2c2e"], lmax=2, lauxmax=2, sph=False, normalize="cgto", seed=103),
    "sph_pgto_l2_laux2": dict(keys=["ovlp", "kin", "dpm", "qpm", "coul", "2c2e", "3c2e"], lmax=2, lauxmax=2, sph=True, normalize="pgto", seed=104),
    "sph_none_l2_laux2": dict(keys=["ovlp", "kin", "dpm", "coul", "2c2e", "3c2e"], lmax=2, lauxmax=2, sph=True, normalize="none", seed=105),
    "sph_cgto_l1_laux1": dict(keys=["ovlp", "2c2e", "3c2e"], lmax=1, lauxmax=1, sph=True, normalize="cgto", seed=106),
    # section 8(f)-1 keys; generated into the same reference directory as the lmax 4 set
    "sph_cgto_l4_gto_prefactor": dict(keys=["gto", "prefactor"], lmax=4, lauxmax=5, sph=True, normalize="cgto", seed=107,
                                      refdir="sph_cgto_l4_laux5"),
    "sph_cgto_l2_multi_sph": dict(keys=["multi_sph"], lmax=2, lauxmax=2, sph=True, normalize="cgto", seed=108,
                                  refdir="sph_cgto_l2_laux2"),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--variants", nargs="*", default=None)
    a = ap.parse_args()
    if a.only in (None, "boys"):
        make_boys()
    if a.only in (None, "blocks"):
        for v, kw in VARIANTS.items():
            if a.variants and v not in a.variants:
                continue
            make_blocks(v, **kw)
