"""TEST INFRASTRUCTURE: loader/evaluator for the REFERENCE's own generated numpy modules in
oracle/_ref (produced by oracle/refgen/generate_ref.py from /root/reference; git-ignored, travels
to the GPU box).  Used by tests/, __graft_entry__.smoke() and bench.py (parity sampling and the
cpu_baseline / reference arm) -- never by the product (sympleints_b200/ does not import oracle/).

Calling convention = the reference's generated functions (templates/py_func.tpl:1-16):
``f(ax, da, A, bx, db, B[, cx, dc, C], R, result)`` with broadcast-shaped exponent/coefficient
arrays; ``result`` is written in place.  The reference's 3-index "equivalent" wrappers for
La < Lb are broken, so swapped classes are evaluated as (Lb, La, Lc) and
transposed here (the intended ``resort_bac_abc`` semantics of graphs/templates/int_3c2e_mod.tpl).
"""
import importlib.util
from pathlib import Path

import numpy as np

REFDIR = Path(__file__).resolve().parent / "_ref"
MODULES = {  # key -> (module / dict name, ncomp)
    "ovlp": ("ovlp3d", 1), "kin": ("kinetic3d", 1), "dpm": ("dipole3d", 3), "qpm": ("quadrupole3d", 6),
    "dqpm": ("diag_quadrupole3d", 3), "coul": ("coulomb3d", 1), "2c2e": ("int2c2e3d", 1),
    "3c2e": ("int3c2e3d_sph", 1),
}
_CACHE = {}


def available(key, variant="sph_cgto_l4_laux5"):
    return (REFDIR / f"ref_{variant}_python" / f"{MODULES[key][0]}.py").exists()


def funcs(key, variant="sph_cgto_l4_laux5"):
    """dict {(La, Lb[, Lc]): callable} of the reference's generated module for ``key``."""
    k = (key, variant)
    if k not in _CACHE:
        modname = MODULES[key][0]
        path = REFDIR / f"ref_{variant}_python" / f"{modname}.py"
        if not path.exists():
            raise FileNotFoundError(f"{path}: run oracle/refgen/generate_ref.py in the build container")
        spec = importlib.util.spec_from_file_location(f"ref_{variant}_{modname}", path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _CACHE[k] = getattr(m, modname)
    return _CACHE[k]


def _size(l, sph):
    return 2 * l + 1 if sph else (l + 1) * (l + 2) // 2


def eval_pair(key, a, b, R=(0.0, 0.0, 0.0), dtype=np.float64, variant="sph_cgto_l4_laux5", sph=True):
    """Block (ncomp, na, nb) of shells a, b (objects with L, exps, coeffs, center)."""
    fd = funcs(key, variant)
    ncomp = MODULES[key][1]
    c = lambda x: np.asarray(x, dtype=dtype)
    res = np.zeros(ncomp * _size(a.L, sph) * _size(b.L, sph), dtype=dtype)
    fd[(a.L, b.L)](c(a.exps)[:, None], c(a.coeffs)[:, None], c(a.center),
                   c(b.exps)[None, :], c(b.coeffs)[None, :], c(b.center), c(R), res)
    return res.reshape(ncomp, _size(a.L, sph), _size(b.L, sph))


def eval_triple(a, b, cshell, dtype=np.float64, variant="sph_cgto_l4_laux5"):
    """Block (na, nb, nc) of the spherical 3c2e integrals (ab|c)."""
    fd = funcs("3c2e", variant)
    swap = a.L < b.L
    if swap:
        a, b = b, a
    c = lambda x: np.asarray(x, dtype=dtype)
    na, nb, nc = 2 * a.L + 1, 2 * b.L + 1, 2 * cshell.L + 1
    res = np.zeros(na * nb * nc, dtype=dtype)
    fd[(a.L, b.L, cshell.L)](c(a.exps)[:, None, None], c(a.coeffs)[:, None, None], c(a.center),
                             c(b.exps)[None, :, None], c(b.coeffs)[None, :, None], c(b.center),
                             c(cshell.exps)[None, None, :], c(cshell.coeffs)[None, None, :], c(cshell.center),
                             c(np.zeros(3)), res)
    res = res.reshape(na, nb, nc)
    return res.transpose(1, 0, 2) if swap else res


def block_check(val, ref64, ref80, atol=1e-14, rtol=1e-12, noise_factor=10.0):
    """The parity criterion of tests/golden_util.py (BASELINE.json: 1e-12 relative / 1e-14 absolute per
    block, or within 10x the reference's own float64 noise of its extended-precision evaluation).
    Returns (ok, ratio) with ratio = error / allowed (<= 1 passes)."""
    m = float(np.abs(ref64).max())
    tol = atol + rtol * m
    e64 = float(np.abs(val - ref64).max())
    if e64 <= tol:
        return True, e64 / tol
    e80 = float(np.abs(val.astype(np.longdouble) - ref80).max())
    noise = float(np.abs(ref64.astype(np.longdouble) - ref80).max())
    allowed = noise_factor * noise + tol
    return e80 <= allowed, e80 / allowed
