"""ctypes access to the TEST-ONLY host emulator (tests/emu/libsi_emu.so)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
KEYS = {"ovlp": 0, "kin": 1, "dpm": 2, "qpm": 3, "dqpm": 4, "coul": 5, "2c2e": 6, "3c2e": 7}
NCOMP = {"ovlp": 1, "kin": 1, "dpm": 3, "qpm": 6, "dqpm": 3, "coul": 1, "2c2e": 1, "3c2e": 1}
NORM = {"none": 0, "cgto": 1, "pgto": 2}
_dp = ctypes.POINTER(ctypes.c_double)


def build_emu(force=False):
    so = HERE / "emu" / "libsi_emu.so"
    srcs = [HERE / "emu" / "si_emu.cpp", REPO / "sympleints_b200" / "csrc" / "si_program.cpp"]
    hdrs = [REPO / "sympleints_b200" / "csrc" / "si_machine.h", REPO / "sympleints_b200" / "csrc" / "si_program.h"]
    newest = max(p.stat().st_mtime for p in srcs + hdrs)
    if force or not so.exists() or so.stat().st_mtime < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so)] + [str(s) for s in srcs])
    return so


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build_emu()))
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


def xlarge_pref(nmax):
    from math import pi, sqrt
    out = np.empty(nmax + 1)
    for n in range(nmax + 1):
        f2 = 1.0
        k = 2 * n - 1
        while k > 1:
            f2 *= k
            k -= 2
        out[n] = f2 / 2 ** (n + 1) * sqrt(pi)
    return out


def program_stats(key, La, Lb, Lc=0, sph=True, normalize="cgto"):
    out = (ctypes.c_longlong * 8)()
    rc = lib().siemu_program_stats(KEYS[key], La, Lb, Lc, int(sph), NORM[normalize], out)
    assert rc == 0
    names = ["w_size", "nstages", "n_fma_prim", "n_fma_contr", "lut_words", "nconst", "chunk", "x_size"]
    return dict(zip(names, list(out)))


def emu_eval(key, La, Lb, Lc, shell_a, shell_b, shell_c=None, R=None, q=None, sph=True, normalize="cgto", table=None):
    """shell_* = (exps, coeffs, center).  Handles La < Lb by swapping + transposing."""
    if La < Lb:
        res = emu_eval(key, Lb, La, Lc, shell_b, shell_a, shell_c, R, q, sph, normalize, table)
        size = (lambda l: 2 * l + 1) if (sph or key == "3c2e") else (lambda l: (l + 1) * (l + 2) // 2)
        na, nb = size(La), size(Lb)
        ncomp = NCOMP[key]
        nc = res.size // (ncomp * na * nb)
        return res.reshape(ncomp, nb, na, nc).transpose(0, 2, 1, 3).ravel().copy()
    if table is None:
        from oracle import boys as oboys
        table = oboys.table()
    table = np.ascontiguousarray(table)
    pref = xlarge_pref(table.shape[0] - 7)
    ax, da, A = [np.ascontiguousarray(x, dtype=float) for x in shell_a]
    bx, db, B = [np.ascontiguousarray(x, dtype=float) for x in shell_b]
    if shell_c is not None:
        cx, dc, C = [np.ascontiguousarray(x, dtype=float) for x in shell_c]
    else:
        cx, dc, C = np.ones(1), np.ones(1), np.zeros(3)
    R = np.zeros(3) if R is None else np.ascontiguousarray(R, dtype=float).reshape(-1)
    ncharge = R.size // 3
    q = np.ones(ncharge) if q is None else np.ascontiguousarray(q, dtype=float)
    size = (lambda l: 2 * l + 1) if (sph or key == "3c2e") else (lambda l: (l + 1) * (l + 2) // 2)
    n = NCOMP[key] * size(La) * size(Lb) * (2 * Lc + 1 if key == "3c2e" else 1)
    res = np.zeros(n)
    rc = lib().siemu_eval(
        KEYS[key], La, Lb, Lc, int(sph), NORM[normalize],
        ax.size, _p(ax), _p(da), _p(A), bx.size, _p(bx), _p(db), _p(B), cx.size, _p(cx), _p(dc), _p(C),
        _p(R), ncharge, _p(q), _p(table), table.shape[0], table.shape[1], _p(pref), _p(res))
    assert rc == 0
    return res


# ---- warp emulator of the GENERATED 3c2e kernels (tests/emu/g3_emu.cpp) ------------------------
# every (La >= Lb) pair with two aux momenta each + the extremes: 34 of the 90 classes
G3_EMU_CLASSES = sorted({(a, b, c) for a in range(5) for b in range(a + 1) for c in ((a + b) % 6, (2 * a + b + 3) % 6)}
                        | {(0, 0, 0), (0, 0, 5), (4, 4, 5), (4, 4, 0), (4, 0, 5), (2, 2, 1)})
_G3 = None


def g3_lib():
    global _G3
    if _G3 is not None:
        return _G3
    from sympleints_b200.codegen import gen3c2e

    inc = HERE / "emu" / "g3_emu_classes.inc"
    gen3c2e.write_emu_source(str(inc), G3_EMU_CLASSES)
    so = HERE / "emu" / "libg3_emu.so"
    deps = [inc, HERE / "emu" / "g3_emu.cpp", REPO / "sympleints_b200" / "csrc" / "si_g3.h",
            REPO / "sympleints_b200" / "csrc" / "si_machine.h"]
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-o", str(so),
                               str(HERE / "emu" / "g3_emu.cpp")])
    _G3 = ctypes.CDLL(str(so))
    return _G3


def g3_eval(La, Lb, Lc, shell_a, shell_b, aux_shells, table=None, tuple_table=False):
    """aux_shells: list of (exps, coeffs, center) all with angular momentum Lc.  Returns
    array (naux, na, nb, nc).  tuple_table=True: through an explicit tuple table; tuple_table=2: additionally a
    second pair (a shifted by (0.5, -0.25, 0.125), b) interleaved in the table; its blocks follow."""
    if table is None:
        from oracle import boys as oboys
        table = oboys.table()
    table = np.ascontiguousarray(table)
    pref = xlarge_pref(table.shape[0] - 7)
    ax, da, A = [np.ascontiguousarray(x, dtype=float) for x in shell_a]
    bx, db, B = [np.ascontiguousarray(x, dtype=float) for x in shell_b]
    Kc = np.array([len(s[0]) for s in aux_shells], dtype=np.int32)
    cx = np.ascontiguousarray(np.concatenate([s[0] for s in aux_shells]), dtype=float)
    dc = np.ascontiguousarray(np.concatenate([s[1] for s in aux_shells]), dtype=float)
    C = np.ascontiguousarray(np.array([s[2] for s in aux_shells], dtype=float).reshape(-1, 3))
    naux = len(aux_shells)
    res = np.zeros((naux * (2 if int(tuple_table) == 2 else 1), 2 * La + 1, 2 * Lb + 1, 2 * Lc + 1))
    rc = g3_lib().g3emu_eval(La, Lb, Lc, ax.size, _p(ax), _p(da), _p(A), bx.size, _p(bx), _p(db), _p(B), naux,
                             Kc.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _p(cx), _p(dc), _p(C), _p(table),
                             table.shape[0], table.shape[1], _p(pref), int(tuple_table), _p(res))
    assert rc == 0, rc
    return res


# ---- warp emulator of the GENERATED nuclear-attraction kernels (tests/emu/gc_emu.cpp) -----------
GC_EMU_CLASSES = [(a, b) for a in range(5) for b in range(a + 1)]
_GC = None


def gc_lib():
    global _GC
    if _GC is not None:
        return _GC
    from sympleints_b200.codegen import gencoul

    inc = HERE / "emu" / "gc_emu_classes.inc"
    gencoul.write_emu_source(str(inc), GC_EMU_CLASSES)
    so = HERE / "emu" / "libgc_emu.so"
    deps = [inc, HERE / "emu" / "gc_emu.cpp", REPO / "sympleints_b200" / "csrc" / "si_g3.h",
            REPO / "sympleints_b200" / "csrc" / "si_machine.h"]
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-o", str(so),
                               str(HERE / "emu" / "gc_emu.cpp")])
    _GC = ctypes.CDLL(str(so))
    return _GC


def gc_eval(La, Lb, pairs, cxyz, cq, table=None, nsplit=1):
    """pairs: list of (shell_a, shell_b), shell = (exps, coeffs, center), all of class (La >= Lb).
    Returns array (npairs, nsa, nsb) = sum_C q_C (a|1/r_C|b)."""
    if table is None:
        from oracle import boys as oboys
        table = oboys.table()
    table = np.ascontiguousarray(table)
    pref = xlarge_pref(table.shape[0] - 7)
    ip = ctypes.POINTER(ctypes.c_int)
    Ka = np.array([len(p[0][0]) for p in pairs], dtype=np.int32)
    Kb = np.array([len(p[1][0]) for p in pairs], dtype=np.int32)
    cat = lambda i, j: np.ascontiguousarray(np.concatenate([np.asarray(p[i][j], dtype=float).ravel() for p in pairs]))
    ax, da, A = cat(0, 0), cat(0, 1), cat(0, 2)
    bx, db, B = cat(1, 0), cat(1, 1), cat(1, 2)
    cxyz = np.ascontiguousarray(cxyz, dtype=float).reshape(-1, 3)
    cq = np.ascontiguousarray(cq, dtype=float)
    res = np.zeros((len(pairs), 2 * La + 1, 2 * Lb + 1))
    rc = gc_lib().gcemu_eval(La, Lb, len(pairs), Ka.ctypes.data_as(ip), _p(ax), _p(da), _p(A), Kb.ctypes.data_as(ip),
                             _p(bx), _p(db), _p(B), len(cq), _p(cxyz), _p(cq), _p(table), table.shape[0],
                             table.shape[1], _p(pref), nsplit, _p(res))
    assert rc == 0, rc
    return res
