"""Oracle integrals -- plain numpy restatement of the sympleints recurrences.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the
reference's own generated numpy code (oracle/_ref, produced by
oracle/refgen/generate_ref.py) and against tests/golden/*.npz.

Every function evaluates ONE shell pair / triple and returns the contracted
block in the reference's flat layout ``(comp, a, b[, c])`` (C order; the
component axis is dropped when ncomp == 1, PythonRenderer.py:62-68).
Primitives are handled by numpy broadcasting exactly like the generated code:
``ax, da`` have shape (Ka,1[,1]), ``bx, db`` (1,Kb[,1]), ``cx, dc`` (1,1,Kc).

Restated from (file:line in /root/reference/sympleints):
  defs/twocenter1d.py:14-51      p, mu, P, PA, PB, PR, K
  defs/multipole.py:17-47        Multipole1d recurrence, component order
  defs/kinetic.py:15-50          Kinetic1d recurrence
  defs/coulomb.py:31-101         nuclear attraction OS recurrence
  defs/coulomb.py:104-184        2c2e recurrence
  defs/coulomb.py:187-377        3c2e (Ahlrichs truncated aux VRR)
  main.py:122-144,152-189,241-280  coefficient product, normalisation, C2S
  cart2sph.py:28-88,163-202      Schlegel-Frisch coefficients, real combination
  helpers.py:42-49               canonical Cartesian order
"""
import functools
import itertools as it
from math import factorial, comb, floor, pi, sqrt

import numpy as np

from oracle.boys import boys_vec as _boys_default


# --------------------------------------------------------------------------
# orderings, normalisation, cart2sph
# --------------------------------------------------------------------------
def canonical_order(L):
    """helpers.py:42-49."""
    out = []
    for i in range(L + 1):
        l = L - i
        for n in range(i + 1):
            out.append((l, i - n, n))
    return out


def ncart(L):
    return (L + 1) * (L + 2) // 2


def nsph(L):
    return 2 * L + 1


def fact2(n):
    """Double factorial with (-1)!! = 1 (testing/normalization.py:9-18)."""
    if n <= 0:
        return 1
    r = 1
    while n > 1:
        r *= n
        n -= 2
    return r


@functools.lru_cache(None)
def lmn_factors(L):
    """main.py:175-183 / testing/normalization.py:20-26."""
    return np.array(
        [1.0 / sqrt(fact2(2 * l - 1) * fact2(2 * m - 1) * fact2(2 * n - 1))
         for (l, m, n) in canonical_order(L)]
    )


def norm_pgto(lmn, exponent):
    """main.py:152-159."""
    L = sum(lmn)
    f = fact2(2 * lmn[0] - 1) * fact2(2 * lmn[1] - 1) * fact2(2 * lmn[2] - 1)
    return np.sqrt(2 ** (2 * L + 1.5) * exponent ** (L + 1.5) / f / pi**1.5)


def norm_cgto_lmn(coeffs, exps, L):
    """testing/normalization.py:29-43 (radial part folded into the coefficients)."""
    coeffs = np.asarray(coeffs, dtype=float)
    exps = np.asarray(exps, dtype=float)
    N = 0.0
    for i, ei in enumerate(exps):
        for j, ej in enumerate(exps):
            tmp = coeffs[i] * coeffs[j] / (ei + ej) ** (L + 1.5)
            tmp *= np.sqrt(ei * ej) ** (L + 1.5)
            N += tmp
    N = np.sqrt(exps ** (L + 1.5) / (np.pi**1.5 / 2**L * N))
    return N * coeffs


def _c2s_complex(lx, ly, lz, m):
    """Schlegel/Frisch Eq. (15), cart2sph.py:28-88, in exact integer arithmetic
    up to the final square root."""
    l = lx + ly + lz
    am = abs(m)
    j2 = lx + ly - am
    if j2 % 2 != 0 or j2 < 0:
        return 0.0
    j = j2 // 2
    lmm = l - am
    sign = 1 if m >= 0 else -1
    tot = 0
    for i in range(lmm // 2 + 1):
        if j > i:
            continue
        i_fact_num = comb(l, i) * comb(i, j) * (-1) ** i * factorial(2 * l - 2 * i)
        i_fact_den = factorial(lmm - 2 * i)
        ksum = 0
        for k in range(j + 1):
            if lx - 2 * k < 0 or lx - 2 * k > am:
                continue
            e2 = sign * (am - lx + 2 * k)  # exponent of (-1)^(e2/2) == i^e2
            ksum += (1j ** (e2 % 4)) * comb(j, k) * comb(am, lx - 2 * k)
        from fractions import Fraction
        tot += complex(ksum) * float(Fraction(i_fact_num, i_fact_den))
    from fractions import Fraction
    rad = Fraction(
        factorial(2 * lx) * factorial(2 * ly) * factorial(2 * lz) * factorial(l) * factorial(lmm),
        factorial(2 * l) * factorial(lx) * factorial(ly) * factorial(lz) * factorial(l + am),
    )
    return tot * sqrt(float(rad)) / (2**l * factorial(l))


@functools.lru_cache(None)
def cart2sph(l, zero_thresh=1e-14):
    """Real C2S matrix, shape (2l+1, ncart(l)); cart2sph.py:163-202.

    Row index is m = -l..l; for m != 0 the real combination is
    (c_m + sign*c_-m)/sqrt(sign*2) with sign = +1 for m < 0 and -1 for m > 0.
    """
    C = np.zeros((ncart(l), 2 * l + 1), dtype=complex)
    for ci, (lx, ly, lz) in enumerate(canonical_order(l)):
        for mi, m in enumerate(range(-l, l + 1)):
            c = _c2s_complex(lx, ly, lz, m)
            if m != 0:
                cm = _c2s_complex(lx, ly, lz, -m)
                sign = 1 if m < 0 else -1
                c = (c + sign * cm) / np.sqrt(complex(sign * 2))
            C[ci, mi] = c
    assert np.abs(C.imag).max() < 1e-13
    C = C.real.copy()
    C[np.abs(C) <= zero_thresh] = 0.0
    return np.ascontiguousarray(C.T)


# --------------------------------------------------------------------------
# geometry helper
# --------------------------------------------------------------------------
class _Pair:
    """defs/twocenter1d.py:14-51 for broadcast primitive arrays."""

    def __init__(self, ax, A, bx, B, R=None):
        self.ax, self.bx = ax, bx
        self.A = [A[i] for i in range(3)]
        self.B = [B[i] for i in range(3)]
        self.p = ax + bx
        self.mu = ax * bx / self.p
        self.AB = [self.A[i] - self.B[i] for i in range(3)]
        self.P = [(ax * self.A[i] + bx * self.B[i]) / self.p for i in range(3)]
        self.PA = [self.P[i] - self.A[i] for i in range(3)]
        self.PB = [self.P[i] - self.B[i] for i in range(3)]
        if R is not None:
            self.PR = [self.P[i] - R[i] for i in range(3)]
        self.K1 = [np.exp(-self.mu * self.AB[i] ** 2) for i in range(3)]


def _argmax_first(args):
    m = max(args)
    return args.index(m)


# --------------------------------------------------------------------------
# 1-electron multipole / kinetic (1-D factorised)
# --------------------------------------------------------------------------
class _Multipole1d:
    """defs/multipole.py:10-32 with Strategy.HIGHEST (defs/strategy.py:31-33)."""

    def __init__(self, pr, d):
        self.p, self.PA, self.PB = pr.p, pr.PA[d], pr.PB[d]
        self.PR = pr.PR[d] if hasattr(pr, "PR") else None
        self.K = pr.K1[d]
        self.cache = {}

    def __call__(self, i, j, e):
        if min(i, j, e) < 0:
            return 0.0
        key = (i, j, e)
        if key in self.cache:
            return self.cache[key]
        if i + j + e == 0:
            val = np.sqrt(pi / self.p) * self.K
        else:
            args = [i, j, e]
            ind = _argmax_first(args)
            args[ind] -= 1
            i_, j_, e_ = args
            X = (self.PA, self.PB, self.PR)[ind]
            val = X * self(i_, j_, e_) + 1 / (2 * self.p) * (
                i_ * self(i_ - 1, j_, e_) + j_ * self(i_, j_ - 1, e_) + e_ * self(i_, j_, e_ - 1)
            )
        self.cache[key] = val
        return val


class _Kinetic1d:
    """defs/kinetic.py:15-44."""

    def __init__(self, pr, d):
        self.pr, self.d = pr, d
        self.S = _Multipole1d(pr, d)
        self.p, self.PA, self.PB = pr.p, pr.PA[d], pr.PB[d]
        self.ax, self.bx = pr.ax, pr.bx
        self.cache = {}

    def ovlp(self, i, j):
        return self.S(i, j, 0)

    def __call__(self, i, j):
        if min(i, j) < 0:
            return 0.0
        key = (i, j)
        if key in self.cache:
            return self.cache[key]
        if i + j == 0:
            val = (self.ax - 2 * self.ax**2 * (self.PA**2 + 1 / (2 * self.p))) * self.ovlp(0, 0)
        else:
            args = [i, j]
            ind = _argmax_first(args)
            args[ind] -= 1
            i_, j_ = args
            X = (self.PA, self.PB)[ind]
            rr = X * self(i_, j_) + 1 / (2 * self.p) * (i_ * self(i_ - 1, j_) + j_ * self(i_, j_ - 1))
            if ind == 0:
                val = rr + self.bx / self.p * (2 * self.ax * self.ovlp(i_ + 1, j_) - i_ * self.ovlp(i_ - 1, j_))
            else:
                val = rr + self.ax / self.p * (2 * self.bx * self.ovlp(i_, j_ + 1) - j_ * self.ovlp(i_, j_ - 1))
        self.cache[key] = val
        return val


# --------------------------------------------------------------------------
# Boys-type recurrences (memoised numeric recursion mirroring the reference)
# --------------------------------------------------------------------------
def _e(i):
    v = [0, 0, 0]
    v[i] = 1
    return tuple(v)


def _sub(a, b):
    return tuple(x - y for x, y in zip(a, b))


def _add(a, b):
    return tuple(x + y for x, y in zip(a, b))


class _Coulomb:
    """defs/coulomb.py:31-92."""

    def __init__(self, pr, boys):
        self.pr = pr
        R2 = sum(x * x for x in pr.PR)
        AB2 = sum(x * x for x in pr.AB)
        self.K = np.exp(-pr.mu * AB2)
        self.x = pr.p * R2
        self.boys = boys
        self.cache = {}

    def base(self, N):
        x64 = np.asarray(self.x, dtype=np.float64)
        return 2 * pi / self.pr.p * self.K * self.boys(int(N), x64)

    def __call__(self, a, b, N):
        if min(a) < 0 or min(b) < 0:
            return 0.0
        key = (a, b, N)
        if key in self.cache:
            return self.cache[key]
        pr = self.pr
        if sum(a) + sum(b) == 0:
            val = self.base(N)
        else:
            # order x(bra), x(ket), y(bra), y(ket), z(bra), z(ket)  (coulomb.py:81-92)
            for d in range(3):
                if a[d] > 0:
                    which, dd = "bra", d
                    break
                if b[d] > 0:
                    which, dd = "ket", d
                    break
            one = _e(dd)
            if which == "bra":
                X = pr.PA[dd]
                bra, ket = _sub(a, one), b
            else:
                X = pr.PB[dd]
                bra, ket = a, _sub(b, one)
            bra_d, ket_d = _sub(bra, one), _sub(ket, one)
            val = (
                X * self(bra, ket, N)
                - pr.PR[dd] * self(bra, ket, N + 1)
                + 1 / (2 * pr.p) * bra[dd] * (self(bra_d, ket, N) - self(bra_d, ket, N + 1))
                + 1 / (2 * pr.p) * ket[dd] * (self(bra, ket_d, N) - self(bra, ket_d, N + 1))
            )
        self.cache[key] = val
        return val


class _TwoC2E:
    """defs/coulomb.py:104-175."""

    def __init__(self, pr, boys):
        self.pr = pr
        AB2 = sum(x * x for x in pr.AB)
        self.x = pr.mu * AB2
        self.boys = boys
        self.cache = {}

    def __call__(self, a, b, N):
        if min(a) < 0 or min(b) < 0:
            return 0.0
        key = (a, b, N)
        if key in self.cache:
            return self.cache[key]
        pr = self.pr
        p = pr.p
        if sum(a) + sum(b) == 0:
            x64 = np.asarray(self.x, dtype=np.float64)
            val = 2 * pi**2.5 / np.sqrt(p) / (pr.ax * pr.bx) * self.boys(int(N), x64)
        else:
            if max(a) > 0:
                which = 0
                dd = next(d for d in range(3) if a[d] > 0)
            else:
                which = 1
                dd = next(d for d in range(3) if b[d] > 0)
            one = _e(dd)
            ang = [a, b]
            exps = (pr.ax, pr.bx)
            X = (pr.PA, pr.PB)[which][dd]
            i1, i2 = which, 1 - which
            l1 = ang[i1][dd] - 1
            l2 = ang[i2][dd]
            d1 = list(ang)
            d1[i1] = _sub(ang[i1], one)
            d11 = list(d1)
            d11[i1] = _sub(d1[i1], one)
            d12 = list(d1)
            d12[i2] = _sub(d1[i2], one)
            val = (
                X * self(d1[0], d1[1], N + 1)
                + l1 / (2 * exps[i1]) * (self(d11[0], d11[1], N) - exps[i2] / p * self(d11[0], d11[1], N + 1))
                + l2 / (2 * p) * self(d12[0], d12[1], N + 1)
            )
        self.cache[key] = val
        return val


class _ThreeC2ESph:
    """defs/coulomb.py:187-351 with aux_vrr = 'sph'."""

    def __init__(self, ax, A, bx, B, cx, C, boys):
        self.pr = pr = _Pair(ax, A, bx, B)
        self.c = cx
        self.PC = [pr.P[i] - C[i] for i in range(3)]
        self.rho = pr.p * cx / (pr.p + cx)
        self.boys = boys
        self.cache = {}

    def __call__(self, a, b, c, N):
        if min(a) < 0 or min(b) < 0 or min(c) < 0:
            return 0.0
        key = (a, b, c, N)
        if key in self.cache:
            return self.cache[key]
        pr, p, cx, rho, PC = self.pr, self.pr.p, self.c, self.rho, self.PC
        if sum(a) + sum(b) + sum(c) == 0:
            r2pc = sum(x * x for x in PC)
            r2ab = sum(x * x for x in pr.AB)
            chi = rho * r2pc
            K = np.exp(-pr.mu * r2ab)
            x64 = np.asarray(chi, dtype=np.float64)
            val = 2 * pi**2.5 / np.sqrt(p + cx) / (p * cx) * K * self.boys(int(N), x64)
        elif max(b) > 0:
            dd = next(d for d in range(3) if b[d] > 0)
            one = _e(dd)
            assert N == 0
            val = self(_add(a, one), _sub(b, one), c, N) + pr.AB[dd] * self(a, _sub(b, one), c, N)
        elif max(c) > 0:
            dd = next(d for d in range(3) if c[d] > 0)
            one = _e(dd)
            cm = _sub(c, one)
            La = a[dd]
            val = rho / cx * (PC[dd] * self(a, b, cm, N + 1) + La / (2 * p) * self(_sub(a, one), b, cm, N + 1))
        else:
            dd = next(d for d in range(3) if a[d] > 0)
            one = _e(dd)
            am = _sub(a, one)
            amm = _sub(am, one)
            ai = a[dd] - 1
            val = (
                pr.PA[dd] * self(am, b, c, N)
                - rho / p * PC[dd] * self(am, b, c, N + 1)
                + ai / (2 * p) * (self(amm, b, c, N) - rho / p * self(amm, b, c, N + 1))
            )
        self.cache[key] = val
        return val


# --------------------------------------------------------------------------
# block assembly: coefficients, normalisation, primitive sum, C2S
# --------------------------------------------------------------------------
def _norm_factors(Ls, exps, normalize):
    """Per-component normalisation, broadcast over primitives (main.py:250-265)."""
    if normalize == "none":
        return None
    comps = [canonical_order(L) for L in Ls]
    out = []
    for idx in it.product(*[range(len(c)) for c in comps]):
        if normalize == "cgto":
            f = 1.0
            for L, k in zip(Ls, idx):
                f = f * lmn_factors(L)[k]
        elif normalize == "pgto":
            f = 1.0
            for L, k, ex, comp in zip(Ls, idx, exps, comps):
                f = f * norm_pgto(comp[k], ex)
        else:
            raise ValueError(normalize)
        out.append(f)
    return out


def _finish(vals, Ls, ncomp, coeffs, exps, sph, normalize):
    """vals: list (comp-major, then Cartesian a, b[, c]) of broadcast primitive arrays."""
    nper = len(vals) // ncomp
    dprod = functools.reduce(lambda x, y: x * y, coeffs)
    norms = _norm_factors(Ls, exps, normalize)
    shape = tuple(ncart(L) for L in Ls)
    dtype = np.result_type(*[np.asarray(e).dtype for e in exps], np.float64)
    out = np.zeros((ncomp,) + shape, dtype=dtype)
    for ci in range(ncomp):
        for k in range(nper):
            v = vals[ci * nper + k] * dprod
            if norms is not None:
                v = v * norms[k]
            out[ci].flat[k] = np.sum(v)
    if sph:
        Cs = [cart2sph(L).astype(dtype) for L in Ls]
        if len(Ls) == 1:
            out = np.einsum("ia,na->ni", Cs[0], out)
        elif len(Ls) == 2:
            out = np.einsum("ia,jb,nab->nij", Cs[0], Cs[1], out)
        else:
            out = np.einsum("ia,jb,kc,nabc->nijk", Cs[0], Cs[1], Cs[2], out)
    if ncomp == 1:
        out = out[0]
    return out.ravel()


def _bc2(ax, da, bx, db):
    ax = np.asarray(ax).reshape(-1, 1)
    da = np.asarray(da).reshape(-1, 1)
    bx = np.asarray(bx).reshape(1, -1)
    db = np.asarray(db).reshape(1, -1)
    return ax, da, bx, db


def multipole(La, Lb, Le_tot, ax, da, A, bx, db, B, R=(0.0, 0.0, 0.0), sph=True, normalize="cgto"):
    """ovlp (Le_tot=0), dpm (1), qpm (2); defs/multipole.py:35-47."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    pr = _Pair(ax, A, bx, B, R)
    m1 = [_Multipole1d(pr, d) for d in range(3)]
    vals = []
    for Le in canonical_order(Le_tot):
        for a in canonical_order(La):
            for b in canonical_order(Lb):
                vals.append(m1[0](a[0], b[0], Le[0]) * m1[1](a[1], b[1], Le[1]) * m1[2](a[2], b[2], Le[2]))
    return _finish(vals, (La, Lb), ncart(Le_tot), (da, db), (ax, bx), sph, normalize)


def ovlp(La, Lb, ax, da, A, bx, db, B, R=None, **kw):
    return multipole(La, Lb, 0, ax, da, A, bx, db, B, (0.0, 0.0, 0.0), **kw)


def dpm(La, Lb, ax, da, A, bx, db, B, R=(0.0, 0.0, 0.0), **kw):
    return multipole(La, Lb, 1, ax, da, A, bx, db, B, R, **kw)


def qpm(La, Lb, ax, da, A, bx, db, B, R=(0.0, 0.0, 0.0), **kw):
    return multipole(La, Lb, 2, ax, da, A, bx, db, B, R, **kw)


def dqpm(La, Lb, ax, da, A, bx, db, B, R=(0.0, 0.0, 0.0), sph=True, normalize="cgto"):
    """defs/multipole.py:50-57."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    pr = _Pair(ax, A, bx, B, R)
    m1 = [_Multipole1d(pr, d) for d in range(3)]
    vals = []
    for Le in ((2, 0, 0), (0, 2, 0), (0, 0, 2)):
        for a in canonical_order(La):
            for b in canonical_order(Lb):
                vals.append(m1[0](a[0], b[0], Le[0]) * m1[1](a[1], b[1], Le[1]) * m1[2](a[2], b[2], Le[2]))
    return _finish(vals, (La, Lb), 3, (da, db), (ax, bx), sph, normalize)


def kin(La, Lb, ax, da, A, bx, db, B, R=None, sph=True, normalize="cgto"):
    """defs/kinetic.py:47-56."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    pr = _Pair(ax, A, bx, B)
    T = [_Kinetic1d(pr, d) for d in range(3)]
    vals = []
    for a in canonical_order(La):
        for b in canonical_order(Lb):
            Tx, Ty, Tz = [T[d](a[d], b[d]) for d in range(3)]
            Sx, Sy, Sz = [T[d].ovlp(a[d], b[d]) for d in range(3)]
            vals.append(Tx * Sy * Sz + Sx * Ty * Sz + Sx * Sy * Tz)
    return _finish(vals, (La, Lb), 1, (da, db), (ax, bx), sph, normalize)


def coul(La, Lb, ax, da, A, bx, db, B, R=(0.0, 0.0, 0.0), sph=True, normalize="cgto", boys=None):
    """Unit positive charge at R (defs/coulomb.py:95-101)."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    pr = _Pair(ax, A, bx, B, R)
    C = _Coulomb(pr, boys or _boys_default)
    vals = [C(a, b, 0) for a in canonical_order(La) for b in canonical_order(Lb)]
    return _finish(vals, (La, Lb), 1, (da, db), (ax, bx), sph, normalize)


def int2c2e(La, Lb, ax, da, A, bx, db, B, R=None, sph=True, normalize="cgto", boys=None):
    """defs/coulomb.py:178-184."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    pr = _Pair(ax, A, bx, B)
    T = _TwoC2E(pr, boys or _boys_default)
    vals = [T(a, b, 0) for a in canonical_order(La) for b in canonical_order(Lb)]
    return _finish(vals, (La, Lb), 1, (da, db), (ax, bx), sph, normalize)


def int3c2e_sph(La, Lb, Lc, ax, da, A, bx, db, B, cx, dc, C, R=None, normalize="cgto", boys=None):
    """defs/coulomb.py:368-377; always spherical (main.py:934-940)."""
    ax = np.asarray(ax).reshape(-1, 1, 1)
    da = np.asarray(da).reshape(-1, 1, 1)
    bx = np.asarray(bx).reshape(1, -1, 1)
    db = np.asarray(db).reshape(1, -1, 1)
    cx = np.asarray(cx).reshape(1, 1, -1)
    dc = np.asarray(dc).reshape(1, 1, -1)
    T = _ThreeC2ESph(ax, A, bx, B, cx, C, boys or _boys_default)
    vals = [
        T(a, b, c, 0)
        for a in canonical_order(La)
        for b in canonical_order(Lb)
        for c in canonical_order(Lc)
    ]
    return _finish(vals, (La, Lb, Lc), 1, (da, db, dc), (ax, bx, cx), True, normalize)


def gto(La, ax, da, A, R, sph=True, normalize="cgto"):
    """Values of the shell's functions at the point R (defs/gto.py:9-28, main.py:520-549):
    sum_k d_k N (R-A)^lmn exp(-a_k |R-A|^2), then lmn/pgto normalisation and C2S."""
    ax = np.asarray(ax).reshape(-1)
    da = np.asarray(da).reshape(-1)
    RA = [R[i] - A[i] for i in range(3)]
    vals = [RA[0] ** l * np.exp(-ax * RA[0] ** 2) * RA[1] ** m * np.exp(-ax * RA[1] ** 2) * RA[2] ** n * np.exp(-ax * RA[2] ** 2)
            for (l, m, n) in canonical_order(La)]
    return _finish(vals, (La,), 1, (da,), (ax,), sph, normalize)


def prefactor(La, Lb, ax, da, A, bx, db, B, R=(0.0, 0.0, 0.0), sph=True, normalize="cgto"):
    """``prefactors`` (defs/overlap.py:16-34, main.py:481-518): da db rP^(a+b) per Cartesian component pair, then
    normalisation and C2S on both indices.  rP is NOT an argument of the reference's generated functions (they read
    module-level ``rP`` / ``rP2``); it is passed as R here, with rP2 = rP**2."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    vals = [R[0] ** (a[0] + b[0]) * R[1] ** (a[1] + b[1]) * R[2] ** (a[2] + b[2]) * (ax * 0 + 1.0) * (bx * 0 + 1.0)
            for a in canonical_order(La) for b in canonical_order(Lb)]
    return _finish(vals, (La, Lb), 1, (da, db), (ax, bx), sph, normalize)


# real regular solid harmonics R_lm, l <= 2, m = -l..l, as polynomials in (x, y, z): {exponents: coefficient}
# (sympleints/sym_solid_harmonics.py Rlm_polys_up_to_l_iter(2, order=mOrder.NATURAL), evaluated once)
_SQ3 = sqrt(3.0)
RLM_POLYS = [
    {(0, 0, 0): 1.0},
    {(0, 1, 0): 1.0}, {(0, 0, 1): 1.0}, {(1, 0, 0): 1.0},
    {(1, 1, 0): _SQ3}, {(0, 1, 1): _SQ3}, {(2, 0, 0): -0.5, (0, 2, 0): -0.5, (0, 0, 2): 1.0}, {(1, 0, 1): _SQ3},
    {(2, 0, 0): 0.5 * _SQ3, (0, 2, 0): -0.5 * _SQ3},
]


def multi_sph(La, Lb, ax, da, A, bx, db, B, R=None, sph=True, normalize="cgto"):
    """``multipole3d_sph`` (defs/multipole.py:60-138, main.py:752-800): integrals over the nine spherical multipole
    operators R_lm, l <= 2, about the centre P = (aA + bB)/(a + b) of EVERY primitive pair; operators of order
    l > La + Lb are set to zero as the reference does (multipole.py:118-121).  Primitive functions in the reference;
    several primitives are summed here, each pair with its own origin."""
    ax, da, bx, db = _bc2(ax, da, bx, db)
    A = [A[i] for i in range(3)]
    B = [B[i] for i in range(3)]
    pr = _Pair(ax, A, bx, B, R=None)
    pr.PR = [0.0 * pr.P[i] for i in range(3)]     # origin = P of each primitive pair
    m1 = [_Multipole1d(pr, d) for d in range(3)]
    vals = []
    for poly in RLM_POLYS:
        order = sum(next(iter(poly)))
        for a in canonical_order(La):
            for b in canonical_order(Lb):
                if order > La + Lb:
                    vals.append(0.0 * ax * bx)
                    continue
                v = 0.0
                for Le, coef in poly.items():
                    v = v + coef * m1[0](a[0], b[0], Le[0]) * m1[1](a[1], b[1], Le[1]) * m1[2](a[2], b[2], Le[2])
                vals.append(v)
    return _finish(vals, (La, Lb), 9, (da, db), (ax, bx), sph, normalize)


KEY_FUNCS = {
    "ovlp": (ovlp, 1),
    "kin": (kin, 1),
    "dpm": (dpm, 3),
    "qpm": (qpm, 6),
    "dqpm": (dqpm, 3),
    "coul": (coul, 1),
    "2c2e": (int2c2e, 1),
    "multi_sph": (multi_sph, 9),
}
