"""Oracle Boys function F_n(x) -- restatement of sympleints/testing/boys.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference algorithm step for step so that values agree to the last
bit (pinned in tests/test_oracle_boys.py against tests/golden/boys_golden.npz,
which was produced by the reference's own ``get_boys_func``):

* table row ``nmax+order`` on x = 0, 0.1, ..., 30 from ``scipy.integrate.quad``
  with epsabs = epsrel = 1e-13                      (testing/boys.py:8-29)
* all lower rows by downward recursion on the grid
  ``F_n = (2x F_{n+1} + exp(-x)) / (2n+1)``           (testing/boys.py:81-91)
* evaluation: x == 0 -> table[n, 0]; x < 30 -> 6-term Taylor expansion around
  the grid point at/below x, ``ind0 = int(x*10)``, ``xg = ind0*0.1``;
  else ``(2n-1)!!/2^(n+1) sqrt(pi) x^(-n-1/2)``       (testing/boys.py:96-140)

The generated reference modules under oracle/_ref import ``boys`` from here
(``--boys-func oracle.boys``), because the reference's default instance has
nmax=8 which is too small for 2c2e (h|h) and 3c2e (gg|h).
"""
import numpy as np
import scipy as sp
from scipy.special import factorial, factorial2

NMAX_DEFAULT = 14
ORDER = 6
DX = 0.1
XLARGE = 30.0


def boys_quad(n, x, err=1e-13):
    def inner(t):
        return t ** (2 * n) * np.exp(-x * t**2)

    y, _ = sp.integrate.quad(inner, 0.0, 1.0, epsabs=err, epsrel=err)
    return y


def build_table(nmax=NMAX_DEFAULT, order=ORDER, dx=DX, xmax=XLARGE):
    """Full table, shape (nmax+order+1, 301)."""
    nsteps = int((xmax / dx) + 1)
    xs = np.linspace(0.0, xmax, num=nsteps)
    top = np.empty_like(xs)
    for i, x in enumerate(xs):
        top[i] = boys_quad(nmax + order, x)
    table = np.empty((nmax + order + 1, nsteps))
    table[-1] = top
    emx = np.exp(-xs)
    _2xs = 2 * xs
    for n in range(nmax + order - 1, -1, -1):
        table[n] = (_2xs * table[n + 1] + emx) / (2 * n + 1)
    return xs, table


def xlarge_prefactors(nmax):
    pref = np.empty(nmax + 1)
    for n in range(nmax + 1):
        pref[n] = (factorial2(2 * n - 1) if n > 0 else 1.0) / 2 ** (n + 1) * np.sqrt(np.pi)
    return pref


def get_boys_func(nmax=NMAX_DEFAULT, order=ORDER, dx=DX, xlarge=XLARGE, table=None):
    if table is None:
        _, table = build_table(nmax, order, dx, xlarge)
    inv_fact = 1 / factorial(np.arange(order, dtype=int))
    pref = xlarge_prefactors(nmax)
    factor = int(1 / dx)

    def scalar(n, x):
        if x == 0.0:
            return table[n, 0]
        elif x < xlarge:
            ind0 = int(x * factor)
            xg = ind0 * dx
            dxt = x - xg
            result = 0.0
            for k in range(order):
                result += table[n + k, ind0] * (-dxt) ** k * inv_fact[k]
            return result
        else:
            return pref[n] * x ** (-n - 0.5)

    def boys(n, xs):
        """Same calling convention as the reference: scalar or ndarray ``xs``;
        arrays are walked element by element (testing/boys.py:131-138)."""
        if isinstance(xs, np.ndarray):
            out = [scalar(n, x) for x in xs.ravel()]
            return np.reshape(out, xs.shape)
        return scalar(n, xs)

    def boys_vec(n, xs):
        """Vectorised evaluation with identical arithmetic per element
        (used by the tests to keep them fast)."""
        xs = np.asarray(xs, dtype=float)
        flat = xs.ravel()
        out = np.empty_like(flat)
        small = flat < xlarge
        xsm = flat[small]
        ind0 = (xsm * factor).astype(np.int64)
        xg = ind0 * dx
        dxt = xsm - xg
        res = np.zeros_like(xsm)
        for k in range(order):
            res += table[n + k, ind0] * (-dxt) ** k * inv_fact[k]
        zero = xsm == 0.0
        res[zero] = table[n, 0]
        out[small] = res
        out[~small] = pref[n] * flat[~small] ** (-n - 0.5)
        return out.reshape(xs.shape)

    boys.table = table
    boys.vec = boys_vec
    boys.scalar = scalar
    boys.nmax = nmax
    return boys


_CACHE = {}


def _default():
    if "b" not in _CACHE:
        _CACHE["b"] = get_boys_func(NMAX_DEFAULT)
    return _CACHE["b"]


def boys(n, xs):
    """Module-level instance used by the generated reference code (nmax=14)."""
    return _default()(n, xs)


def boys_vec(n, xs):
    return _default().vec(n, xs)


def table():
    return _default().table
