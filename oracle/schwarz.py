"""TEST INFRASTRUCTURE -- CPU restatement of the Schwarz integrals (ab|ab) of sympleints' graph path.

PARITY UNPINNED: the reference defines the four-centre integral only on its graph/Fortran path
(sympleints/graphs/defs/int_4c2e.py:11-75, classes from l_iters.py:23-28 ``schwarz_iter``), which cannot be
compiled here (no gfortran) and has no Python rendering, golden vector or test.  What is restated is the published
definition used there (base ``2 pi^2.5 / (p q sqrt(p+q)) K_AB K_CD F_n(alpha |PQ|^2)``, all four indices spherical,
lmn normalisation as for every other key), evaluated with the McMurchie-Davidson scheme (Helgaker, Jorgensen, Olsen,
Molecular Electronic-Structure Theory, sections 9.5 and 9.9):

    (ab|cd) = 2 pi^2.5 / (p q sqrt(p+q)) sum_{tuv} E^{ab}_{tuv} sum_{t'u'v'} (-1)^{t'+u'+v'} E^{cd}_{t'u'v'} R_{t+t',u+u',v+v'}(alpha, P-Q)

It is checked (tests/test_schwarz.py) against the closed form for s functions, against numerical differentiation of
lower angular momenta, against the PINNED 2c2e/3c2e oracles through the Cauchy-Schwarz inequality
|(ab|P)|^2 <= (ab|ab)(P|P), and for positivity.  Only the diagonal (ab|ab) is exposed.
"""
import numpy as np

from . import boys as oboys
from .ints import canonical_order, cart2sph, lmn_factors, ncart, norm_pgto


def _e1d(la, lb, a, b, XA, XB):
    """E[i, j, t] for one direction: Hermite expansion coefficients of x_A^i x_B^j exp(-a x_A^2 - b x_B^2)."""
    p = a + b
    mu = a * b / p
    P = (a * XA + b * XB) / p
    PA, PB = P - XA, P - XB
    E = np.zeros((la + 1, lb + 1, la + lb + 2))
    E[0, 0, 0] = np.exp(-mu * (XA - XB) ** 2)
    for i in range(la):
        for t in range(i + 1 + 1):
            E[i + 1, 0, t] = (E[i, 0, t - 1] / (2 * p) if t > 0 else 0.0) + PA * E[i, 0, t] + (t + 1) * E[i, 0, t + 1]
    for i in range(la + 1):
        for j in range(lb):
            for t in range(i + j + 1 + 1):
                E[i, j + 1, t] = (E[i, j, t - 1] / (2 * p) if t > 0 else 0.0) + PB * E[i, j, t] + (t + 1) * E[i, j, t + 1]
    return E[:, :, : la + lb + 1]


def _hermite_r(L, alpha, PQ, boys):
    """R[t, u, v] = R^0_{tuv}(alpha, PQ), t + u + v <= L."""
    T = alpha * float(np.dot(PQ, PQ))
    R = np.zeros((L + 1, L + 1, L + 1, L + 1))   # [n, t, u, v]
    for n in range(L + 1):
        R[n, 0, 0, 0] = (-2.0 * alpha) ** n * boys(n, T)
    for n in range(L - 1, -1, -1):
        for t in range(L - n + 1):
            for u in range(L - n - t + 1):
                for v in range(L - n - t - u + 1):
                    if t + u + v == 0:
                        continue
                    if t > 0:
                        val = (t - 1) * (R[n + 1, t - 2, u, v] if t > 1 else 0.0) + PQ[0] * R[n + 1, t - 1, u, v]
                    elif u > 0:
                        val = (u - 1) * (R[n + 1, t, u - 2, v] if u > 1 else 0.0) + PQ[1] * R[n + 1, t, u - 1, v]
                    else:
                        val = (v - 1) * (R[n + 1, t, u, v - 2] if v > 1 else 0.0) + PQ[2] * R[n + 1, t, u, v - 1]
                    R[n, t, u, v] = val
    return R[0]


def _e_sph(La, Lb, a, da, A, b, db, B, sph, normalize):
    """Spherical (or Cartesian) expansion coefficients E[(i, j), t, u, v] of one primitive pair, with the contraction
    coefficients, the normalisation and C2S on both indices applied."""
    L = La + Lb
    Ex, Ey, Ez = (_e1d(La, Lb, a, b, A[d], B[d]) for d in range(3))
    ca, cb = canonical_order(La), canonical_order(Lb)
    E = np.zeros((len(ca), len(cb), L + 1, L + 1, L + 1))
    for i, (ax, ay, az) in enumerate(ca):
        for j, (bx, by, bz) in enumerate(cb):
            blk = np.einsum("t,u,v->tuv", Ex[ax, bx, : ax + bx + 1], Ey[ay, by, : ay + by + 1], Ez[az, bz, : az + bz + 1])
            f = da * db
            if normalize == "cgto":
                f *= lmn_factors(La)[i] * lmn_factors(Lb)[j]
            elif normalize == "pgto":
                f *= norm_pgto(ca[i], a) * norm_pgto(cb[j], b)
            E[i, j, : ax + bx + 1, : ay + by + 1, : az + bz + 1] = f * blk
    if sph:
        E = np.einsum("ia,jb,abtuv->ijtuv", cart2sph(La), cart2sph(Lb), E)
    return E


def _prim_pairs(La, Lb, ax, da, A, bx, db, B, sph, normalize):
    ax, da, bx, db = (np.asarray(x, dtype=float).reshape(-1) for x in (ax, da, bx, db))
    A, B = np.asarray(A, dtype=float), np.asarray(B, dtype=float)
    out = []
    for a, d1 in zip(ax, da):
        for b, d2 in zip(bx, db):
            p = a + b
            out.append((p, (a * A + b * B) / p, _e_sph(La, Lb, a, d1, A, b, d2, B, sph, normalize)))
    return out


def eri_block(Ls, bra, ket, sph=True, normalize="cgto", boys=None, diagonal=False):
    """(ab|cd) of two contracted shell pairs: ``bra = (ax, da, A, bx, db, B)``, ``ket`` likewise; array
    (na, nb, nc, nd), or with ``diagonal=True`` (ket == bra) the (na, nb) array of (a_i b_j | a_i b_j)."""
    boys = boys or oboys.boys
    La, Lb, Lc, Ld = Ls
    P1 = _prim_pairs(La, Lb, *bra, sph, normalize)
    P2 = _prim_pairs(Lc, Ld, *ket, sph, normalize)
    n1, n2 = La + Lb + 1, Lc + Ld + 1
    s2 = np.array([(-1.0) ** k for k in range(n2)])
    i1, i2 = np.arange(n1), np.arange(n2)
    out = None
    for (p, P, EP) in P1:
        for (q, Q, EQ) in P2:
            alpha = p * q / (p + q)
            R = _hermite_r(n1 + n2 - 2, alpha, P - Q, boys)
            M = R[i1[:, None, None, None, None, None] + i2[None, None, None, :, None, None],
                  i1[None, :, None, None, None, None] + i2[None, None, None, None, :, None],
                  i1[None, None, :, None, None, None] + i2[None, None, None, None, None, :]]
            EQs = EQ * s2[:, None, None] * s2[None, :, None] * s2[None, None, :]
            pref = 2 * np.pi**2.5 / (p * q * np.sqrt(p + q))
            if diagonal:
                val = np.einsum("ijtuv,tuvxyz,ijxyz->ij", EP, M, EQs) * pref
            else:
                val = np.einsum("ijtuv,tuvxyz,klxyz->ijkl", EP, M, EQs) * pref
            out = val if out is None else out + val
    return out


def schwarz_block(La, Lb, ax, da, A, bx, db, B, sph=True, normalize="cgto", boys=None):
    """Diagonal (ab|ab) of the contracted shell pair: array (na, nb) of (a_i b_j | a_i b_j)."""
    sh = (ax, da, A, bx, db, B)
    return eri_block((La, Lb, La, Lb), sh, sh, sph=sph, normalize=normalize, boys=boys, diagonal=True)


def schwarz_q(La, Lb, *args, **kw):
    """Q_ab = sqrt(max_ij |(a_i b_j | a_i b_j)|): the Schwarz factor of the shell pair."""
    return float(np.sqrt(np.abs(schwarz_block(La, Lb, *args, **kw)).max()))
