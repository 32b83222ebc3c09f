"""TEST INFRASTRUCTURE -- CPU restatement of the Schwarz integrals (ab|ab) of sympleints' graph path.

PARITY UNPINNED: the reference defines the four-centre integral only on its graph/Fortran path
(sympleints/graphs/defs/int_4c2e.py:11-75, classes from l_iters.py:23-28 ``schwarz_iter``), which cannot be
compiled here (no gfortran) and has no Python rendering, golden vector or test.  What is restated is the published
definition used there (base ``2 pi^2.5 / (p q sqrt(p+q)) K_AB K_CD F_n(alpha |PQ|^2)``, all four indices spherical,
lmn normalisation as for every other key), evaluated with the McMurchie-Davidson scheme (Helgaker, Jorgensen, Olsen,
Molecular Electronic-Structure Theory, sections 9.5 and 9.9):

    (ab|cd) = 2 pi^2.5 / (p q sqrt(p+q)) sum_{tuv} E^{ab}_{tuv} sum_{t'u'v'} (-1)^{t'+u'+v'} E^{cd}_{t'u'v'} R_{t+t',u+u',v+v'}(alpha, P-Q)

``eri_block_os`` restates the reference's own transformations of that file literally (vrr0, vrr2, hrr1, hrr3 and the four
C2S steps, int_4c2e.py:21-75) -- the algorithm the reference would run -- and agrees with the McMurchie-Davidson
evaluation on every element of full four-index blocks to 1e-12 (tests/test_schwarz.py).  Further checks there: the
closed form for s functions, numerical differentiation of lower angular momenta, the PINNED 2c2e/3c2e oracles through
the Cauchy-Schwarz inequality |(ab|P)|^2 <= (ab|ab)(P|P), and positivity.  What stays impossible is a comparison with
reference OUTPUT, hence "unpinned".
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


def eri_block_os(Ls, bra, ket, sph=True, normalize="cgto", boys=None):
    """(ab|cd) by a LITERAL restatement of the reference's transformations (sympleints/graphs/defs/int_4c2e.py):

        base   2 pi^2.5 / (px qx sqrt(px+qx)) K_AB K_CD F_n(rho |PQ|^2)                                  (:21-23)
        vrr0   [a+1_i 0|c 0]^n = PA_i [a|c]^n - qx/(px+qx) PQ_i [a|c]^{n+1}
                                 + a_i/(2 px) ([a-1_i|c]^n - qx/(px+qx) [a-1_i|c]^{n+1})
                                 + c_i/(2 (px+qx)) [a|c-1_i]^{n+1}                                      (:25-35)
        vrr2   [a 0|c+1_i 0]^n = QC_i [a|c]^n + px/(px+qx) PQ_i [a|c]^{n+1}
                                 + c_i/(2 qx) ([a|c-1_i]^n - px/(px+qx) [a|c-1_i]^{n+1})
                                 + a_i/(2 (px+qx)) [a-1_i|c]^{n+1}                                      (:37-47)
        hrr1   (a b+1_i|..) = (a+1_i b|..) + AB_i (a b|..)      hrr3   (..|c d+1_i) = (..|c+1_i d) + CD_i (..|c d)   (:49-68)
        c2s on all four indices                                                                           (:55-74)

    evaluated by memoised recursion per primitive quadruple (plain Python: small angular momenta only).  An algorithm
    independent of the McMurchie-Davidson evaluation of eri_block; tests/test_schwarz.py checks that the two agree."""
    import itertools
    from functools import lru_cache

    boys = boys or oboys.boys
    if normalize not in ("cgto", "none"):
        raise ValueError("eri_block_os: cgto / none only")
    La, Lb, Lc, Ld = Ls
    (ax, da, A, bx, db, B), (cx, dc, C, dx, dd, D) = bra, ket
    ax, da, bx, db, cx, dc, dx, dd = (np.asarray(x, dtype=float).reshape(-1) for x in (ax, da, bx, db, cx, dc, dx, dd))
    A, B, C, D = (np.asarray(x, dtype=float) for x in (A, B, C, D))
    AB, CD = A - B, C - D
    a_shells = [canonical_order(n) for n in range(La + Lb + 1)]
    c_shells = [canonical_order(n) for n in range(Lc + Ld + 1)]
    acc = {}
    for (a, d1), (b, d2), (c, d3), (d, d4) in itertools.product(zip(ax, da), zip(bx, db), zip(cx, dc), zip(dx, dd)):
        px, qx = a + b, c + d
        P, Q = (a * A + b * B) / px, (c * C + d * D) / qx
        PA, QC, PQ = P - A, Q - C, P - Q
        KAB, KCD = np.exp(-a * b / px * AB @ AB), np.exp(-c * d / qx * CD @ CD)
        rho = px * qx / (px + qx)
        pre = 2 * np.pi**2.5 / (px * qx * np.sqrt(px + qx)) * KAB * KCD * d1 * d2 * d3 * d4
        nmax = La + Lb + Lc + Ld
        F = [pre * float(boys(n, rho * PQ @ PQ)) for n in range(nmax + 1)]

        @lru_cache(maxsize=None)
        def I(av, cv, n):
            if min(av) < 0 or min(cv) < 0:
                return 0.0
            if sum(av) == 0 and sum(cv) == 0:
                return F[n]
            if sum(av) > 0:
                i = next(k for k in range(3) if av[k] > 0)
                am = tuple(av[k] - (k == i) for k in range(3))
                amm = tuple(am[k] - (k == i) for k in range(3))
                cm = tuple(cv[k] - (k == i) for k in range(3))
                val = PA[i] * I(am, cv, n) - qx / (px + qx) * PQ[i] * I(am, cv, n + 1)
                if am[i] > 0:
                    val += am[i] / (2 * px) * (I(amm, cv, n) - qx / (px + qx) * I(amm, cv, n + 1))
                if cv[i] > 0:
                    val += cv[i] / (2 * (px + qx)) * I(am, cm, n + 1)
                return val
            i = next(k for k in range(3) if cv[k] > 0)
            cm = tuple(cv[k] - (k == i) for k in range(3))
            cmm = tuple(cm[k] - (k == i) for k in range(3))
            val = QC[i] * I(av, cm, n) + px / (px + qx) * PQ[i] * I(av, cm, n + 1)
            if cm[i] > 0:
                val += cm[i] / (2 * qx) * (I(av, cmm, n) - px / (px + qx) * I(av, cmm, n + 1))
            return val                      # (a == 0 here: the a_i / (2 (px+qx)) term vanishes)

        for na in range(La, La + Lb + 1):
            for av in a_shells[na]:
                for nc in range(Lc, Lc + Ld + 1):
                    for cv in c_shells[nc]:
                        acc[(av, cv)] = acc.get((av, cv), 0.0) + I(av, cv, 0)

    def hrr_b(av, bv, cv):              # (a b | c 0), contracted
        if sum(bv) == 0:
            return acc[(av, cv)]
        i = next(k for k in range(3) if bv[k] > 0)
        bm = tuple(bv[k] - (k == i) for k in range(3))
        ap = tuple(av[k] + (k == i) for k in range(3))
        return hrr_b(ap, bm, cv) + AB[i] * hrr_b(av, bm, cv)

    ab_c = {}

    def hrr_d(av, bv, cv, dv):
        if sum(dv) == 0:
            key = (av, bv, cv)
            if key not in ab_c:
                ab_c[key] = hrr_b(av, bv, cv)
            return ab_c[key]
        i = next(k for k in range(3) if dv[k] > 0)
        dm = tuple(dv[k] - (k == i) for k in range(3))
        cp = tuple(cv[k] + (k == i) for k in range(3))
        return hrr_d(av, bv, cp, dm) + CD[i] * hrr_d(av, bv, cv, dm)

    comps = [canonical_order(L) for L in Ls]
    out = np.zeros(tuple(len(c) for c in comps))
    for (i, av), (j, bv), (k, cv), (l, dv) in itertools.product(*(enumerate(c) for c in comps)):
        out[i, j, k, l] = hrr_d(tuple(av), tuple(bv), tuple(cv), tuple(dv))
    if normalize == "cgto":
        for axis, L in enumerate(Ls):
            shape = [1, 1, 1, 1]
            shape[axis] = -1
            out = out * np.asarray(lmn_factors(L)).reshape(shape)
    if sph:
        out = np.einsum("ia,jb,kc,ld,abcd->ijkl", cart2sph(La), cart2sph(Lb), cart2sph(Lc), cart2sph(Ld), out)
    return out


def schwarz_block(La, Lb, ax, da, A, bx, db, B, sph=True, normalize="cgto", boys=None):
    """Diagonal (ab|ab) of the contracted shell pair: array (na, nb) of (a_i b_j | a_i b_j)."""
    sh = (ax, da, A, bx, db, B)
    return eri_block((La, Lb, La, Lb), sh, sh, sph=sph, normalize=normalize, boys=boys, diagonal=True)


def schwarz_q(La, Lb, *args, **kw):
    """Q_ab = sqrt(max_ij |(a_i b_j | a_i b_j)|): the Schwarz factor of the shell pair."""
    return float(np.sqrt(np.abs(schwarz_block(La, Lb, *args, **kw)).max()))
