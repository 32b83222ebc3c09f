"""Schwarz integrals (ab|ab) (sympleints/graphs/defs/int_4c2e.py, l_iters.py:23-28).

The reference has this integral only on its Fortran graph path, so the oracle (oracle/schwarz.py) is UNPINNED (no
reference OUTPUT exists to compare with); these CPU tests tie it to what can be checked: a literal restatement of the
reference's own transformations (vrr0, vrr2, hrr1, hrr3, c2s of int_4c2e.py) against the McMurchie-Davidson evaluation,
the closed form for s functions, numerical differentiation, and the Cauchy-Schwarz inequality against the pinned
2c2e / 3c2e oracle.  The GPU tests compare the CUDA kernel with it."""
import numpy as np
import pytest

from oracle import boys, ints, schwarz


def test_ssss_closed_form_and_positivity():
    rng = np.random.default_rng(0)
    for _ in range(5):
        a, b = 10 ** rng.uniform(-1, 1, 2)
        A, B = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        v = schwarz.schwarz_block(0, 0, [a], [1.0], A, [b], [1.0], B, normalize="none")[0, 0]
        p = a + b
        K = np.exp(-a * b / p * np.dot(A - B, A - B))
        ref = 2 * np.pi**2.5 / (p * p * np.sqrt(2 * p)) * K * K * boys.boys(0, 0.0)
        assert abs(v - ref) <= 1e-14 * ref
    for (La, Lb) in [(1, 1), (2, 1), (3, 2), (2, 2)]:
        ea, eb = 10 ** rng.uniform(-1, 1, 2), 10 ** rng.uniform(-1, 1, 1)
        D = schwarz.schwarz_block(La, Lb, ea, ints.norm_cgto_lmn(rng.uniform(.2, 1, 2), ea, La), rng.uniform(-1, 1, 3),
                                  eb, ints.norm_cgto_lmn(np.ones(1), eb, Lb), rng.uniform(-1, 1, 3))
        assert D.shape == (2 * La + 1, 2 * Lb + 1) and (D > 0).all()


def test_angular_momentum_by_numerical_differentiation():
    """(p_x b|cd) = 1/(2a) d/dA_x (s_A b|cd) for un-normalised Cartesian primitives: ties the Hermite expansion and
    the R recursion of higher angular momentum to the s-type closed form."""
    rng = np.random.default_rng(1)
    a, b, c, d = 0.9, 0.6, 1.1, 0.4
    A, B, C, D = (rng.uniform(-1, 1, 3) for _ in range(4))
    one = [1.0]

    def ssss(Ax):
        return schwarz.eri_block((0, 0, 0, 0), ([a], one, Ax, [b], one, B), ([c], one, C, [d], one, D),
                                 sph=False, normalize="none")[0, 0, 0, 0]

    pxs = schwarz.eri_block((1, 0, 0, 0), ([a], one, A, [b], one, B), ([c], one, C, [d], one, D), sph=False, normalize="none")
    h = 1e-4
    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        fd = (ssss(A + e) - ssss(A - e)) / (2 * h) / (2 * a)
        assert abs(pxs[k, 0, 0, 0] - fd) <= 1e-7 * max(1.0, abs(fd))
    # second derivative: (d_xx s|ss) = 1/(4a^2) d2/dAx2 (ss|ss) + 1/(2a) (ss|ss)
    dss = schwarz.eri_block((2, 0, 0, 0), ([a], one, A, [b], one, B), ([c], one, C, [d], one, D), sph=False, normalize="none")
    e = np.array([h, 0, 0])
    d2 = (ssss(A + e) - 2 * ssss(A) + ssss(A - e)) / h**2
    assert abs(dss[0, 0, 0, 0] - (d2 / (4 * a * a) + ssss(A) / (2 * a))) <= 1e-5 * abs(dss[0, 0, 0, 0])


def _rshell(rng, L, K):
    ex = 10 ** rng.uniform(-0.5, 0.8, K)
    return ex, ints.norm_cgto_lmn(rng.uniform(0.3, 1, K), ex, L), rng.uniform(-1, 1, 3)


@pytest.mark.parametrize("Ls", [(0, 0, 0, 0), (1, 0, 0, 0), (1, 1, 1, 0), (2, 1, 1, 1), (2, 0, 1, 2), (3, 1, 2, 0), (2, 2, 2, 2)])
def test_reference_recurrences_against_mcmurchie_davidson_full_blocks(Ls):
    """(ab|cd) for four different contracted shells: the reference's Obara-Saika / HRR transformations
    (graphs/defs/int_4c2e.py:21-75, restated literally in eri_block_os) and the McMurchie-Davidson evaluation the
    Schwarz oracle and the CUDA kernel use are two independent algorithms; every element of the block must agree."""
    rng = np.random.default_rng(sum(10**k * l for k, l in enumerate(Ls)))
    s = [_rshell(rng, L, 2) for L in Ls]
    bra, ket = (*s[0], *s[1]), (*s[2], *s[3])
    o = schwarz.eri_block_os(Ls, bra, ket)
    m = schwarz.eri_block(Ls, bra, ket)
    assert o.shape == tuple(2 * L + 1 for L in Ls)
    assert np.abs(o - m).max() <= 1e-12 * np.abs(m).max()
    # Cartesian, un-normalised variant (no C2S, no lmn factors)
    oc = schwarz.eri_block_os(Ls, bra, ket, sph=False, normalize="none")
    mc = schwarz.eri_block(Ls, bra, ket, sph=False, normalize="none")
    assert np.abs(oc - mc).max() <= 1e-12 * np.abs(mc).max()


@pytest.mark.parametrize("La,Lb,tol", [(3, 2, 1e-12), (3, 3, 1e-10), (4, 1, 1e-12), (4, 3, 1e-10)])
def test_schwarz_diagonal_against_the_reference_recurrences(La, Lb, tol):
    """The diagonal (ab|ab) the oracle exposes against the diagonal of the block built with the reference's recurrences
    (the HRR loses digits for f and g kets, SURVEY 7.2, hence the looser bound there)."""
    rng = np.random.default_rng(10 * La + Lb)
    a, b = _rshell(rng, La, 1), _rshell(rng, Lb, 2 if La + Lb < 7 else 1)
    sh = (*a, *b)
    d = schwarz.schwarz_block(La, Lb, *sh)
    o = np.einsum("ijij->ij", schwarz.eri_block_os((La, Lb, La, Lb), sh, sh))
    assert (d > 0).all() and np.abs(o - d).max() <= tol * np.abs(d).max()


@pytest.mark.parametrize("Ls", [(1, 0, 1), (2, 1, 2), (2, 2, 3), (3, 1, 0), (3, 3, 4)])
def test_cauchy_schwarz_against_the_pinned_oracle(Ls):
    """|(ab|P)|^2 <= (ab|ab) (P|P) with (ab|P), (P|P) from the oracle that is pinned on the reference's code."""
    La, Lb, Lc = Ls
    rng = np.random.default_rng(sum(Ls))
    ea, eb, ec = 10 ** rng.uniform(-1, 1, 2), 10 ** rng.uniform(-1, 1, 2), 10 ** rng.uniform(-0.5, 1, 1)
    A, B, C = (rng.uniform(-1.2, 1.2, 3) for _ in range(3))
    da, db, dc = (ints.norm_cgto_lmn(rng.uniform(.2, 1, len(e)), e, L) for e, L in ((ea, La), (eb, Lb), (ec, Lc)))
    D = schwarz.schwarz_block(La, Lb, ea, da, A, eb, db, B)
    T = ints.int3c2e_sph(La, Lb, Lc, ea, da, A, eb, db, B, ec, dc, C).reshape(2 * La + 1, 2 * Lb + 1, 2 * Lc + 1)
    M = ints.int2c2e(Lc, Lc, ec, dc, C, ec, dc, C).reshape(2 * Lc + 1, 2 * Lc + 1)
    ratio = T**2 / (D[:, :, None] * np.diag(M)[None, None, :])
    assert ratio.max() <= 1.0 + 1e-10
    # and the bound is not vacuous: an aux function sitting on the pair's charge centre gets within a factor of it
    assert ratio.max() > 1e-4


@pytest.mark.gpu
def test_schwarz_kernel_against_the_oracle_and_the_inequality():
    """si_schwarz: D[mu, nu] = (mu nu|mu nu) per block against oracle/schwarz.py (same Boys table), the shell-pair
    factors Q, and the Cauchy-Schwarz inequality for EVERY element of the GPU's own 3c2e tensor and 2c2e metric."""
    import sympleints_b200
    from oracle import boys as oboys
    from sympleints_b200 import synthetic

    bf = oboys.get_boys_func(16)      # (gg|gg) needs Boys orders up to 16
    eng = sympleints_b200.get_engine(0, 4, 5, boys_table=bf.table)
    shells, sites = synthetic.orbital_basis(natoms=2, seed=31)
    aux = synthetic.aux_basis(sites, seed=32)
    ob, ab = eng.basis(shells), eng.basis(aux)
    D, Q = eng.schwarz(ob)
    D, Q = D.cpu().numpy(), Q.cpu().numpy()
    off = np.concatenate([[0], np.cumsum([2 * s.L + 1 for s in shells])])
    assert np.array_equal(D, D.T), float(np.abs(D - D.T).max())
    assert (D >= 0).all(), float(D.min())
    rng = np.random.default_rng(2)
    pairs = [(i, j) for i in range(len(shells)) for j in range(len(shells))]
    for k in rng.permutation(len(pairs))[:40]:
        i, j = pairs[k]
        a, b = shells[i], shells[j]
        o = schwarz.schwarz_block(a.L, b.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center, boys=bf)
        g = D[off[i]:off[i + 1], off[j]:off[j + 1]]
        assert np.abs(g - o).max() <= 1e-14 + 1e-11 * np.abs(o).max(), (i, j, a.L, b.L)
        assert abs(Q[i, j] - np.sqrt(np.abs(o).max())) <= 1e-11 * Q[i, j]
    # ... and directly against the literal restatement of the reference's recurrences (one pair per class up to f)
    seen = set()
    for (i, j) in pairs:
        a, b = shells[i], shells[j]
        if a.L > 3 or b.L > 3 or (a.L, b.L) in seen or len(a.exps) * len(b.exps) > 6:
            continue
        seen.add((a.L, b.L))
        sh = (a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center)
        o = np.einsum("ijij->ij", schwarz.eri_block_os((a.L, b.L, a.L, b.L), sh, sh, boys=bf))
        g = D[off[i]:off[i + 1], off[j]:off[j + 1]]
        assert np.abs(g - o).max() <= 1e-14 + 1e-10 * np.abs(o).max(), (i, j, a.L, b.L)
    assert len(seen) >= 10
    T = eng.int3c2e(ob, ab, layout="full").cpu().numpy()
    M = eng.int2c2e(ab).cpu().numpy()
    bound = D[:, :, None] * np.diag(M)[None, None, :]
    assert (T**2 <= bound * (1 + 1e-9) + 1e-28).all()
    assert (T**2 / bound).max() > 0.05      # the bound is not vacuous
    # a table that is too small is reported, as the reference's Boys function would raise IndexError
    small = sympleints_b200.get_engine(0, 4, 5, boys_table=oboys.table())   # nmax = 14 < 16
    with pytest.raises(sympleints_b200.SympleintsB200Error, match="Boys"):
        small.schwarz(small.basis(shells))
    small.close()
    eng.close()
