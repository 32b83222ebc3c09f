"""GPU parity tests: every call goes through the C ABI (libsympleints_b200.so) and is compared
with the reference's golden vectors and with the oracle on seeded inputs."""
import numpy as np
import pytest

import golden_util as gu
from oracle import boys as oboys
from oracle import ints

pytestmark = pytest.mark.gpu


def _gpu_block(engine, blk, sph, normalize):
    key, Ls = blk["key"], blk["Ls"]
    a, b, c = blk["a"], blk["b"], blk["c"]
    if key == "3c2e":
        return engine.eval_block("3c2e", Ls, *a, *b, *c, R=blk["R"], normalize=normalize)
    if key == "gto":
        return engine.eval_block("gto", Ls[:1], *a, R=blk["R"], sph=sph, normalize=normalize)
    return engine.eval_block(key, Ls, *a, *b, R=blk["R"], sph=sph, normalize=normalize)


def test_boys_kernel_matches_reference_values(engine):
    z = np.load(gu.GOLDEN / "boys_golden.npz")
    assert np.array_equal(engine.boys_table(), z["table"])
    xs, vals = z["xs"], z["vals"]
    for n in range(vals.shape[0]):
        got = engine.boys(n, xs)
        # same table, same cell, same six Taylor terms; only pow()/FMA contraction may differ
        np.testing.assert_allclose(got, vals[n], rtol=5e-15, atol=0, err_msg=f"n={n}")


@pytest.mark.parametrize("mode,rtol", [("fast", 5e-15), ("coul", 5e-15)])
def test_hot_path_boys_variants_match_reference_values(engine, mode, rtol):
    """The variants the generated kernels actually call (si_boys_fast: Horner/FMA Taylor sum on the
    transposed table + binary-power asymptote; the coul kernels' shared-rsqrt asymptote) against the
    reference's golden values.  Same table, same cell, same six terms: the stated bound is 5e-15 relative (23 ulp
    of 2.2e-16; measured worst case 2.4e-15 in the asymptotic branch at x = 200, where x^(-n-1/2) is built from
    ~10 multiplications), 200 times below the 1e-12 parity tolerance."""
    z = np.load(gu.GOLDEN / "boys_golden.npz")
    xs, vals = z["xs"], z["vals"]
    worst = 0.0
    for n in range(vals.shape[0]):
        got = engine.boys(n, xs, mode=mode)
        rel = np.abs(got - vals[n]) / np.abs(vals[n])
        worst = max(worst, float(rel.max()))
        np.testing.assert_allclose(got, vals[n], rtol=rtol, atol=0, err_msg=f"mode={mode} n={n}")
    assert worst > 0.0  # a different operation order than mode 'reference', not an alias of it


def _two_shell_block(engine, blk, sph, normalize):
    """A golden (a|b) block through the BATCHED entry points: a two-shell basis, off-diagonal block of the
    matrix -- the vectors hit the generated gc_kernel_* / the (a s0|b) 2c2e path, not the single-block entry."""
    from sympleints_b200.shells import Shell

    (ax, da, A), (bx, db, B) = blk["a"], blk["b"]
    La, Lb = blk["Ls"]
    basis = engine.basis([Shell(La, A, da, ax, normalized=True), Shell(Lb, B, db, bx, normalized=True)])
    size = (lambda l: 2 * l + 1) if sph else (lambda l: (l + 1) * (l + 2) // 2)
    na, nb = size(La), size(Lb)
    if blk["key"] == "coul":
        M = engine.coul(basis, blk["R"].reshape(1, 3), np.ones(1), sph=sph, normalize=normalize).cpu().numpy()
    else:
        M = engine.int2c2e(basis, sph=sph, normalize=normalize).cpu().numpy()
    assert np.array_equal(M[:na, na:], M[na:, :na].T)
    return M[:na, na:na + nb].ravel()


@pytest.mark.parametrize("variant", gu.variants())
def test_golden_coul_and_2c2e_blocks_through_batched_kernels(engine, variant):
    sph, normalize, blocks = gu.load_variant(variant)
    bad, n = [], 0
    for blk in blocks:
        if blk["key"] not in ("coul", "2c2e"):
            continue
        n += 1
        ok, rep = gu.block_error(_two_shell_block(engine, blk, sph, normalize), blk)
        if not ok:
            bad.append(rep)
    assert not bad, bad[:5]
    if variant == "sph_cgto_l4_laux5":
        assert n == 20 + 26   # 15 + 5 swapped coul classes, 21 + 5 swapped 2c2e classes


def test_coul_rejects_a_boys_table_that_is_too_small():
    """The batched coul path checks the Boys order against the caller's table (IndexError in the reference)."""
    import sympleints_b200
    from sympleints_b200.shells import Shell

    small = sympleints_b200.get_engine(0, 4, 5, boys_table=oboys.build_table(2)[1])   # rows for n <= 2
    sh = [Shell(2, [0.0, 0.0, 0.0], [1.0], [1.0]), Shell(1, [0.5, 0.0, 0.0], [1.0], [0.7])]
    with pytest.raises(sympleints_b200.SympleintsB200Error, match="Boys"):
        small.coul(small.basis(sh), np.zeros((1, 3)), np.ones(1))
    small.close()


def test_own_boys_table_is_accurate():
    """Without a caller-supplied table the library builds its own; it must agree with quadrature
    better than the reference's table does (the reference's depends on nmax at the 1e-9 level)."""
    import sympleints_b200

    eng = sympleints_b200.get_engine(0, 4, 5)
    tab = eng.boys_table()
    eng.close()
    ref = oboys.table()
    assert tab.shape[1] == 301 and tab.shape[0] >= 20
    nr = min(tab.shape[0], ref.shape[0])   # own table: orders up to 4 lmax for the Schwarz integrals, the reference's: nmax = 14
    np.testing.assert_allclose(tab[:nr], ref[:nr], rtol=5e-9)
    for n, i in ((0, 7), (13, 120), (19, 300)):
        np.testing.assert_allclose(tab[n, i], oboys.boys_quad(n, i * 0.1, err=1e-15), rtol=1e-12)


@pytest.mark.parametrize("variant", gu.variants())
def test_blocks_match_reference_golden(engine, variant):
    sph, normalize, blocks = gu.load_variant(variant)
    bad = []
    for blk in blocks:
        ok, rep = gu.block_error(_gpu_block(engine, blk, sph, normalize), blk)
        if not ok:
            bad.append(rep)
    assert not bad, bad[:5]


def test_gto_on_a_grid(engine):
    """Batched form of the gto key: all functions of a shell set on a grid of points, (nbf, npoints)."""
    from sympleints_b200 import synthetic

    shells, sites = synthetic.orbital_basis(natoms=3, seed=5)
    rng = np.random.default_rng(8)
    pts = rng.uniform(-2, 8, (257, 3))
    for sph, normalize in ((True, "cgto"), (True, "pgto"), (False, "cgto"), (True, "none")):
        V = engine.gto(engine.basis(shells), pts, sph=sph, normalize=normalize).cpu().numpy()
        row = 0
        for s in shells:
            n = 2 * s.L + 1 if sph else (s.L + 1) * (s.L + 2) // 2
            for g in (0, 100, 256):
                o = ints.gto(s.L, s.exps, s.coeffs, s.center, pts[g], sph=sph, normalize=normalize)
                assert np.abs(V[row:row + n, g] - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()
            row += n
        assert row == V.shape[0]


def _rand_shell(rng, L, K):
    ex = np.sort(10 ** rng.uniform(-1, 1.5, K))[::-1].copy()
    return ex, ints.norm_cgto_lmn(rng.uniform(0.2, 1, K), ex, L), rng.uniform(-1.5, 1.5, 3)


def test_all_90_3c2e_classes_vs_oracle(engine):
    """All (La>=Lb, Lc) classes of lmax=4 / lauxmax=5 plus swapped ones, float64 and longdouble oracle."""
    rng = np.random.default_rng(99)
    ld = lambda t: tuple(np.asarray(x, dtype=np.longdouble) for x in t)
    bad = []
    classes = [(a, b, c) for a in range(5) for b in range(a + 1) for c in range(6)]
    classes += [(b, a, c) for (a, b, c) in classes[::7] if a != b]
    for (La, Lb, Lc) in classes:
        K = (2, 1, 1) if La + Lb + Lc > 8 else (2, 2, 2)
        sa, sb, sc = _rand_shell(rng, La, K[0]), _rand_shell(rng, Lb, K[1]), _rand_shell(rng, Lc, K[2])
        o64 = ints.int3c2e_sph(La, Lb, Lc, *sa, *sb, *sc)
        o80 = ints.int3c2e_sph(La, Lb, Lc, *ld(sa), *ld(sb), *ld(sc))
        g = engine.eval_block("3c2e", (La, Lb, Lc), *sa, *sb, *sc)
        blk = dict(tag=f"3c2e_{La}{Lb}{Lc}", ref64=o64, ref80=o80)
        ok, rep = gu.block_error(g, blk)
        if not ok:
            bad.append(rep)
    assert not bad, bad[:5]


def test_error_conventions(engine):
    import sympleints_b200

    sa = (np.ones(1), np.ones(1), np.zeros(3))
    with pytest.raises(sympleints_b200.SympleintsB200Error, match="lmax"):
        engine.eval_block("ovlp", (5, 0), *sa, *sa)      # KeyError in the reference
    with pytest.raises(sympleints_b200.SympleintsB200Error):
        engine.eval_block("3c2e", (0, 0, 6), *sa, *sa, *sa)
    small = sympleints_b200.get_engine(0, 1, 1, boys_table=oboys.build_table(2)[1])
    with pytest.raises(sympleints_b200.SympleintsB200Error):  # IndexError from Boys in the reference
        small.eval_block("3c2e", (1, 1, 1), *sa, *sa, *sa)
    small.close()


def test_result_buffer_is_overwritten_in_place(engine):
    """PythonRenderer semantics: result[i] = ... (overwrite, templates/py_func.tpl:14-16)."""
    rng = np.random.default_rng(3)
    sa, sb = _rand_shell(rng, 2, 2), _rand_shell(rng, 1, 3)
    res = np.full(5 * 3, 123.0)
    out = engine.eval_block("ovlp", (2, 1), sa[0][:, None], sa[1][:, None], sa[2], sb[0][None, :], sb[1][None, :],
                            sb[2], R=np.zeros(3), result=res)
    assert out is res
    np.testing.assert_allclose(res, ints.ovlp(2, 1, *sa, *sb), rtol=1e-12, atol=1e-14)
