"""GPU tests of the batched drivers (the counterpart of sympleints/testing/intor.py) on small
synthetic systems, plus size-independent properties on larger ones."""
import numpy as np
import pytest

from oracle import ints
from sympleints_b200 import synthetic

pytestmark = pytest.mark.gpu


def _sizes(shells, sph):
    return [(2 * s.L + 1) if sph else (s.L + 1) * (s.L + 2) // 2 for s in shells]


def oracle_matrix(key, shells, sph=True, normalize="cgto", **kw):
    """intor_2c (testing/intor.py:18-76) driven by the oracle functions."""
    f, ncomp = ints.KEY_FUNCS[key]
    off = np.concatenate([[0], np.cumsum(_sizes(shells, sph))])
    M = np.zeros((ncomp, off[-1], off[-1]))
    for i, a in enumerate(shells):
        for j in range(i, len(shells)):
            b = shells[j]
            blk = f(a.L, b.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center, sph=sph,
                    normalize=normalize, **kw).reshape(ncomp, off[i + 1] - off[i], off[j + 1] - off[j])
            M[:, off[i]:off[i + 1], off[j]:off[j + 1]] = blk
            M[:, off[j]:off[j + 1], off[i]:off[i + 1]] = blk.transpose(0, 2, 1)
    return M


@pytest.fixture(scope="module")
def small(engine):
    shells, sites = synthetic.orbital_basis(natoms=3, seed=5)
    aux = synthetic.aux_basis(sites, seed=6)
    return shells, sites, aux, engine.basis(shells), engine.basis(aux)


@pytest.mark.parametrize("key", ["ovlp", "kin", "dpm", "qpm", "dqpm"])
@pytest.mark.parametrize("sph", [True, False])
def test_one_electron_matrices(engine, small, key, sph):
    shells, _, _, ob, _ = small
    R = np.array([0.3, -0.2, 0.5])
    g = engine.int1e(key, ob, sph=sph, R=R).cpu().numpy()
    o = oracle_matrix(key, shells, sph=sph, R=R)
    g = g.reshape(o.shape)
    assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()
    # off-diagonal shell blocks are mirrored exactly; diagonal blocks are computed in full
    assert np.abs(g - g.transpose(0, 2, 1)).max() <= 1e-15 * np.abs(o).max()
    h = engine.int1e_host(key, ob, sph=sph, R=R).reshape(o.shape)
    assert np.array_equal(h, g)


@pytest.mark.parametrize("normalize", ["cgto", "pgto"])
def test_one_electron_set_in_one_launch(engine, small, normalize):
    """si_int1e_set: ovlp, kin, dpm, qpm of all classes from the fused kernel k_onee (three concurrent launches: the
    classes by block size), against the oracle and bit-identical to the per-key calls (same
    kernel, operator mask)."""
    shells, _, _, ob, _ = small
    R = np.array([0.3, -0.2, 0.5])
    l0 = engine.launch_count()
    S = engine.int1e_set(ob, normalize=normalize, R=R)
    assert engine.launch_count() - l0 <= 3
    S = S.cpu().numpy()
    for key, o0, n in (("ovlp", 0, 1), ("kin", 1, 1), ("dpm", 2, 3), ("qpm", 5, 6)):
        o = oracle_matrix(key, shells, normalize=normalize, R=R)
        g = S[o0:o0 + n]
        assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max(), key
        single = engine.int1e(key, ob, normalize=normalize, R=R).cpu().numpy().reshape(o.shape)
        assert np.array_equal(single, g), key
    dq = engine.int1e("dqpm", ob, normalize=normalize, R=R).cpu().numpy()
    assert np.array_equal(dq, S[[5, 8, 10]])


def test_spherical_multipoles_about_the_primitive_pair_centres(engine, small):
    """multi_sph (section 8f-3): the nine R_lm operators, origin = P of every primitive pair, orders above La + Lb
    zeroed as in the reference (defs/multipole.py:60-138)."""
    shells, _, _, ob, _ = small
    g = engine.int1e("multi_sph", ob).cpu().numpy()
    o = oracle_matrix("multi_sph", shells)
    assert g.shape == o.shape == (9, ob.nbf_sph, ob.nbf_sph)
    assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()
    assert np.array_equal(engine.int1e_host("multi_sph", ob), g)
    with pytest.raises(Exception):
        engine.int1e("multi_sph", ob, sph=False)


def test_plumbing_config_ovlp_lmax1(engine):
    """BASELINE config 1: 20-atom s/p shell set, ovlp, against the intor_2c-style driver."""
    shells, _ = synthetic.plumbing_basis()
    assert len(shells) == 60 and sum(_sizes(shells, True)) == 100  # 20 x (s, s, p)
    b = engine.basis(shells)
    g = engine.int1e("ovlp", b).cpu().numpy()
    o = oracle_matrix("ovlp", shells)[0]
    assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()
    assert np.allclose(np.diag(g), 1.0, atol=1e-13)  # cGTO-normalised shells


def test_two_center_two_electron_metric(engine, small):
    _, _, aux, _, ab = small
    g = engine.int2c2e(ab).cpu().numpy()
    o = oracle_matrix("2c2e", aux)[0]
    assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()
    np.linalg.cholesky(g)  # the metric is positive definite


def test_nuclear_attraction_with_point_charges(engine, small):
    shells, sites, _, ob, _ = small
    xyz, q = synthetic.point_charges(sites, n=24, seed=3)
    g = engine.coul(ob, xyz, q).cpu().numpy()
    os_ = [oracle_matrix("coul", shells, R=rr)[0] for rr in xyz]
    o = sum(qq * oc for oc, qq in zip(os_, q))
    assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()
    h = engine.coul_host(ob, xyz, q)
    assert np.array_equal(h, g)
    # linearity in the charges
    g2 = engine.coul(ob, xyz, 2.0 * q).cpu().numpy()
    assert np.abs(g2 - 2 * g).max() <= 1e-13 * np.abs(g).max()
    # mirrored exactly
    assert np.abs(g - g.T).max() <= 1e-15 * np.abs(g).max()
    # every shell-pair block on its own scale (generated McMurchie-Davidson kernels, all 15 classes)
    off = np.concatenate([[0], np.cumsum(_sizes(shells, True))])
    for i in range(len(shells)):
        for j in range(len(shells)):
            blk = (slice(off[i], off[i + 1]), slice(off[j], off[j + 1]))
            tol = sum(abs(qq) * (1e-14 + 1e-12 * np.abs(oc[blk]).max()) for oc, qq in zip(os_, q))
            assert np.abs(g[blk] - o[blk]).max() <= tol, (i, j)


def test_nuclear_attraction_charge_split(engine, small):
    """Many charges: the generated kernels split them over several tuples per shell pair (partial matrices
    summed by k_sum_slices); the result must agree with the sum of unsplit evaluations of 25-charge chunks."""
    shells, sites, _, ob, _ = small
    xyz, q = synthetic.point_charges(sites, n=200, seed=9)
    g = engine.coul(ob, xyz, q).cpu().numpy()
    parts = sum(engine.coul(ob, xyz[i:i + 25], q[i:i + 25]).cpu().numpy() for i in range(0, 200, 25))
    assert np.abs(g - parts).max() <= 1e-13 * np.abs(g).max()
    assert np.abs(g - g.T).max() <= 1e-15 * np.abs(g).max()
    # a single charge
    g1 = engine.coul(ob, xyz[:1], q[:1]).cpu().numpy()
    o1 = q[0] * oracle_matrix("coul", shells, R=xyz[0])[0]
    assert np.abs(g1 - o1).max() <= 1e-14 + 1e-12 * np.abs(o1).max()


@pytest.mark.parametrize("sph,normalize", [(True, "pgto"), (False, "cgto"), (True, "none")])
def test_nuclear_attraction_other_variants(engine, small, sph, normalize):
    """pgto goes through the generated kernels, Cartesian / un-normalised output through the stage machine."""
    shells, sites, _, ob, _ = small
    xyz, q = synthetic.point_charges(sites, n=5, seed=4)
    g = engine.coul(ob, xyz, q, sph=sph, normalize=normalize).cpu().numpy()
    o = sum(qq * oracle_matrix("coul", shells, sph=sph, normalize=normalize, R=rr)[0] for rr, qq in zip(xyz, q))
    assert np.abs(g - o).max() <= 1e-14 + 1e-12 * np.abs(o).max()


def _oracle_3c_block(a, b, c):
    return ints.int3c2e_sph(a.L, b.L, c.L, a.exps, a.coeffs, a.center, b.exps, b.coeffs, b.center,
                            c.exps, c.coeffs, c.center)


def test_three_center_tensor_full_and_packed(engine, small):
    shells, _, aux, ob, ab = small
    T = engine.int3c2e(ob, ab, layout="full").cpu().numpy()
    Tp = engine.int3c2e(ob, ab, layout="packed").cpu().numpy()
    offo = np.concatenate([[0], np.cumsum(_sizes(shells, True))])
    offa = np.concatenate([[0], np.cumsum(_sizes(aux, True))])
    gmax = np.abs(T).max()
    rng = np.random.default_rng(1)
    # (a) every shell triple of a random subset against the oracle.  Blocks that vanish by
    # symmetry (same-centre shells) are compared on the absolute scale of the tensor.
    for _ in range(300):
        i, j, k = rng.integers(len(shells)), rng.integers(len(shells)), rng.integers(len(aux))
        o = _oracle_3c_block(shells[i], shells[j], aux[k])
        blk = T[offo[i]:offo[i + 1], offo[j]:offo[j + 1], offa[k]:offa[k + 1]].ravel()
        ld = lambda s: (s.L, s.exps.astype(np.longdouble), s.coeffs.astype(np.longdouble), s.center.astype(np.longdouble))
        tol = 1e-14 + 1e-12 * np.abs(o).max()
        if np.abs(blk - o).max() > tol:
            a, b, c = shells[i], shells[j], aux[k]
            o80 = ints.int3c2e_sph(a.L, b.L, c.L, *ld(a)[1:], *ld(b)[1:], *ld(c)[1:])
            noise = float(np.abs(o - o80).max())
            assert float(np.abs(blk - o80).max()) <= 10 * noise + tol + 1e-15 * gmax, (i, j, k)
    # (b) symmetry and packed == full
    assert np.abs(T - T.transpose(1, 0, 2)).max() <= 1e-15 * gmax
    row = 0
    for i in range(len(shells)):
        for j in range(i + 1):
            ni, nj = offo[i + 1] - offo[i], offo[j + 1] - offo[j]
            assert np.array_equal(Tp[row:row + ni * nj].reshape(ni, nj, -1), T[offo[i]:offo[i + 1], offo[j]:offo[j + 1]])
            row += ni * nj
    assert row == ob.packed_rows


def test_three_center_aux_sharding_is_exact(engine, small):
    """Sharding over contiguous aux-shell ranges (the multi-GPU partition) reproduces the
    unsharded tensor bit for bit."""
    _, _, aux, ob, ab = small
    T = engine.int3c2e(ob, ab, layout="packed").cpu().numpy()
    cuts = [0, 5, 6, 31, len(aux)]
    parts = [engine.int3c2e(ob, ab, shell_range=(cuts[i], cuts[i + 1]), layout="packed").cpu().numpy()
             for i in range(len(cuts) - 1)]
    assert np.array_equal(np.concatenate(parts, axis=1), T)
    h = engine.int3c2e_host(ob, ab, shell_range=(5, 31), layout="packed")
    assert np.array_equal(h, np.concatenate(parts[1:3], axis=1))


def test_translation_invariance_medium_system(engine):
    """Size-independent property at a larger size: shifting every centre leaves ovlp/kin/2c2e/3c2e
    unchanged (to rounding)."""
    shells, sites = synthetic.orbital_basis(natoms=8, seed=11)
    aux = synthetic.aux_basis(sites, seed=12)
    shift = np.array([0.7, -1.3, 2.1])
    from sympleints_b200.shells import Shell
    mv = lambda ss: [Shell(s.L, s.center + shift, s.coeffs, s.exps, normalized=True) for s in ss]
    b0, b1 = engine.basis(shells), engine.basis(mv(shells))
    a0, a1 = engine.basis(aux), engine.basis(mv(aux))
    for key in ("ovlp", "kin"):
        m0, m1 = engine.int1e(key, b0).cpu().numpy(), engine.int1e(key, b1).cpu().numpy()
        assert np.abs(m0 - m1).max() <= 1e-11 * np.abs(m0).max()
    t0 = engine.int3c2e(b0, a0, layout="packed")
    t1 = engine.int3c2e(b1, a1, layout="packed")
    assert float((t0 - t1).abs().max()) <= 1e-10 * float(t0.abs().max())
    m0, m1 = engine.int2c2e(a0).cpu().numpy(), engine.int2c2e(a1).cpu().numpy()
    assert np.abs(m0 - m1).max() <= 1e-10 * np.abs(m0).max()


def test_host_path_chunking_is_exact(engine, small, monkeypatch):
    """The host-buffer 3c2e path (row chunks for the packed layout, column chunks for the full layout,
    double-buffered D2H) reproduces the device-resident result bit for bit, whatever the chunk size."""
    _, _, _, ob, ab = small
    Tp = engine.int3c2e(ob, ab, layout="packed").cpu().numpy()
    Tf = engine.int3c2e(ob, ab, layout="full").cpu().numpy()
    for nbytes in ("40000", "700000", "100000000"):
        monkeypatch.setenv("SI_STAGE_BYTES", nbytes)
        assert np.array_equal(engine.int3c2e_host(ob, ab, layout="packed"), Tp), nbytes
        assert np.array_equal(engine.int3c2e_host(ob, ab, layout="full"), Tf), nbytes
        assert np.array_equal(engine.int3c2e_host(ob, ab, shell_range=(3, 40), layout="packed"),
                              engine.int3c2e(ob, ab, shell_range=(3, 40), layout="packed").cpu().numpy()), nbytes


def test_edge_cases_single_shell_and_reversed_order(engine):
    """Smallest inputs (one shell, one primitive, one charge, one aux shell) and a shell list in descending
    angular momentum (every off-diagonal pair is a "swapped" one, L(i) < L(j) never holds for i > j ... and
    the reverse list makes it hold for all)."""
    from sympleints_b200.shells import Shell

    g = Shell(4, [0.1, -0.2, 0.3], [1.0], [0.8])
    h = Shell(5, [0.4, 0.1, -0.3], [1.0], [1.1])
    bg, bh = engine.basis([g]), engine.basis([h])
    S = engine.int1e("ovlp", bg).cpu().numpy().reshape(9, 9)
    assert np.abs(S - np.eye(9)).max() <= 1e-13            # cgto-normalised spherical g shell
    V = engine.coul(bg, np.array([[0.5, 0.5, -0.5]]), np.array([-2.0])).cpu().numpy()
    oV = -2.0 * oracle_matrix("coul", [g], R=np.array([0.5, 0.5, -0.5]))[0]
    assert np.abs(V - oV).max() <= 1e-14 + 1e-12 * np.abs(oV).max()
    M = engine.int2c2e(bh).cpu().numpy()
    oM = oracle_matrix("2c2e", [h])[0]
    assert np.abs(M - oM).max() <= 1e-14 + 1e-12 * np.abs(oM).max()
    T = engine.int3c2e(bg, bh, layout="full").cpu().numpy()
    oT = _oracle_3c_block(g, g, h).reshape(9, 9, 11)
    assert np.abs(T - oT).max() <= 1e-14 + 1e-12 * np.abs(oT).max()
    Tp = engine.int3c2e(bg, bh, layout="packed").cpu().numpy()
    assert np.array_equal(Tp.reshape(9, 9, 11), T)
    # mixed list in both orders: the matrices are permutations of each other
    rng = np.random.default_rng(11)
    shells = [Shell(L, rng.uniform(-1, 1, 3), rng.uniform(0.2, 1, k), 10 ** rng.uniform(-0.5, 1, k))
              for L, k in ((0, 3), (1, 2), (2, 1), (3, 1), (4, 1))]
    rev = shells[::-1]
    a = engine.int1e("kin", engine.basis(shells)).cpu().numpy().reshape(25, 25)
    b = engine.int1e("kin", engine.basis(rev)).cpu().numpy().reshape(25, 25)
    off_a = np.concatenate([[0], np.cumsum(_sizes(shells, True))])
    off_b = np.concatenate([[0], np.cumsum(_sizes(rev, True))])
    perm = np.concatenate([np.arange(off_b[len(rev) - 1 - i], off_b[len(rev) - i]) for i in range(len(shells))])
    assert np.abs(a - b[np.ix_(perm, perm)]).max() <= 1e-13 * np.abs(a).max()
    aux = [h, Shell(2, [0.0, 0.3, 0.1], [1.0], [0.6])]
    ta = engine.int3c2e(engine.basis(shells), engine.basis(aux), layout="full").cpu().numpy()
    tb = engine.int3c2e(engine.basis(rev), engine.basis(aux), layout="full").cpu().numpy()
    assert np.abs(ta - tb[np.ix_(perm, perm)]).max() <= 1e-13 * np.abs(ta).max()
    va = engine.coul(engine.basis(shells), np.array([[0.2, 0.1, 0.4], [1.0, -1.0, 0.5]]), np.array([1.0, -0.5])).cpu().numpy()
    vb = engine.coul(engine.basis(rev), np.array([[0.2, 0.1, 0.4], [1.0, -1.0, 0.5]]), np.array([1.0, -0.5])).cpu().numpy()
    assert np.abs(va - vb[np.ix_(perm, perm)]).max() <= 1e-13 * np.abs(va).max()


def test_per_class_microbenchmark_and_row_selection(engine):
    """sympleints_b200.benchmark (counterpart of sympleints/benchmark.py:65-142): pure same-class batches through
    si_basis_select_rows; the selected blocks agree with the full evaluation and nothing else is written."""
    from sympleints_b200 import benchmark
    from sympleints_b200.shells import Shell

    rows = benchmark.run_key(engine, "3c2e_sph", 2, 2, nprims=2, npairs=16, naux=8, reps=2,
                             classes=[(2, 1, 2), (1, 1, 0), (2, 0, 1)])
    assert [(r["La"], r["Lb"], r["Lc"]) for r in rows] == [(2, 1, 2), (1, 1, 0), (2, 0, 1)]
    assert all(r["tot"] > 0 and r["tuples"] == 128 and r["tflops"] > 0 for r in rows)
    rows = benchmark.run_key(engine, "ovlp", 2, 2, nprims=3, npairs=8, reps=2, classes=[(2, 1, None)])
    assert rows[0]["integrals"] == 8 * 5 * 3 and rows[0]["tflops"] is None
    assert "La Lb Lc" in benchmark.format_rows("ovlp", rows)
    # selection semantics
    rng = np.random.default_rng(4)
    shells = [Shell(L, rng.uniform(-2, 2, 3), rng.uniform(0.2, 1, 2), 10 ** rng.uniform(-0.5, 1, 2)) for L in (1, 1, 0, 2)]
    aux = [Shell(1, rng.uniform(-2, 2, 3), [1.0], [0.9]), Shell(2, rng.uniform(-2, 2, 3), [1.0], [0.5])]
    b, ab = engine.basis(shells), engine.basis(aux)
    full = engine.int3c2e(b, ab, layout="full").cpu().numpy()
    S = engine.int1e("kin", b).cpu().numpy()
    b.select_rows(3, 4, skip_diagonal=True)          # pairs (d-shell, each of the three earlier shells)
    assert b.packed_rows == 5 * (3 + 3 + 1)
    T = engine.int3c2e(b, ab, layout="packed").cpu().numpy()
    assert np.array_equal(T.reshape(7 * 5, 8)[:15].reshape(5, 3, 8), full[7:12, 0:3])
    import torch
    out = torch.full((12, 12), 7.0, dtype=torch.float64, device=engine.tdev)
    engine.int1e("kin", b, out=out)
    o = out.cpu().numpy()
    assert np.array_equal(o[7:12, 0:7], S[7:12, 0:7]) and np.array_equal(o[0:7, 7:12], S[0:7, 7:12])
    assert (o[0:7, 0:7] == 7.0).all() and (o[7:12, 7:12] == 7.0).all()
    b.select_rows()                                   # back to all pairs
    assert np.array_equal(engine.int1e("kin", b).cpu().numpy(), S)


def test_entry_points_from_two_host_threads(engine, oracle_table):
    """The C ABI from several host threads (the reference's functions are pure / nogil): two contexts
    used concurrently, and ONE context shared by two threads that pass different CUDA streams (calls on a
    context are serialised by its mutex and ordered on the device through its completion event)."""
    import threading

    import torch

    import sympleints_b200

    shells_a, sites = synthetic.orbital_basis(natoms=3, seed=21)
    shells_b, _ = synthetic.orbital_basis(natoms=2, seed=22)
    aux = synthetic.aux_basis(sites, seed=23)
    ref = {}
    for name, sh in (("a", shells_a), ("b", shells_b)):
        b = engine.basis(sh)
        ref[name] = (engine.int1e("kin", b).cpu().numpy(), engine.int1e("qpm", b, R=(0.1, 0.2, 0.3)).cpu().numpy())
    ref3 = engine.int3c2e(engine.basis(shells_a), engine.basis(aux), layout="packed").cpu().numpy()
    errors = []

    def worker(eng, name, sh, with_3c):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                b = eng.basis(sh)
                ab = eng.basis(aux) if with_3c else None
                for _ in range(10):
                    k = eng.int1e("kin", b)
                    q = eng.int1e("qpm", b, R=(0.1, 0.2, 0.3))
                    t = eng.int3c2e(b, ab, layout="packed") if with_3c else None
                    st.synchronize()
                    if not (np.array_equal(k.cpu().numpy(), ref[name][0]) and np.array_equal(q.cpu().numpy(), ref[name][1])):
                        errors.append((name, "2-index result differs"))
                    if with_3c and not np.array_equal(t.cpu().numpy(), ref3):
                        errors.append((name, "3-index result differs"))
        except Exception as e:   # noqa: BLE001
            errors.append((name, repr(e)))

    # (1) two contexts, one thread each
    e1 = sympleints_b200.get_engine(0, 4, 5, boys_table=oracle_table)
    e2 = sympleints_b200.get_engine(0, 4, 5, boys_table=oracle_table)
    th = [threading.Thread(target=worker, args=(e1, "a", shells_a, True)),
          threading.Thread(target=worker, args=(e2, "b", shells_b, False))]
    [t.start() for t in th]
    [t.join() for t in th]
    e1.close()
    e2.close()
    assert not errors, errors[:3]
    # (2) one context shared by two threads on different streams
    th = [threading.Thread(target=worker, args=(engine, "a", shells_a, True)),
          threading.Thread(target=worker, args=(engine, "b", shells_b, False))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors[:3]


def test_opt_in_primitive_prescreening(engine, small):
    """si_set_screening: off by default (bit-identical), a tiny threshold changes nothing measurable, a generous one
    drops only contributions of that size (pPRE2 estimate of graphs/templates/int_3c2e_func.tpl:46-54)."""
    _, _, _, ob, ab = small
    ref = engine.int3c2e(ob, ab, layout="packed").cpu().numpy()
    try:
        engine.set_screening(1e-30)
        assert np.abs(engine.int3c2e(ob, ab, layout="packed").cpu().numpy() - ref).max() <= 1e-25
        engine.set_screening(1e-9)
        dev = np.abs(engine.int3c2e(ob, ab, layout="packed").cpu().numpy() - ref).max()
        assert dev <= 1e-6 * np.abs(ref).max()
    finally:
        engine.set_screening(0.0)
    assert np.array_equal(engine.int3c2e(ob, ab, layout="packed").cpu().numpy(), ref)


def test_density_fitting_consumers(engine, small):
    """B = (ab|P) L^-T with (P|Q) = L L^T reproduces (ab|P)(P|Q)^-1(Q|cd) (sympleints_b200/df.py)."""
    from sympleints_b200 import df

    _, _, _, ob, ab = small
    T = engine.int3c2e(ob, ab, layout="packed").cpu().numpy()
    M = engine.int2c2e(ab).cpu().numpy()
    B = df.fitted_tensor(engine, ob, ab).cpu().numpy()
    want = T @ np.linalg.solve(M, T.T)
    got = B @ B.T
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
