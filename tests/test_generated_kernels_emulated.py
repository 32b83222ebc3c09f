"""CPU tests: the GENERATED class-specialised 3c2e kernels (codegen/gen3c2e.py), executed by the
TEST-ONLY warp emulator (32 host threads, tests/emu/g3_emu.cpp), against the oracle."""
import numpy as np
import pytest

import emu_lib
from oracle import ints


def _shell(rng, L, K):
    ex = 10 ** rng.uniform(-1, 1.2, K)
    return ex, ints.norm_cgto_lmn(rng.uniform(0.2, 1, K), ex, L), rng.uniform(-1.5, 1.5, 3)


@pytest.mark.parametrize("Ls", emu_lib.G3_EMU_CLASSES)
def test_generated_class_vs_oracle(Ls):
    La, Lb, Lc = Ls
    rng = np.random.default_rng(100 * La + 10 * Lb + Lc)
    sa, sb = _shell(rng, La, 2), _shell(rng, Lb, 3)
    # several aux shells with different contraction lengths: exercises tuples-per-warp packing,
    # the padded primitive loop and the "repeat last valid tuple" path
    auxs = [_shell(rng, Lc, k) for k in (1, 2, 1, 3, 1, 1, 2)]
    g = emu_lib.g3_eval(La, Lb, Lc, sa, sb, auxs)
    # the same through an explicit tuple table (reversed order, padded groups): bit-identical blocks
    gt = emu_lib.g3_eval(La, Lb, Lc, sa, sb, auxs, tuple_table=True)
    assert np.array_equal(g, gt)
    # two pairs interleaved in the table (tuples of different pairs inside one warp)
    g2 = emu_lib.g3_eval(La, Lb, Lc, sa, sb, auxs, tuple_table=2)
    assert np.array_equal(g2[:len(auxs)], g)
    sa_shift = (sa[0], sa[1], np.asarray(sa[2]) + np.array([0.5, -0.25, 0.125]))
    for k, sc in enumerate(auxs):
        o = ints.int3c2e_sph(La, Lb, Lc, *sa_shift, *sb, *sc)
        assert np.abs(o - g2[len(auxs) + k].ravel()).max() <= 1e-14 + 1e-11 * np.abs(o).max(), (Ls, k, "second pair")
    ld = lambda t: tuple(np.asarray(x, dtype=np.longdouble) for x in t)
    for k, sc in enumerate(auxs):
        o = ints.int3c2e_sph(La, Lb, Lc, *sa, *sb, *sc)
        m = np.abs(o).max()
        err = np.abs(o - g[k].ravel()).max()
        if err > 1e-14 + 1e-12 * m:
            o80 = ints.int3c2e_sph(La, Lb, Lc, *ld(sa), *ld(sb), *ld(sc))
            noise = float(np.abs(o - o80).max())
            assert float(np.abs(g[k].ravel() - o80).max()) <= 10 * noise + 1e-14 + 1e-12 * m, (Ls, k)


def test_generator_covers_all_classes_and_fits_shared_memory():
    from sympleints_b200.codegen import gen3c2e

    info = gen3c2e.class_info(4, 5)
    assert len(info) == 96   # 90 3c2e classes + (5,0|Lc) used by the 2c2e metric
    for cls, g in info.items():
        assert g.WSZ * g.TPW * 8 <= 227 * 1024, cls
        assert g.EPT * g.TPW == 32


# ---- generated nuclear-attraction kernels (codegen/gencoul.py, McMurchie-Davidson) -----------------
@pytest.mark.parametrize("Ls", emu_lib.GC_EMU_CLASSES)
def test_generated_coul_class_vs_oracle(Ls):
    La, Lb = Ls
    rng = np.random.default_rng(10 * La + Lb)
    # several shell pairs with different contraction lengths share a warp (pairs-per-warp packing, the
    # padded primitive-pair loop, the "repeat last valid pair" path); several charges per pair
    pairs = [(_shell(rng, La, int(ka)), _shell(rng, Lb, int(kb))) for ka, kb in ((1, 1), (2, 3), (3, 1), (1, 2), (2, 2))]
    xyz = rng.uniform(-2, 2, (5, 3))
    q = rng.uniform(-3, 3, 5)
    g = emu_lib.gc_eval(La, Lb, pairs, xyz, q)
    for k, (sa, sb) in enumerate(pairs):
        o = sum(qq * ints.coul(La, Lb, *sa, *sb, R=rr) for rr, qq in zip(xyz, q)).reshape(g[k].shape)
        assert np.abs(o - g[k]).max() <= 1e-14 + 1e-12 * np.abs(o).max(), (Ls, k)
    # charges split over 3 tuples per pair (partial matrices summed afterwards)
    g3 = emu_lib.gc_eval(La, Lb, pairs, xyz, q, nsplit=3)
    assert np.abs(g3 - g).max() <= 1e-13 * np.abs(g).max()


def test_generated_coul_vs_golden_blocks():
    """Per-block parity criterion against the reference's own output (tests/golden), incl. the (g|g)
    block with very unequal exponents that a VRR+HRR evaluation fails."""
    import golden_util as gu

    sph, normalize, blocks = gu.load_variant("sph_cgto_l4_laux5")
    n = 0
    for blk in blocks:
        if blk["key"] != "coul":
            continue
        La, Lb = blk["Ls"]
        R = blk["R"].reshape(1, 3)
        if La >= Lb:
            got = emu_lib.gc_eval(La, Lb, [(blk["a"], blk["b"])], R, np.ones(1))[0]
        else:
            got = emu_lib.gc_eval(Lb, La, [(blk["b"], blk["a"])], R, np.ones(1))[0].T
        ok, rep = gu.block_error(got.reshape(blk["ref64"].shape), blk)
        assert ok, rep
        n += 1
    assert n >= 15
