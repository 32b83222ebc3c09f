"""CPU tests: the stage programs built by csrc/si_program.cpp, executed by the TEST-ONLY host
emulator (tests/emu), against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

import emu_lib
import golden_util as gu
from oracle import ints


def _emu(blk, sph, normalize):
    key, Ls = blk["key"], blk["Ls"]
    Lc = Ls[2] if key == "3c2e" else 0
    return emu_lib.emu_eval(key, Ls[0], Ls[1], Lc, blk["a"], blk["b"], blk["c"], R=blk["R"], sph=sph,
                            normalize=normalize)


@pytest.mark.parametrize("variant", gu.variants())
def test_emulated_programs_match_reference_blocks(variant):
    sph, normalize, blocks = gu.load_variant(variant)
    bad = []
    for blk in blocks:
        if blk["key"] in ("gto", "prefactor", "multi_sph"):
            continue   # closed-form keys with kernels of their own (csrc/si_gto.cuh), not stage programs
        coefs_scaled = blk
        if normalize == "pgto":
            # the C ABI folds the radial pgto norm into the coefficients (si_basis_create);
            # the emulator takes pre-scaled coefficients
            def scale(t, L):
                ex, co, cen = t
                return ex, co * np.sqrt(2 ** (2 * L + 1.5) * ex ** (L + 1.5) / np.pi**1.5), cen
            Ls = blk["Ls"]
            coefs_scaled = dict(blk)
            coefs_scaled["a"] = scale(blk["a"], Ls[0])
            coefs_scaled["b"] = scale(blk["b"], Ls[1])
            if blk["c"] is not None:
                coefs_scaled["c"] = scale(blk["c"], Ls[2])
        val = _emu(coefs_scaled, sph, normalize)
        ok, rep = gu.block_error(val, blk)
        if not ok:
            bad.append(rep)
    assert not bad, bad[:5]


def test_c2s_matrices_match_oracle():
    lib = emu_lib.lib()
    import ctypes
    for l in range(8):
        out = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2))
        lib.siemu_cart2sph(l, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert np.abs(out - ints.cart2sph(l)).max() < 5e-16


@pytest.mark.parametrize("Ls", [(0, 0, 0), (1, 1, 2), (2, 1, 4), (3, 3, 3), (4, 2, 5), (4, 4, 5), (2, 4, 3), (0, 3, 5)])
def test_3c2e_all_stage_orders_vs_oracle(Ls):
    rng = np.random.default_rng(sum(Ls) * 7 + Ls[0])
    def shell(L, K):
        ex = 10 ** rng.uniform(-1, 1.5, K)
        return ex, ints.norm_cgto_lmn(rng.uniform(0.2, 1, K), ex, L), rng.uniform(-1.5, 1.5, 3)
    sa, sb, sc = shell(Ls[0], 2), shell(Ls[1], 3), shell(Ls[2], 2)
    o = ints.int3c2e_sph(*Ls, *sa, *sb, *sc)
    e = emu_lib.emu_eval("3c2e", *Ls, sa, sb, sc)
    ld = lambda t: tuple(np.asarray(x, dtype=np.longdouble) for x in t)
    o80 = ints.int3c2e_sph(*Ls, *ld(sa), *ld(sb), *ld(sc))
    m = np.abs(o).max()
    noise = float(np.abs(o - o80).max())
    assert np.abs(e - o).max() <= 1e-14 + 1e-12 * m or float(np.abs(e - o80).max()) <= 10 * noise + 1e-12 * m


def test_program_resources_fit_shared_memory():
    """Every class of the benchmark configuration must fit one slot into 227 KB of shared memory."""
    worst = 0
    for La in range(5):
        for Lb in range(La + 1):
            for Lc in range(6):
                st = emu_lib.program_stats("3c2e", La, Lb, Lc)
                worst = max(worst, st["w_size"])
                assert (st["w_size"] + 64 + st["nconst"]) * 8 <= 227 * 1024
    for La in range(6):
        for Lb in range(La + 1):
            assert emu_lib.program_stats("2c2e", La, Lb)["w_size"] * 8 < 227 * 1024
    assert worst < 8192
