"""CPU tests: the oracle restatements against the reference's golden vectors."""
import numpy as np
import pytest

import golden_util as gu
from oracle import boys as oboys
from oracle import ints


def test_boys_table_and_values_bit_exact():
    z = np.load(gu.GOLDEN / "boys_golden.npz")
    _, table = oboys.build_table(14)
    assert np.array_equal(table, z["table"]), "oracle Boys table differs from the reference's"
    b = oboys.get_boys_func(14, table=table)
    xs, vals = z["xs"], z["vals"]
    for n in range(vals.shape[0]):
        got = np.array([b(n, float(x)) for x in xs])
        assert np.array_equal(got, vals[n]), f"scalar Boys differs for n={n}"
        # numpy's vector pow may differ from the scalar pow in the last bit
        np.testing.assert_allclose(b.vec(n, xs), vals[n], rtol=2e-15, atol=0)
        assert np.array_equal(b(n, xs.reshape(4, -1)), vals[n].reshape(4, -1))


def test_boys_reference_accuracy_claim():
    """tests/test_boys/test_boys.py of the reference: 2e-10 against quadrature."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(0, 25, 200), [0.0], np.linspace(30, 35, 10)])
    for n in (0, 3, 10):
        ref = np.array([oboys.boys_quad(n, x) for x in xs])
        np.testing.assert_allclose(oboys.boys_vec(n, xs), ref, atol=2e-10)


def test_cart2sph_ordering():
    # l=1 rows are x, z, y  
    C = ints.cart2sph(1)
    assert np.allclose(C, [[1, 0, 0], [0, 0, 1], [0, 1, 0]])
    C2 = ints.cart2sph(2)
    assert [int((row != 0).sum()) for row in C2] == [2, 1, 3, 1, 1]
    assert [int((ints.cart2sph(l) != 0).sum()) for l in range(6)] == [1, 3, 8, 16, 28, 46]


@pytest.mark.parametrize("variant", gu.variants())
def test_oracle_matches_reference_blocks(variant):
    sph, normalize, blocks = gu.load_variant(variant)
    bad = []
    for blk in blocks:
        val = gu.oracle_block(blk, sph, normalize)
        ok, rep = gu.block_error(val, blk)
        if not ok:
            bad.append(rep)
    assert not bad, bad[:5]


def test_oracle_longdouble_tracks_reference_longdouble():
    """The oracle evaluated in extended precision agrees with the reference's generated code
    evaluated in extended precision far below the FP64 noise level."""
    sph, normalize, blocks = gu.load_variant("sph_cgto_l4_laux5")
    for blk in blocks[::7]:
        val = gu.oracle_block(blk, sph, normalize, dtype=np.longdouble)
        m = float(np.abs(blk["ref80"]).max())
        noise = float(np.abs(blk["ref64"] - blk["ref80"]).max())
        # Boys values and the printed numeric constants are float64 in both, hence a residual
        assert float(np.abs(val - blk["ref80"]).max()) <= max(2 * noise, 1e-13 * m) + 1e-17, blk["tag"]
