"""Loading of tests/golden fixtures + the parity criterion used everywhere.

Criterion (per block; BASELINE.json: 1e-12 relative / 1e-14 absolute in FP64):
    max|x - ref| <= ATOL + RTOL * max|ref|,   ATOL = 1e-14, RTOL = 1e-12
against the reference's float64 result, OR -- for blocks where the reference's own FP64
rounding noise exceeds that bound (high-l Boys classes, SURVEY.md section 7 hard part 2) --
    max|x - ref80| <= NOISE_FACTOR * max|ref64 - ref80| + ATOL + RTOL * max|ref|
i.e. we are no further from the reference's extended-precision evaluation than a small
multiple of the reference's own float64 evaluation is.
"""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
ATOL = 1e-14
RTOL = 1e-12
NOISE_FACTOR = 10.0


def variants():
    return sorted(p.stem[len("blocks_"):] for p in GOLDEN.glob("blocks_*.npz"))


def load_variant(v):
    z = np.load(GOLDEN / f"blocks_{v}.npz")
    sph = bool(int(z["_sph"]))
    normalize = str(z["_normalize"])
    blocks = []
    for tag in z["_meta"]:
        tag = str(tag)
        key, ls = tag.rsplit("_", 1)
        Ls = tuple(int(c) for c in ls)

        def split(arr):
            K = (arr.size - 3) // 2
            return arr[:K], arr[K:2 * K], arr[2 * K:]

        blk = dict(tag=tag, key=key, Ls=Ls, a=split(z[tag + "_a"]), b=split(z[tag + "_b"]) if tag + "_b" in z else None,
                   c=split(z[tag + "_c"]) if tag + "_c" in z else None, R=z[tag + "_R"],
                   ref64=z[tag + "_ref64"],
                   ref80=z[tag + "_ref80hi"].astype(np.longdouble) + z[tag + "_ref80lo"].astype(np.longdouble))
        blocks.append(blk)
    return sph, normalize, blocks


def block_error(val, blk):
    """Returns (ok, report dict)."""
    ref64, ref80 = blk["ref64"], blk["ref80"]
    m = float(np.abs(ref64).max())
    tol = ATOL + RTOL * m
    e64 = float(np.abs(val - ref64).max())
    e80 = float(np.abs(val.astype(np.longdouble) - ref80).max())
    noise = float(np.abs(ref64.astype(np.longdouble) - ref80).max())
    ok = (e64 <= tol) or (e80 <= NOISE_FACTOR * noise + tol)
    return ok, dict(tag=blk["tag"], max=m, e64=e64, e80=e80, noise=noise, tol=tol)


def oracle_block(blk, sph, normalize, dtype=np.float64):
    from oracle import ints

    key, Ls = blk["key"], blk["Ls"]
    c = lambda t: tuple(np.asarray(x, dtype=dtype) for x in t)
    a, b = c(blk["a"]), (c(blk["b"]) if blk["b"] is not None else None)
    if key == "3c2e":
        return ints.int3c2e_sph(*Ls, *a, *b, *c(blk["c"]), normalize=normalize)
    if key == "gto":   # one-index function: f(ax, da, A, R)
        return ints.gto(Ls[0], *a, np.asarray(blk["R"], dtype=dtype), sph=sph, normalize=normalize)
    if key == "prefactor":   # R carries rP (module-level name in the reference's generated code)
        return ints.prefactor(Ls[0], Ls[1], *a, *b, R=np.asarray(blk["R"], dtype=dtype), sph=sph, normalize=normalize)
    if key == "multi_sph":   # origin = centre of every primitive pair, R unused
        return ints.multi_sph(Ls[0], Ls[1], *a, *b, sph=sph, normalize=normalize)
    f = ints.KEY_FUNCS[key][0]
    return f(Ls[0], Ls[1], *a, *b, R=np.asarray(blk["R"], dtype=dtype), sph=sph, normalize=normalize)
