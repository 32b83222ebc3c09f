#!/usr/bin/env python
"""Run the UNMODIFIED reference generator (sympleints.main.run with its own
PythonRenderer) from /root/reference and write its generated numpy modules to
oracle/_ref/ref_python/.  These files are the reference's own output, executed
here as the primary oracle; they are git-ignored but
travel to the GPU box with the gpurun snapshot.

Usage (in the build container only -- /root/reference does not exist on the GPU box):
    python oracle/refgen/generate_ref.py --keys ovlp kin dpm qpm coul 2c2e --lmax 4 --lauxmax 5
    python oracle/refgen/generate_ref.py --keys 3c2e_sph --lmax 4 --lauxmax 5 --tag full
"""
import argparse
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path(os.environ.get("SYMPLEINTS_REFERENCE", "/root/reference"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", nargs="+", required=True)
    ap.add_argument("--lmax", type=int, default=4)
    ap.add_argument("--lauxmax", type=int, default=5)
    ap.add_argument("--cart", action="store_true", help="Cartesian instead of --sph")
    ap.add_argument("--normalize", default="cgto", choices=["pgto", "cgto", "none"])
    ap.add_argument("--prefix", default=None)
    ap.add_argument("--workers", type=int, default=6)
    args = ap.parse_args()

    os.environ["SYMPLEINTS_REF_WORKERS"] = str(args.workers)
    sys.path.insert(0, str(HERE / "stubs"))
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REPO))

    from sympleints.main import run, normalization_map
    from sympleints.PythonRenderer import PythonRenderer

    prefix = args.prefix
    if prefix is None:
        prefix = "ref_%s_%s_l%d_laux%d" % (
            "cart" if args.cart else "sph", args.normalize, args.lmax, args.lauxmax)
    out_dir = REPO / "oracle" / "_ref"
    out_dir.mkdir(exist_ok=True)
    res = run(
        l_max=args.lmax,
        l_aux_max=args.lauxmax,
        sph=not args.cart,
        normalization=normalization_map[args.normalize],
        out_dir=str(out_dir),
        prefix=prefix,
        keys=args.keys,
        boys_func="oracle.boys",
        renderers=[PythonRenderer()],
    )
    print(res)


if __name__ == "__main__":
    main()
