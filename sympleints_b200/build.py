"""Build the CUDA library in-tree (sm_100a only).

``python -m sympleints_b200.build`` or ``build_library()``.  The resulting
``sympleints_b200/lib/libsympleints_b200.so`` is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libsympleints_b200.so"
SOURCES = [CSRC / "si_lib.cu", CSRC / "si_program.cpp"]
HEADERS = [CSRC / "si_machine.h", CSRC / "si_program.h", PKG.parent / "include" / "sympleints_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(p).exists():
        raise RuntimeError("nvcc not found: the sympleints_b200 CUDA library cannot be built")
    return p


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + [str(s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
