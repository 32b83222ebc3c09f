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
GENDIR = CSRC / "generated"
OBJDIR = PKG / "lib" / "obj"
SOURCES = [CSRC / "si_lib.cu", CSRC / "si_program.cpp"]
HEADERS = [CSRC / "si_machine.h", CSRC / "si_program.h", PKG.parent / "include" / "sympleints_b200.h"]
LIB_HEADERS = [CSRC / "si_onee.cuh", CSRC / "si_gto.cuh", CSRC / "si_schwarz.cuh"]   # included by si_lib.cu only
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


def generate_sources():
    """Run the code generator (class-specialised 3c2e kernels) into csrc/generated/."""
    from .codegen import gen3c2e, gencoul, genonee

    genonee.write_sources(str(GENDIR))   # header only (included by si_onee.cuh)
    return [Path(f) for f in gen3c2e.write_sources(str(GENDIR))] + \
           [Path(f) for f in gencoul.write_sources(str(GENDIR))]


def _compile_obj(src, obj, verbose):
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-c", "-o", str(obj), str(src)] + (["-Xptxas", "-v"] if verbose else [])
    log = obj.with_suffix(".log")
    with open(log, "w") as fh:
        subprocess.check_call(cmd, stdout=fh, stderr=subprocess.STDOUT)
    return obj


def build_library(force=False, verbose=False, jobs=None):
    gen = generate_sources()
    deps = SOURCES + HEADERS + LIB_HEADERS + [CSRC / "si_g3.h", PKG / "codegen" / "gen3c2e.py", PKG / "codegen" / "gencoul.py", PKG / "codegen" / "genonee.py",
                                GENDIR / "g3_table.inc", GENDIR / "gc_table.inc"] + gen
    for tj in (PKG / "codegen" / "tuning.json", PKG / "codegen" / "tuning_coul.json"):
        if tj.exists():
            deps.append(tj)
    if not force and LIB.exists() and all(p.stat().st_mtime <= LIB.stat().st_mtime for p in deps):
        return LIB
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    import concurrent.futures as cf
    import os

    todo = []
    objs = []
    hdr_time = max(p.stat().st_mtime for p in HEADERS + [CSRC / "si_g3.h"])
    for src in gen + SOURCES:
        obj = OBJDIR / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_time) or \
                (src.name == "si_lib.cu" and obj.stat().st_mtime < max([(GENDIR / "g3_table.inc").stat().st_mtime,
                                                                         (GENDIR / "gc_table.inc").stat().st_mtime,
                                                                         (GENDIR / "onee_tables.inc").stat().st_mtime] +
                                                                        [h.stat().st_mtime for h in LIB_HEADERS])):
            todo.append((src, obj))
    jobs = jobs or max(1, min(8, (os.cpu_count() or 2)))
    with cf.ThreadPoolExecutor(jobs) as ex:
        list(ex.map(lambda so: _compile_obj(so[0], so[1], verbose), todo))
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           "-o", str(LIB)] + [str(o) for o in objs]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
