"""ctypes binding of the C ABI declared in include/sympleints_b200.h.

There is NO CPU fallback: if the shared library is missing or no CUDA device is
usable, loading / context creation raises.
"""
import ctypes
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_void_p
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("SYMPLEINTS_B200_LIB", PKG / "lib" / "libsympleints_b200.so"))

KEYS = {"ovlp": 0, "kin": 1, "dpm": 2, "qpm": 3, "dqpm": 4, "coul": 5, "2c2e": 6, "3c2e_sph": 7, "3c2e": 7, "gto": 8,
        "prefactor": 9, "multi_sph": 10}
NCOMP = {"ovlp": 1, "kin": 1, "dpm": 3, "qpm": 6, "dqpm": 3, "coul": 1, "2c2e": 1, "3c2e_sph": 1, "3c2e": 1, "gto": 1,
         "prefactor": 1, "multi_sph": 9}
NORMALIZE = {"none": 0, "cgto": 1, "pgto": 2}
LAYOUT = {"full": 0, "packed": 1}

c_double_p = POINTER(c_double)
c_int_p = POINTER(c_int)

# every symbol include/sympleints_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "si_strerror": (c_char_p, [c_int]),
    "si_last_error": (c_char_p, []),
    "si_init": (c_int, [c_int, c_int, c_int, c_double_p, c_int, c_int, POINTER(c_void_p)]),
    "si_finalize": (c_int, [c_void_p]),
    "si_boys_table": (c_int, [c_void_p, c_double_p, c_int_p, c_int_p]),
    "si_boys_eval": (c_int, [c_void_p, c_int, c_double_p, c_size_t, c_double_p]),
    "si_boys_eval_mode": (c_int, [c_void_p, c_int, c_int, c_double_p, c_size_t, c_double_p]),
    "si_basis_create": (c_int, [c_void_p, c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_double_p, POINTER(c_void_p)]),
    "si_basis_destroy": (c_int, [c_void_p]),
    "si_basis_nbf": (c_int, [c_void_p, c_int]),
    "si_basis_nshell": (c_int, [c_void_p]),
    "si_basis_packed_rows": (c_longlong, [c_void_p]),
    "si_basis_select_rows": (c_int, [c_void_p, c_int, c_int, c_int]),
    "si_basis_selected_rows": (c_longlong, [c_void_p]),
    "si_int1e": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double_p, c_void_p, c_void_p]),
    "si_int1e_set": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double_p, c_void_p, c_void_p]),
    "si_coul": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double_p, c_double_p, c_void_p, c_void_p]),
    "si_int2c2e": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "si_int3c2e": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "si_gto": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_double_p, c_void_p, c_void_p]),
    "si_gto_h": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_double_p, c_double_p]),
    "si_int1e_h": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double_p, c_double_p]),
    "si_coul_h": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double_p, c_double_p, c_double_p]),
    "si_int2c2e_h": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double_p]),
    "si_int3c2e_h": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double_p, c_size_t]),
    "si_eval_block": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_double_p, c_double_p, c_double_p,
                              c_int, c_double_p, c_double_p, c_double_p,
                              c_int, c_double_p, c_double_p, c_double_p,
                              c_double_p, c_double_p]),
    "si_schwarz": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "si_set_screening": (c_int, [c_void_p, c_double]),
    "si_host_breakdown": (c_int, [c_void_p, c_double_p]),
    "si_fp64_peak": (c_int, [c_void_p, c_double, c_double_p]),
    "si_launch_count": (c_longlong, [c_void_p]),
    "si_class_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_longlong)]),
}

_LIB = None


class SympleintsB200Error(RuntimeError):
    pass


def load():
    """Load the shared library and bind every declared symbol (raises if absent)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise SympleintsB200Error(
            f"{LIB_PATH} is missing: build it with `python -m sympleints_b200.build` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(status):
    if status != 0:
        lib = load()
        msg = lib.si_strerror(status).decode()
        detail = lib.si_last_error().decode()
        raise SympleintsB200Error(f"sympleints_b200 error {status}: {msg}" + (f" ({detail})" if detail else ""))
