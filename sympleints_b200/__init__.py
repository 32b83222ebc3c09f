"""sympleints_b200 -- B200-native evaluation engine for sympleints' shell integrals.

The package holds only what the hot path needs: the CUDA kernels and C ABI
(``csrc/``, ``include/sympleints_b200.h``), the ctypes binding (``lib``), the
host shell-batching driver (``engine``), the ``CudaRenderer`` plugin mirroring
the reference's Renderer surface, and the multi-GPU sharding helper (``dist``).
"""
from .lib import KEYS, NCOMP, SympleintsB200Error  # noqa: F401
from .shells import Shell, norm_cgto_lmn  # noqa: F401


def get_engine(*args, **kwargs):
    from .engine import Engine

    return Engine(*args, **kwargs)
