"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol and
fails loudly (no CPU fallback) when no CUDA device is present."""
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (REPO / "include" / "sympleints_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(si_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from sympleints_b200 import build, lib

    build.build_library()
    L = lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sympleints_b200.h but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(lib.SIGNATURES) == declared


def test_strerror_and_argument_checks_without_gpu():
    from sympleints_b200 import lib

    L = lib.load()
    assert L.si_strerror(0) == b"ok"
    assert b"lmax" in L.si_strerror(2)
    # null arguments are rejected before any CUDA call
    assert L.si_init(0, 4, 5, None, 0, 0, None) == 1
    assert L.si_finalize(None) == 1
    assert L.si_basis_destroy(None) == 1


def test_no_cpu_fallback():
    import torch

    import sympleints_b200

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(sympleints_b200.SympleintsB200Error):
        sympleints_b200.get_engine()


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle."""
    for p in (REPO / "sympleints_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
    for p in (REPO / "sympleints_b200" / "csrc").rglob("*"):
        if p.is_file():
            assert "oracle" not in p.read_text(), p
