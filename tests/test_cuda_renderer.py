"""The CudaRenderer plugs into the reference's Renderer surface (sympleints/Renderer.py)."""
import importlib.util
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def _load(path):
    spec = importlib.util.spec_from_file_location(path.stem + "_cuda_test", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_stub_payload_renders_all_classes(tmp_path):
    from sympleints_b200.CudaRenderer import CudaRenderer, stub_functions

    r = CudaRenderer(normalize="cgto")
    f2 = stub_functions("ovlp3d", 4, spherical=True)
    fn = r.render_write(f2, tmp_path)
    mod = _load(fn)
    assert sorted(mod.ovlp3d) == sorted((a, b) for a in range(5) for b in range(5))   # native + swapped classes
    f3 = stub_functions("int3c2e3d_sph", 2, 2, boys_func="oracle.boys", three_center=True)
    fn3 = r.render_write(f3, tmp_path)
    mod3 = _load(fn3)
    assert len(mod3.int3c2e3d_sph) == 9 * 3 and (0, 2, 1) in mod3.int3c2e3d_sph
    cu = (tmp_path / "int3c2e3d_sph.cu").read_text()
    assert "g3_kernel_221" in cu and "si_g3_syncwarp" in cu
    # the Boys table is taken from the --boys-func module
    assert mod3._boys_table().shape == (21, 301)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
def test_plugs_into_reference_run(tmp_path):
    """sympleints.main.run(..., renderers=[CudaRenderer()]) -- the reference's own orchestration."""
    sys.path.insert(0, str(REPO / "oracle" / "refgen" / "stubs"))
    sys.path.insert(0, str(REF))
    os.environ["SYMPLEINTS_REF_WORKERS"] = "1"
    from sympleints.main import Normalization, run
    from sympleints.Renderer import Renderer

    import importlib

    import sympleints_b200.CudaRenderer as cr

    cr = importlib.reload(cr)   # pick up the reference's Renderer ABC now that it is importable
    CudaRenderer = cr.CudaRenderer
    r = CudaRenderer(normalize="cgto")
    assert isinstance(r, Renderer)
    res = run(l_max=1, l_aux_max=1, sph=True, normalization=Normalization.CGTO, out_dir=str(tmp_path), prefix="t",
              keys=["ovlp", "dpm", "3c2e_sph", "gto", "prefactor"], boys_func="oracle.boys", renderers=[r])
    assert set(res["cuda"]) == {"ovlp3d", "dipole3d", "int3c2e3d_sph", "sph_gto3d", "prefactors"}
    gmod = _load(Path(res["cuda"]["sph_gto3d"]))
    assert set(gmod.sph_gto3d) == {(0,), (1,)}
    pmod = _load(Path(res["cuda"]["prefactors"]))
    assert set(pmod.prefactors) == {(0, 0), (1, 0), (0, 1), (1, 1)}
    mod = _load(Path(res["cuda"]["int3c2e3d_sph"]))
    assert set(mod.int3c2e3d_sph) == {(a, b, c) for a in range(2) for b in range(2) for c in range(2)}
    assert mod.NORMALIZE == "cgto" and mod.SPHERICAL


@pytest.mark.gpu
def test_rendered_module_evaluates_on_gpu(tmp_path):
    from oracle import ints
    from sympleints_b200.CudaRenderer import CudaRenderer, stub_functions

    r = CudaRenderer(normalize="cgto")
    mod = _load(r.render_write(stub_functions("dipole3d", 2, ncomponents=3, spherical=True), tmp_path))
    rng = np.random.default_rng(0)
    ax, bx = np.array([1.3, 0.4]), np.array([0.9])
    da, db = ints.norm_cgto_lmn(np.array([0.5, 0.6]), ax, 1), ints.norm_cgto_lmn(np.array([1.0]), bx, 2)
    A, B, R = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    res = np.empty(3 * 3 * 5)
    mod.dipole3d[(1, 2)](ax[:, None], da[:, None], A, bx[None, :], db[None, :], B, R, res)   # reference call style
    np.testing.assert_allclose(res, ints.dpm(1, 2, ax, da, A, bx, db, B, R=R), rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
def test_rendered_gto_and_prefactor_modules(tmp_path):
    """Section 8(f)-1 keys through the plugin: sph_gto3d[(La,)](ax, da, A, R, result) and
    prefactors[(La, Lb)](ax, da, A, bx, db, B, result) with the module-level rP of the reference's code."""
    from oracle import ints
    from sympleints_b200.CudaRenderer import CudaRenderer, stub_functions

    r = CudaRenderer(normalize="cgto")
    gmod = _load(r.render_write(stub_functions("sph_gto3d", 3, spherical=True, one_center=True), tmp_path))
    assert sorted(gmod.sph_gto3d) == [(0,), (1,), (2,), (3,)]
    rng = np.random.default_rng(1)
    ax = np.array([2.1, 0.6, 0.2])
    da = ints.norm_cgto_lmn(np.array([0.3, 0.5, 0.4]), ax, 3)
    A, R = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
    res = np.empty(7)
    gmod.sph_gto3d[(3,)](ax, da, A, R, res)
    np.testing.assert_allclose(res, ints.gto(3, ax, da, A, R), rtol=1e-12, atol=1e-14)
    pmod = _load(r.render_write(stub_functions("prefactors", 2, spherical=True, with_ref_center=False), tmp_path))
    pmod.rP = rng.uniform(-1, 1, 3)
    res = np.empty(5 * 3)
    pmod.prefactors[(2, 1)](1.1, 0.7, A, 0.4, 1.3, A, res)
    np.testing.assert_allclose(res, ints.prefactor(2, 1, [1.1], [0.7], A, [0.4], [1.3], A, R=pmod.rP), rtol=1e-12, atol=1e-14)
