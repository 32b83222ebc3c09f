"""Host-side driver: the batched counterpart of sympleints/testing/intor.py.

``Engine`` owns a C-ABI context on one CUDA device; ``Basis`` a shell set on
that device.  PyTorch is used for buffer ownership and streams only; all
arithmetic happens in the CUDA kernels of libsympleints_b200.so.
"""
import ctypes
import weakref
from ctypes import c_double, c_int, c_longlong, c_void_p

import numpy as np

from . import lib as _lib
from .lib import KEYS, LAYOUT, NCOMP, NORMALIZE, c_double_p, c_int_p, check
from .shells import shells_to_soa


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _lib.SympleintsB200Error("no CUDA device: sympleints_b200 has no CPU fallback")
    return torch


class Basis:
    def __init__(self, engine, shells):
        self.engine = engine
        self.shells = list(shells)
        L, nprim, centers, exps, coefs = shells_to_soa(self.shells)
        self.L = L
        h = c_void_p()
        check(engine.lib.si_basis_create(engine.ctx, len(self.shells), _ip(L), _ip(nprim), _dp(centers),
                                         _dp(exps), _dp(coefs), ctypes.byref(h)))
        self.handle = h
        engine._bases.add(self)
        self.nshell = len(self.shells)
        self.nbf_sph = engine.lib.si_basis_nbf(h, 1)
        self.nbf_cart = engine.lib.si_basis_nbf(h, 0)
        self.packed_rows = engine.lib.si_basis_packed_rows(h)
        sizes = 2 * L + 1
        self.foff_sph = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    def nbf(self, sph=True):
        return self.nbf_sph if sph else self.nbf_cart

    def select_rows(self, shell_begin=0, shell_end=None, skip_diagonal=False):
        """Restrict the batched entry points to the shell pairs (i, j <= i), shell_begin <= i < shell_end."""
        check(self.engine.lib.si_basis_select_rows(self.handle, shell_begin, self.nshell if shell_end is None else shell_end,
                                                   int(skip_diagonal)))
        self.packed_rows = self.engine.lib.si_basis_selected_rows(self.handle)

    def close(self):
        if self.handle is not None:
            if getattr(self.engine, "ctx", None) is not None:   # the context releases its shell sets itself
                self.engine.lib.si_basis_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One context per GPU (``si_init``)."""

    def __init__(self, device=0, lmax=4, lauxmax=5, boys_table=None):
        self.lib = _lib.load()
        self.torch = _torch()
        self.device = int(device)
        self.lmax, self.lauxmax = lmax, lauxmax
        ctx = c_void_p()
        if boys_table is not None:
            tab = np.ascontiguousarray(boys_table, dtype=np.float64)
            check(self.lib.si_init(self.device, lmax, lauxmax, _dp(tab), tab.shape[0], tab.shape[1], ctypes.byref(ctx)))
        else:
            check(self.lib.si_init(self.device, lmax, lauxmax, None, 0, 0, ctypes.byref(ctx)))
        self.ctx = ctx
        self._bases = weakref.WeakSet()
        self.tdev = self.torch.device("cuda", self.device)

    # -- plumbing ---------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None) is not None:
            for b in list(self._bases):
                b.handle = None
            self.lib.si_finalize(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def basis(self, shells):
        return Basis(self, shells)

    def _stream(self):
        return c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    def launch_count(self):
        return int(self.lib.si_launch_count(self.ctx))

    def boys_table(self):
        nr, nc = c_int(), c_int()
        check(self.lib.si_boys_table(self.ctx, None, ctypes.byref(nr), ctypes.byref(nc)))
        tab = np.empty((nr.value, nc.value))
        check(self.lib.si_boys_table(self.ctx, _dp(tab), None, None))
        return tab

    BOYS_MODES = {"reference": 0, "fast": 1, "coul": 2}

    def boys(self, n, xs, mode="reference"):
        """F_n(xs) on the device.  ``mode``: 'reference' (summation order of testing/boys.py), 'fast' (the
        variant the generated 3c2e/2c2e kernels call), 'coul' (the variant of the generated coul kernels)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        out = np.empty_like(xs)
        check(self.lib.si_boys_eval_mode(self.ctx, self.BOYS_MODES[mode], int(n), _dp(xs), xs.size, _dp(out)))
        return out

    def fp64_peak(self, seconds=0.5):
        out = c_double()
        check(self.lib.si_fp64_peak(self.ctx, float(seconds), ctypes.byref(out)))
        return out.value

    def class_stats(self, key, La, Lb, Lc=0, sph=True, normalize="cgto"):
        out = (c_longlong * 4)()
        check(self.lib.si_class_stats(self.ctx, KEYS[key], La, Lb, Lc, int(sph), NORMALIZE[normalize], out))
        return {"fma_prim": out[0], "fma_contr": out[1], "smem_doubles": out[2], "lut_words": out[3]}

    # -- batched integrals (device-resident results, torch tensors) ---------
    def int1e(self, key, basis, sph=True, normalize="cgto", R=(0.0, 0.0, 0.0), out=None):
        """(ncomp, nbf, nbf) tensor (component axis dropped for ncomp == 1, as intor_2c does)."""
        torch = self.torch
        ncomp, nbf = NCOMP[key], basis.nbf(sph)
        if out is None:
            out = torch.empty((ncomp, nbf, nbf), dtype=torch.float64, device=self.tdev)
        Rv = np.ascontiguousarray(R, dtype=np.float64)
        if key == "2c2e":
            check(self.lib.si_int2c2e(self.ctx, basis.handle, int(sph), NORMALIZE[normalize],
                                      c_void_p(out.data_ptr()), self._stream()))
        else:
            check(self.lib.si_int1e(self.ctx, basis.handle, KEYS[key], int(sph), NORMALIZE[normalize], _dp(Rv),
                                    c_void_p(out.data_ptr()), self._stream()))
        return out[0] if ncomp == 1 else out

    def int1e_set(self, basis, sph=True, normalize="cgto", R=(0.0, 0.0, 0.0), out=None):
        """ovlp, kin, dpm, qpm in one call: (11, nbf, nbf) = [ovlp, kin, dpm x y z, qpm xx xy xz yy yz zz]."""
        torch = self.torch
        nbf = basis.nbf(sph)
        if out is None:
            out = torch.empty((11, nbf, nbf), dtype=torch.float64, device=self.tdev)
        Rv = np.ascontiguousarray(R, dtype=np.float64)
        check(self.lib.si_int1e_set(self.ctx, basis.handle, int(sph), NORMALIZE[normalize], _dp(Rv),
                                    c_void_p(out.data_ptr()), self._stream()))
        return out

    def int2c2e(self, basis, sph=True, normalize="cgto", out=None):
        return self.int1e("2c2e", basis, sph, normalize, out=out)

    def coul(self, basis, charges_xyz, q, sph=True, normalize="cgto", out=None):
        """sum_C q_C (mu|1/|r-R_C||nu); q = -Z for nuclei."""
        torch = self.torch
        nbf = basis.nbf(sph)
        if out is None:
            out = torch.empty((nbf, nbf), dtype=torch.float64, device=self.tdev)
        xyz = np.ascontiguousarray(charges_xyz, dtype=np.float64).reshape(-1, 3)
        qv = np.ascontiguousarray(q, dtype=np.float64).reshape(-1)
        assert xyz.shape[0] == qv.shape[0]
        check(self.lib.si_coul(self.ctx, basis.handle, int(sph), NORMALIZE[normalize], qv.shape[0], _dp(xyz), _dp(qv),
                               c_void_p(out.data_ptr()), self._stream()))
        # the charge arrays are read asynchronously: keep them alive until the stream is done
        self.torch.cuda.current_stream(self.tdev).synchronize()
        return out

    def gto(self, basis, points, sph=True, normalize="cgto", out=None):
        """Values of all basis functions on a grid of points: (nbf, npoints) tensor (batched ``gto`` key)."""
        torch = self.torch
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        nbf = basis.nbf(sph)
        if out is None:
            out = torch.empty((nbf, pts.shape[0]), dtype=torch.float64, device=self.tdev)
        check(self.lib.si_gto(self.ctx, basis.handle, int(sph), NORMALIZE[normalize], pts.shape[0], _dp(pts),
                              c_void_p(out.data_ptr()), self._stream()))
        return out

    def aux_cols(self, aux, shell_begin, shell_end):
        return int(aux.foff_sph[shell_end] - aux.foff_sph[shell_begin])

    def int3c2e(self, orb, aux, shell_range=None, normalize="cgto", layout="full", out=None):
        """(nbf, nbf, naux_local) ['full'] or (packed_rows, naux_local) ['packed'] tensor for the
        auxiliary shells ``shell_range = (begin, end)`` (default: all)."""
        torch = self.torch
        sb, se = (0, aux.nshell) if shell_range is None else shell_range
        ncol = self.aux_cols(aux, sb, se)
        shape = (orb.nbf_sph, orb.nbf_sph, ncol) if layout == "full" else (orb.packed_rows, ncol)
        if out is None:
            out = torch.empty(shape, dtype=torch.float64, device=self.tdev)
        check(self.lib.si_int3c2e(self.ctx, orb.handle, aux.handle, sb, se, NORMALIZE[normalize], LAYOUT[layout],
                                  c_void_p(out.data_ptr()), out.numel(), self._stream()))
        # the aux shell list is uploaded asynchronously from a host temporary inside the call
        return out

    # -- host-buffer entry points (H2D/D2H inside the C call) ----------------
    def int1e_host(self, key, basis, sph=True, normalize="cgto", R=(0.0, 0.0, 0.0)):
        ncomp, nbf = NCOMP[key], basis.nbf(sph)
        out = np.empty((ncomp, nbf, nbf))
        Rv = np.ascontiguousarray(R, dtype=np.float64)
        if key == "2c2e":
            check(self.lib.si_int2c2e_h(self.ctx, basis.handle, int(sph), NORMALIZE[normalize], _dp(out)))
        else:
            check(self.lib.si_int1e_h(self.ctx, basis.handle, KEYS[key], int(sph), NORMALIZE[normalize], _dp(Rv), _dp(out)))
        return out[0] if ncomp == 1 else out

    def coul_host(self, basis, charges_xyz, q, sph=True, normalize="cgto"):
        nbf = basis.nbf(sph)
        out = np.empty((nbf, nbf))
        xyz = np.ascontiguousarray(charges_xyz, dtype=np.float64).reshape(-1, 3)
        qv = np.ascontiguousarray(q, dtype=np.float64).reshape(-1)
        check(self.lib.si_coul_h(self.ctx, basis.handle, int(sph), NORMALIZE[normalize], qv.shape[0], _dp(xyz), _dp(qv), _dp(out)))
        return out

    def int3c2e_host(self, orb, aux, shell_range=None, normalize="cgto", layout="full", out=None):
        sb, se = (0, aux.nshell) if shell_range is None else shell_range
        ncol = self.aux_cols(aux, sb, se)
        shape = (orb.nbf_sph, orb.nbf_sph, ncol) if layout == "full" else (orb.packed_rows, ncol)
        if out is None:
            out = np.empty(shape)
        check(self.lib.si_int3c2e_h(self.ctx, orb.handle, aux.handle, sb, se, NORMALIZE[normalize], LAYOUT[layout],
                                    _dp(out), out.size))
        return out

    def schwarz(self, basis, normalize="cgto"):
        """(D, Q): D[mu, nu] = (mu nu|mu nu) (nbf, nbf) and the shell-pair Schwarz factors Q (nshell, nshell)."""
        torch = self.torch
        D = torch.empty((basis.nbf_sph, basis.nbf_sph), dtype=torch.float64, device=self.tdev)
        Q = torch.empty((basis.nshell, basis.nshell), dtype=torch.float64, device=self.tdev)
        check(self.lib.si_schwarz(self.ctx, basis.handle, NORMALIZE[normalize], c_void_p(D.data_ptr()), c_void_p(Q.data_ptr()),
                                  self._stream()))
        return D, Q

    def set_screening(self, eps):
        """Opt-in primitive pre-screening of the 3c2e kernels (0 = off = parity / benchmark mode)."""
        check(self.lib.si_set_screening(self.ctx, float(eps)))

    def host_breakdown(self):
        """Timing of the last ``int3c2e_host`` call (si_host_breakdown)."""
        out = np.zeros(8)
        check(self.lib.si_host_breakdown(self.ctx, _dp(out)))
        return {"total_s": out[0], "batcher_s": out[1], "alloc_s": out[2], "enqueue_s": out[3], "compute_s": out[4],
                "copy_wait_s": out[5], "chunks": int(out[6]), "chunk_bytes": int(out[7])}

    # -- one block, generated-function signature ------------------------------
    def eval_block(self, key, Ls, ax, da, A, bx=None, db=None, B=None, cx=None, dc=None, C=None, R=None, result=None,
                   sph=True, normalize="cgto"):
        """``f(ax, da, A, bx, db, B[, cx, dc, C], R, result)`` of the generated modules; broadcast-shaped
        exponent/coefficient arrays are accepted (they are flattened).  ``gto``: ``f(ax, da, A, R, result)`` with
        ``Ls = (La,)``; ``prefactor``: rP is passed as ``R``."""
        three = KEYS[key] == 7
        f = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        if key == "gto":
            La = Ls[0]
            ax, da, A, Rv = map(f, (ax, da, A, R))
            n = (2 * La + 1) if sph else (La + 1) * (La + 2) // 2
            if result is None:
                result = np.empty(n)
            buf = result if (result.dtype == np.float64 and result.flags.c_contiguous and result.size == n) else np.empty(n)
            check(self.lib.si_eval_block(self.ctx, KEYS[key], La, 0, 0, int(sph), NORMALIZE[normalize], ax.size, _dp(ax),
                                         _dp(da), _dp(A), 0, None, None, None, 0, None, None, None, _dp(Rv), _dp(buf)))
            if buf is not result:
                result[:n] = buf
            return result
        La, Lb = Ls[0], Ls[1]
        Lc = Ls[2] if three else 0
        ax, da, A, bx, db, B = map(f, (ax, da, A, bx, db, B))
        if three:
            cx, dc, C = map(f, (cx, dc, C))
            kc, pcx, pdc, pC = cx.size, _dp(cx), _dp(dc), _dp(C)
        else:
            kc, pcx, pdc, pC = 0, None, None, None
        Rv = f(R) if R is not None else np.zeros(3)
        size = (lambda l: 2 * l + 1) if (sph or three) else (lambda l: (l + 1) * (l + 2) // 2)
        n = NCOMP[key] * size(La) * size(Lb) * (size(Lc) if three else 1)
        if result is None:
            result = np.empty(n)
        buf = result if (result.dtype == np.float64 and result.flags.c_contiguous and result.size == n) else np.empty(n)
        check(self.lib.si_eval_block(self.ctx, KEYS[key], La, Lb, Lc, int(sph), NORMALIZE[normalize],
                                     ax.size, _dp(ax), _dp(da), _dp(A), bx.size, _dp(bx), _dp(db), _dp(B),
                                     kc, pcx, pdc, pC, _dp(Rv), _dp(buf)))
        if buf is not result:
            result[:n] = buf
        return result
