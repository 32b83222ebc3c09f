"""Shell container mirroring sympleints/testing/shells.py:8-26 and the cGTO
normalisation of sympleints/testing/normalization.py:29-43 (input contract of
``--normalize cgto``: the radial norm lives inside the coefficients)."""
from dataclasses import dataclass

import numpy as np


def norm_cgto_lmn(coeffs, exps, L):
    """Contraction coefficients times per-primitive radial norms.

    N_i = sqrt(a_i^(L+3/2) / (pi^(3/2)/2^L * sum_jk d_j d_k (sqrt(a_j a_k))^(L+3/2) / (a_j+a_k)^(L+3/2)))
    """
    coeffs = np.asarray(coeffs, dtype=float)
    exps = np.asarray(exps, dtype=float)
    e = L + 1.5
    s = np.sum(coeffs[:, None] * coeffs[None, :] * np.sqrt(exps[:, None] * exps[None, :]) ** e
               / (exps[:, None] + exps[None, :]) ** e)
    return coeffs * np.sqrt(exps**e / (np.pi**1.5 / 2**L * s))


@dataclass
class Shell:
    """Same fields as the reference's test Shell; coefficients are normalised on
    construction unless ``normalized=True``."""
    L: int
    center: np.ndarray
    coeffs: np.ndarray
    exps: np.ndarray
    center_ind: int = 0
    atomic_num: int = 0
    normalized: bool = False

    def __post_init__(self):
        self.center = np.asarray(self.center, dtype=float)
        self.exps = np.asarray(self.exps, dtype=float)
        self.coeffs = np.asarray(self.coeffs, dtype=float)
        if not self.normalized:
            self.coeffs = norm_cgto_lmn(self.coeffs, self.exps, self.L)
            self.normalized = True

    @property
    def cart_size(self):
        return (self.L + 2) * (self.L + 1) // 2

    @property
    def sph_size(self):
        return 2 * self.L + 1


def shells_to_soa(shells):
    """Struct-of-arrays view consumed by si_basis_create."""
    L = np.array([s.L for s in shells], dtype=np.int32)
    nprim = np.array([len(s.exps) for s in shells], dtype=np.int32)
    centers = np.ascontiguousarray(np.array([s.center for s in shells], dtype=np.float64).reshape(-1, 3))
    exps = np.ascontiguousarray(np.concatenate([s.exps for s in shells]).astype(np.float64))
    coefs = np.ascontiguousarray(np.concatenate([s.coeffs for s in shells]).astype(np.float64))
    return L, nprim, centers, exps, coefs


def class_major_order(shells):
    """Stable permutation that sorts shells by angular momentum ("class-major" order).

    The 3c2e kernels write one (2Lc+1)-wide column chunk per shell triple; with the auxiliary shells of one class
    adjacent, the chunks of neighbouring tuples fill whole 32-byte sectors and the tensor is written 7 % faster
    (60.6 vs 65.4 ms on the benchmark system; atom-major order: 14 GB of DRAM read-modify-write traffic per 34 GB
    written, profiles/r02_final_traffic.json).  A contiguous shard of such a list also holds only one or two classes
    (15-30 kernel launches instead of 90).  ``[shells[i] for i in order]`` is the reordered list; column block k of a
    tensor built from it belongs to input shell ``order[k]``."""
    return np.argsort(np.array([s.L for s in shells]), kind="stable")


def function_permutation(shells, order, sph=True):
    """Index array ``idx`` with ``T_reordered[..., k] == T_input[..., idx[k]]`` for the shell permutation ``order``."""
    size = [(2 * s.L + 1) if sph else (s.L + 1) * (s.L + 2) // 2 for s in shells]
    off = np.concatenate([[0], np.cumsum(size)])
    return np.concatenate([np.arange(off[i], off[i + 1]) for i in order]) if len(order) else np.zeros(0, dtype=int)
