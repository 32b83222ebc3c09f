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
