"""Synthetic shell sets of the benchmark configurations.

Seeded with ``numpy.random.default_rng(20261019)``; used by bench.py, the tests
and benchmark.py.  These are inputs only -- no integrals are computed here.
"""
import numpy as np

from .shells import Shell

SEED = 20261019


def lattice_sites(n, spacing=2.8, jitter=0.3, rng=None, side=None):
    """First ``n`` sites of a cubic lattice with uniform jitter."""
    if side is None:
        side = int(np.ceil(n ** (1 / 3) - 1e-9))
    pts = []
    for i in range(side):
        for j in range(side):
            for k in range(side):
                pts.append((i, j, k))
    pts = np.array(pts[:n], dtype=float) * spacing
    pts += rng.uniform(-jitter, jitter, size=pts.shape)
    return pts


def _shell(rng, L, K, center, lo, hi, atom):
    exps = np.sort(10 ** rng.uniform(lo, hi, K))[::-1].copy()
    coeffs = rng.uniform(0.2, 1.0, K)
    return Shell(L, center, coeffs, exps, center_ind=atom, atomic_num=6)


def plumbing_basis(natoms=20, seed=SEED):
    """Config 1: 20 atoms, {s: K=3, s: K=1, p: K=2} -> 60 shells, 160 functions."""
    rng = np.random.default_rng(seed)
    sites = lattice_sites(natoms, rng=rng, side=3)
    shells = []
    for a, c in enumerate(sites):
        shells.append(_shell(rng, 0, 3, c, -1, 2, a))
        shells.append(_shell(rng, 0, 1, c, -1, 2, a))
        shells.append(_shell(rng, 1, 2, c, -1, 2, a))
    return shells, sites


ORBITAL_LAYOUT = ((0, 6), (0, 1), (1, 3), (2, 2), (3, 1), (4, 1))  # (L, K) per atom -> 26 sph functions
AUX_LAYOUT = ((0, 7), (1, 6), (2, 5), (3, 3), (4, 2), (5, 1))      # (L, count) per atom, K=1 -> 100 functions


def orbital_basis(natoms=50, seed=SEED, layout=ORBITAL_LAYOUT, lmax=None):
    """Config 2: 50 atoms -> 300 shells, 1300 spherical functions (Ru(bpy)3-sized)."""
    rng = np.random.default_rng(seed)
    sites = lattice_sites(natoms, rng=rng, side=4 if natoms > 27 else None)
    shells = []
    for a, c in enumerate(sites):
        for L, K in layout:
            if lmax is not None and L > lmax:
                continue
            shells.append(_shell(rng, L, K, c, -1, 2.5 if L == 0 else 1.0, a))
    return shells, sites


def aux_basis(sites, seed=SEED + 1, layout=AUX_LAYOUT, lmax=None):
    """Config 4: per atom {s:7, p:6, d:5, f:3, g:2, h:1}, K=1 -> 1200 shells, 5000 functions."""
    rng = np.random.default_rng(seed)
    shells = []
    for a, c in enumerate(sites):
        for L, n in layout:
            if lmax is not None and L > lmax:
                continue
            for _ in range(n):
                shells.append(_shell(rng, L, 1, c, -0.7, 1.5, a))
    return shells


def point_charges(sites, n=1024, seed=SEED + 2):
    """Config 3: nuclei (q = -6) + external charges q in U(-1,1) in the padded bounding box,
    rejecting points closer than 1.5 bohr to any atom."""
    rng = np.random.default_rng(seed)
    nat = len(sites)
    lo, hi = sites.min(0) - 6.0, sites.max(0) + 6.0
    pts = []
    while len(pts) < n - nat:
        cand = rng.uniform(lo, hi, size=(4 * (n - nat), 3))
        d = np.linalg.norm(cand[:, None, :] - sites[None, :, :], axis=2).min(1)
        for p in cand[d >= 1.5]:
            pts.append(p)
            if len(pts) == n - nat:
                break
    xyz = np.concatenate([sites, np.array(pts)], axis=0)
    q = np.concatenate([-6.0 * np.ones(nat), rng.uniform(-1, 1, n - nat)])
    return xyz, q
