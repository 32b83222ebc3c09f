"""Density-fitting consumers either side of the hot path.

The 3c2e tensor (ab|P) and the 2c2e metric (P|Q) are what sympleints' generated functions are used for in a
density-fitting code: the fitted tensor is ``B[ab, Q] = sum_P (ab|P) (L^-T)[P, Q]`` with ``(P|Q) = L L^T``, so
that ``(ab|cd) ~= sum_Q B[ab, Q] B[cd, Q]``.  Both steps are dense FP64 linear algebra on tensors that are already
resident in HBM -- plain library calls (cuSOLVER Cholesky, cuBLAS triangular solve through torch.linalg), kept
on the device so that the sharded tensor never has to be gathered on the host.
"""


def metric_cholesky(engine, aux_basis, normalize="cgto"):
    """Lower Cholesky factor L of the 2c2e metric (P|Q) = L L^T, on the device."""
    torch = engine.torch
    M = engine.int2c2e(aux_basis, normalize=normalize)
    return torch.linalg.cholesky(M)


def fitted_tensor(engine, orb_basis, aux_basis, normalize="cgto", L=None, shell_range=None, out=None):
    """B[(ab), Q] = sum_P (ab|P) (L^-T)[P, Q]; packed rows (shell pairs a >= b), all auxiliary functions.

    With ``shell_range`` only that rank's slab of (ab|P) is computed; the caller all-gathers the slabs first
    (``sympleints_b200.dist.allgather_slabs``), because the triangular solve couples all auxiliary functions."""
    torch = engine.torch
    if L is None:
        L = metric_cholesky(engine, aux_basis, normalize)
    if shell_range is not None:
        raise ValueError("the fit couples all auxiliary functions: gather the slabs first (dist.allgather_slabs)")
    T = engine.int3c2e(orb_basis, aux_basis, normalize=normalize, layout="packed", out=out)
    # solve X L^T = T  <=>  X = T L^-T, in place, row blocks bound the workspace of the library call
    step = 1 << 16
    for r0 in range(0, T.shape[0], step):
        blk = T[r0:r0 + step]
        blk.copy_(torch.linalg.solve_triangular(L, blk.T, upper=False).T)
    return T
