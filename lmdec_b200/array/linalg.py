"""Orthonormalisation and small-matrix algebra.

The reference orthonormalises the iterate with dask's ``tsqr`` (lmdec/decomp/iter_methods.py:382; with
``compute_svd=True`` at lmdec/array/matrix_ops.py:72,94 and lmdec/decomp/init_methods.py:35,144).  Here the tall
factor never leaves the GPU: a Gram matrix per rank (``lmdec_gram``), one K x K all-reduce, and the K x K Cholesky /
SVD on the host in float64 (CholeskyQR2).  ``x = q r`` holds with ``q`` orthonormal; ``r`` is upper triangular unless
the rank-revealing fallback had to be taken.
"""
from __future__ import annotations

import numpy as np
import torch

from ..device import F64, gram


def _allreduce(t: torch.Tensor, group):
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, group=group)
    return t


def _whiten_factor(G: np.ndarray):
    """Return (r_inv, r) with r_inv^T G r_inv = I.  Cholesky when G is numerically SPD, eigen-decomposition otherwise."""
    if not np.all(np.isfinite(G)):
        # what LAPACK reports through dask/numpy when NaN/Inf reach the QR: tests/test_iter_methods.py:318-325
        raise np.linalg.LinAlgError("SVD did not converge (non-finite values in the iterate)")
    K = G.shape[0]
    G = (G + G.T) / 2
    try:
        dmax = float(np.max(np.diag(G)))
        if dmax <= 0 or np.min(np.diag(G)) <= dmax * 1e-28:
            raise np.linalg.LinAlgError("degenerate column")
        L = np.linalg.cholesky(G)
        r = L.T
        r_inv = np.linalg.solve(r, np.eye(K))
        if not np.all(np.isfinite(r_inv)):
            raise np.linalg.LinAlgError("singular")
        return r_inv, r, False
    except np.linalg.LinAlgError:
        lam, V = np.linalg.eigh(G)
        lam_max = max(float(lam[-1]), 0.0)
        keep = lam > lam_max * 1e-24 if lam_max > 0 else np.zeros(K, dtype=bool)
        inv_sqrt = np.where(keep, 1.0 / np.sqrt(np.where(keep, lam, 1.0)), 0.0)
        r_inv = V * inv_sqrt
        r = (V * np.sqrt(np.clip(lam, 0, None))).T
        return r_inv, r, not bool(np.all(keep))


def orthonormalize(x: torch.Tensor, group=None, seed: int = 0):
    """CholeskyQR2: returns (q [n,K] orthonormal, r [K,K] float64 on the host) with x = q r.

    Rank-deficient iterates (the reference's Householder QR completes them with arbitrary orthonormal columns) get
    their null directions refilled with seeded Gaussian columns before the second pass.
    """
    if x.ndim != 2:
        raise ValueError("expected a 2-D iterate")
    K = x.shape[1]
    r_total = np.eye(K)
    q = x
    for it in range(4):
        G = _allreduce(gram(q), group).cpu().numpy()
        if it >= 2 and np.max(np.abs(G - np.eye(K))) < 1e-9:
            break
        r_inv, r, deficient = _whiten_factor(G)
        q = q @ torch.as_tensor(r_inv, device=x.device)
        r_total = r @ r_total
        if deficient:
            dead = np.where(np.abs(r_inv).sum(axis=0) == 0)[0]
            gen = torch.Generator(device="cpu").manual_seed(seed + 17 * it)
            fill = torch.randn((x.shape[0], len(dead)), generator=gen, dtype=F64).to(x.device)
            q = q.clone()
            q[:, torch.as_tensor(dead, device=x.device)] = fill
            # the refilled columns carry no part of x: zero rows of r
            r_total[dead, :] = 0.0
    return q, r_total


def tsqr_svd(x: torch.Tensor, group=None):
    """``tsqr(x, compute_svd=True)``: (u [n,K] on the device, s [K] host, vh [K,K] host)."""
    q, r = orthonormalize(x, group)
    u_r, s, vh = np.linalg.svd(r, full_matrices=False)
    u = q @ torch.as_tensor(u_r, device=x.device)
    return u, s, vh
