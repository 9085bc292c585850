"""Convergence metrics (reference lmdec/array/metrics.py:37-235) evaluated on the device.

The reductions over the long axes are C-ABI kernels (``lmdec_qvals``, ``lmdec_gram`` for Vi^T Vj, ``lmdec_sqdiff`` for
the Frobenius norm of rmse_k); only k x k singular values are taken on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from ..device import as_tensor, gram, qvals, sqdiff
from .transform import acc_format_svd


def q_value_converge(si, sj, scale: bool = True, norm: float = 2) -> float:
    """|| si - sj ||_norm, scaled by the mean norm of the two (metrics.py:219-235).

    k-vectors that are already on the host (the drivers get S from the host-side K x K SVD) are reduced there; device
    operands go through the warp-reduction kernel ``lmdec_qvals``."""
    if isinstance(si, np.ndarray) and isinstance(sj, np.ndarray):
        a, b = si.reshape(-1).astype(np.float64), sj.reshape(-1).astype(np.float64)
        if a.shape != b.shape:
            raise ValueError("Shapes of si and sj must match. Current {} != {}".format(a.shape, b.shape))
        acc = np.linalg.norm(a - b, norm)
        if scale:
            acc /= (np.linalg.norm(a, norm) + np.linalg.norm(b, norm)) / 2
        return float(acc)
    si_t, sj_t = as_tensor(si).reshape(-1), as_tensor(sj).reshape(-1)
    if si_t.shape != sj_t.shape:
        raise ValueError("Shapes of si and sj must match. Current {} != {}".format(tuple(si_t.shape),
                                                                                   tuple(sj_t.shape)))
    if norm == 2:
        return float(qvals(si_t, sj_t, scale).item())
    acc = torch.linalg.vector_norm(si_t - sj_t, ord=norm)
    if scale:
        acc = acc / ((torch.linalg.vector_norm(si_t, ord=norm) + torch.linalg.vector_norm(sj_t, ord=norm)) / 2)
    return float(acc.item())


def subspace_dist(vi, vj, s, power=2, epsilon: float = 1e-6, group=None) -> float:
    """1 - <s^power, sv(vi^T vj)> / (sum s^power + epsilon)   (metrics.py:70-105)."""
    vi_t, vj_t = as_tensor(vi), as_tensor(vj)
    s_np = np.asarray(as_tensor(s).cpu().numpy() if not isinstance(s, np.ndarray) else s, dtype=np.float64).reshape(-1)
    if vi_t.ndim == 1:
        vi_t = vi_t[:, None]
    if vj_t.ndim == 1:
        vj_t = vj_t[:, None]
    if vi_t.shape[0] == len(s_np):
        vi_t = vi_t.T
    if vj_t.shape[0] == len(s_np):
        vj_t = vj_t.T
    if vi_t.shape[0] != vj_t.shape[0]:
        raise ValueError("Shape Error between vi, {},  and vj, {}".format(tuple(vi_t.shape), tuple(vj_t.shape)))
    if vi_t.shape[1] == vj_t.shape[1]:
        g = gram(vi_t, vj_t)
    else:
        g = vi_t.T @ vj_t
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(g, group=group)
    l = np.linalg.svd(g.cpu().numpy(), compute_uv=False)
    if max(l) > 1 + 1e-6 or min(l) < -1e-6:
        raise ValueError("Norm of vi or vj was not 1.")
    if power in (float("inf"), -1):
        idx = np.argmin(s_np) if power == -1 else np.argmax(s_np)
        val = s_np[idx]
        s_np = np.zeros_like(s_np)
        s_np[idx] = val
        power = 1
    weighted = float(np.squeeze((s_np ** power).dot(l)))
    return float(1 - weighted / (np.sum(s_np ** power) + epsilon))


def rmse_k(array, u, s, format_us: bool = True, factor=None, group=None) -> float:
    """|| A A^T u / factor - u diag(s) ||_F / sqrt(n k p)   (metrics.py:151-172)."""
    from .operators import as_operator
    op = as_operator(array)
    n, p = op.shape
    if format_us:
        u_t, s_t = acc_format_svd(u, s, op.shape, square=False)
    else:
        u_t, s_t = as_tensor(u), as_tensor(s)
        if s_t.ndim == 2:
            s_t = torch.diagonal(s_t)
    k = u_t.shape[1]
    aatx = op.dot_t(op.tdot_t(u_t) if hasattr(op, "tdot_t") else op.T.dot_t(u_t))
    return _rmse_from_products(aatx, u_t, s_t, n, k, p, factor, group)


def _rmse_from_products(aatx, u_t, s_t, n, k, p, factor, group=None) -> float:
    if factor is not None:
        if factor is False or factor == 0:
            # reference quirk kept (iter_methods.py:364 + metrics.py:158-159): factor=None upstream -> divide by False
            return float("inf")
        aatx = aatx / factor
    sq = sqdiff(aatx, u_t, s_t)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(sq, group=group)
    return float(torch.sqrt(sq).item() / np.sqrt(n * k * p))
