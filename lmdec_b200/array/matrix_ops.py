"""Iterate -> (U, S, V) and small helpers (reference lmdec/array/matrix_ops.py:19-217)."""
from __future__ import annotations

import numpy as np
import torch

from ..device import F64, DeviceArray, as_tensor
from .linalg import tsqr_svd


def svd_to_trunc_svd(u=None, s=None, v=None, k=None):
    """matrix_ops.py:140-155"""
    out = []
    if u is not None:
        out.append(u[:, :k])
    if s is not None:
        out.append(s[:k])
    if v is not None:
        out.append(v[:k, :])
    return out[0] if len(out) == 1 else tuple(out)


def _tdot(a, x):
    return a.tdot_t(x) if hasattr(a, "tdot_t") else a.T.dot_t(x)


def subspace_to_V(x, a, k=None, group=None):
    """Top-k right singular vectors of ``a`` from the active subspace x: left singular vectors of a^T x
    (matrix_ops.py:89-99).  Returns a [k, p] DeviceArray."""
    from .operators import as_operator
    a = as_operator(a)
    x_t = _tdot(a, as_tensor(x)).to(F64)
    # sample shards: a^T x is replicated on every rank, no collective; SNP shards: its rows are sharded
    v, _, _ = tsqr_svd(x_t, getattr(getattr(a, "store", None), "snp_group", None))
    v = v.T
    if k:
        v = v[:k, :]
    return DeviceArray(v)


def subspace_to_SVD(x, a=None, k=None, full_v: bool = False, sqrt_s: bool = True, log: int = 0, group=None):
    """tsqr-SVD of the iterate, sqrt of the singular values, truncation (matrix_ops.py:71-86)."""
    x_t = as_tensor(x)
    if x_t.ndim == 1:
        x_t = x_t[:, None]
    u, s, vh = tsqr_svd(x_t, group)
    if full_v:
        v = subspace_to_V(x_t, a, k, group).t
    else:
        v = torch.as_tensor(vh, device=x_t.device)
    s_t = torch.as_tensor(s, device=x_t.device)
    if sqrt_s:
        s_t = torch.sqrt(s_t)
    if k:
        u, s_t, v = u[:, :k], s_t[:k], v[:k, :]
    out = (DeviceArray(u), DeviceArray(s_t), DeviceArray(v))
    return out + ({},) if log > 0 else out


def diag_dot(diag_array, x, log: int = 0):
    """diag(d) @ x without forming the diagonal (matrix_ops.py:174-198)."""
    d = as_tensor(diag_array)
    xt = as_tensor(x)
    if xt.ndim not in (1, 2):
        raise ValueError("x must have (M, K) or (K, ). Current Shape = {}".format(tuple(xt.shape)))
    if d.shape[0] != xt.shape[0]:
        raise ValueError("shapes {} and {} not aligned".format(tuple(d.shape), tuple(xt.shape)))
    if d.ndim not in (1, 2):
        raise ValueError("diag_array must have dimension (K, ) or (K, 1)")
    d = d.reshape(-1)
    return DeviceArray(d * xt if xt.ndim == 1 else d[:, None] * xt)


def reshape_degenerate_2d_array(f):
    """(n, 1) inputs are squeezed for the call and re-expanded afterwards (matrix_ops.py:201-217)."""
    from functools import wraps

    @wraps(f)
    def wrapped(self, x, *args, **kwargs):
        shape = tuple(x.shape) if hasattr(x, "shape") else np.shape(x)
        reshape = len(shape) == 2 and shape[1] == 1
        if reshape:
            x = as_tensor(x).reshape(-1)
        r = f(self, x, *args, **kwargs)
        if reshape:
            t = r.t if isinstance(r, DeviceArray) else as_tensor(r)
            return DeviceArray(t[:, None])
        return r

    return wrapped
