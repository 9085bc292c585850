"""StackedArray: the lazy SUM  A = B + C + ...  as an input type (reference lmdec/array/stacked.py:13-284).

Only the algebra the decomposition path exercises is kept: 2-D members of the full shape, and broadcast members of
shape (p,), (1, p) or (n, 1) (the rank-1 centring term ``-mean`` of make_snp_array).  Members live on the GPU; a member
may itself be Stacked / Chained / a genotype operator.
"""
from __future__ import annotations

import numpy as np
import torch

from ..device import F64, DeviceArray, as_tensor, require_cuda
from .operators import DenseOperator, as_operator, is_sparse


class _Vector:
    """1-D or degenerate 2-D broadcast member kept as a device vector with its declared shape."""

    def __init__(self, v: torch.Tensor, shape):
        self.v = v.reshape(-1).to(F64)
        self.shape = tuple(shape)
        self.ndim = len(self.shape)

    def to_dense(self):
        return self.v.reshape(self.shape)


def _member(a, device):
    from ..store import GenotypeOperator, _Transposed
    from .chained import ChainedArray
    from .scaled_legacy import ScaledCenterArray
    if isinstance(a, (_Vector, StackedArray, ChainedArray, GenotypeOperator, _Transposed, ScaledCenterArray)):
        return a
    if hasattr(a, "dot_t"):
        return a
    if is_sparse(a):
        return as_operator(a, device)
    if isinstance(a, DeviceArray):
        a = a.t
    shape = tuple(a.shape)
    if len(shape) == 1 or (len(shape) == 2 and 1 in shape):
        return _Vector(as_tensor(a, device), shape)
    if len(shape) != 2:
        raise NotImplementedError("StackedArray members must be 1-D or 2-D on this path")
    return DenseOperator(a, device=device, keep_dtype=True)


class StackedArray:
    ndim = 2
    group = None

    def __init__(self, arrays, tree_reduce: bool = True, device=None):
        if isinstance(arrays, (np.ndarray, torch.Tensor, DeviceArray)):
            raise ValueError("expected Iterable of arrays. Got a single array")
        device = device or require_cuda()
        self.tree_reduce = tree_reduce
        self._arrays = [_member(a, device) for a in arrays]
        try:
            self.shape = tuple(np.broadcast_shapes(*[tuple(a.shape) for a in self._arrays]))
        except ValueError:
            raise ValueError("expected arrays to have broadcast-able shapes, got {}".format(
                set(tuple(a.shape) for a in self._arrays)))
        if len(self.shape) == 1:
            self.ndim = 1
        self.device = device

    @property
    def arrays(self):
        return self._arrays

    @property
    def local_rows(self):
        return self.shape[0]

    @property
    def T(self):
        if self.ndim != 2:
            raise NotImplementedError
        n, p = self.shape
        out = []
        for a in self._arrays:
            if tuple(a.shape) == (n, p):
                out.append(a.T)
            elif tuple(a.shape) in [(p,), (1, p)]:
                out.append(_Vector(a.v, (p, 1)))
            else:  # (n, 1)
                out.append(_Vector(a.v, (1, n)))
        return StackedArray(out, tree_reduce=self.tree_reduce, device=self.device)

    def dot_t(self, x: torch.Tensor) -> torch.Tensor:
        if self.ndim != 2 or x.ndim > 2:
            raise NotImplementedError
        n, p = self.shape
        if x.shape[0] != p:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(x.shape)))
        total = None
        for a in self._arrays:
            if tuple(a.shape) == (n, p):
                out = a.dot_t(x)
            elif tuple(a.shape) == (n, 1):
                out = a.v[:, None] * x.sum(dim=0) if x.ndim == 2 else a.v * x.sum()
            elif tuple(a.shape) in [(1, p), (p,)]:
                out = a.v @ x  # broadcast over rows when added
            else:
                raise NotImplementedError
            total = out if total is None else total + out
        if total.ndim < x.ndim or (x.ndim == 1 and total.ndim == 0):
            total = total.expand((n,) + tuple(x.shape[1:])).clone()
        return total

    def tdot_t(self, y):
        return self.T.dot_t(y)

    def dot(self, x):
        return DeviceArray(self.dot_t(as_tensor(x, self.device)))

    def __getitem__(self, item):
        if not isinstance(item, tuple):
            item = (item,)
        if len(item) > self.ndim:
            raise IndexError("Too many indices for array")
        item = item + (slice(None),) * (self.ndim - len(item))
        subs = []
        for a in self._arrays:
            if isinstance(a, _Vector):
                v = a.to_dense()
                sub_item = tuple(i if s > 1 else (slice(None) if isinstance(i, slice) else 0)
                                 for i, s in zip(item[-a.ndim:], a.shape))
                sub_item = tuple(torch.as_tensor(i, device=v.device) if isinstance(i, np.ndarray) else i
                                 for i in sub_item)
                sub = v[sub_item]
                subs.append(_Vector(sub, tuple(sub.shape)))
            else:
                subs.append(a[item])
        return StackedArray(subs, tree_reduce=self.tree_reduce, device=self.device)

    def to_dense(self) -> torch.Tensor:
        total = None
        for a in self._arrays:
            d = a.to_dense()
            total = d if total is None else total + d
        return total.expand(self.shape) if tuple(total.shape) != self.shape else total

    @property
    def array(self):
        return DeviceArray(self.to_dense())

    def mean(self, axis=None):
        return DeviceArray(self.to_dense().mean() if axis is None else self.to_dense().mean(dim=axis))

    def compute(self):
        return self.to_dense().cpu().numpy()

    def persist(self):
        return self

    def rechunk(self, *a, **k):
        return self

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a.astype(dtype) if dtype is not None else a
