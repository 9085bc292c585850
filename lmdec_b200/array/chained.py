"""ChainedArray: the lazy PRODUCT  A = B @ C @ ...  with 1-D members read as diagonals (reference
lmdec/array/chained.py:12-269).  ``dot`` applies the members right to left (chained.py:169-177); ``T`` reverses and
transposes (chained.py:159-163); ``__getitem__`` takes slices only and cuts the first member's rows / last member's
columns (chained.py:182-252)."""
from __future__ import annotations

import numpy as np
import torch

from ..device import F64, DeviceArray, as_tensor, require_cuda
from .operators import DenseOperator, as_operator, is_sparse


class _Diag:
    ndim = 1

    def __init__(self, v: torch.Tensor):
        self.v = v.reshape(-1).to(F64)
        self.shape = (self.v.shape[0],)

    @property
    def T(self):
        return self

    def dot_t(self, x):
        if x.shape[0] != self.v.shape[0]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(x.shape)))
        return self.v[:, None] * x if x.ndim == 2 else self.v * x

    def to_dense(self):
        return torch.diag(self.v)

    def __getitem__(self, item):
        return _Diag(self.v[item])


def _member(a, device):
    if hasattr(a, "dot_t"):
        return a
    if is_sparse(a):
        return as_operator(a, device)
    if isinstance(a, DeviceArray):
        a = a.t
    if len(a.shape) == 1:
        return _Diag(as_tensor(a, device))
    if len(a.shape) != 2:
        raise NotImplementedError("expected 1D or 2D arrays")
    return DenseOperator(a, device=device, keep_dtype=True)


class ChainedArray:
    ndim = 2
    group = None

    def __init__(self, arrays, device=None):
        if isinstance(arrays, (np.ndarray, torch.Tensor, DeviceArray)):
            raise ValueError("expected Iterable of arrays. Got a single array")
        device = device or require_cuda()
        self.device = device
        self._arrays = [_member(a, device) for a in arrays]
        prev = None
        for a in self._arrays:
            n, p = (a.shape[0], a.shape[0]) if a.ndim == 1 else a.shape
            if prev is not None and prev != n:
                raise ValueError("expected chain-able dimensions, got {}".format([tuple(b.shape) for b in self._arrays]))
            prev = p
        first, last = self._arrays[0], self._arrays[-1]
        self.shape = (first.shape[0], last.shape[0] if last.ndim == 1 else last.shape[1])

    @property
    def arrays(self):
        return self._arrays

    @property
    def local_rows(self):
        first = self._arrays[0]
        return getattr(first, "local_rows", first.shape[0])

    @property
    def T(self):
        return ChainedArray(list(reversed([a.T for a in self._arrays])), device=self.device)

    def dot_t(self, x):
        if x.ndim > 2:
            raise NotImplementedError
        if x.shape[0] != self.shape[1]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(x.shape)))
        for a in reversed(self._arrays):
            x = a.dot_t(x)
        return x

    def tdot_t(self, y):
        for a in self._arrays:
            y = a.tdot_t(y) if hasattr(a, "tdot_t") else a.T.dot_t(y)
        return y

    def dot(self, x):
        return DeviceArray(self.dot_t(as_tensor(x, self.device)))

    def __getitem__(self, item):
        if not isinstance(item, tuple):
            item = (item,)
        if len(item) > 2:
            raise IndexError("Too many indices for array")
        if any(not isinstance(i, slice) for i in item):
            raise NotImplementedError("Only implements slice indexing")
        arrays = list(self._arrays)
        full = slice(None, None, None)
        if item[0] != full:
            arrays[0] = arrays[0][item[0]] if arrays[0].ndim == 1 else arrays[0][item[0], :]
            if arrays[0].ndim == 1 and len(arrays) > 1:
                # a sliced leading diagonal no longer chains with the next member: cut its rows too
                nxt = arrays[1]
                arrays[1] = nxt[item[0]] if nxt.ndim == 1 else nxt[item[0], :]
        if len(item) == 2 and item[1] != full:
            arrays[-1] = arrays[-1][item[1]] if arrays[-1].ndim == 1 else arrays[-1][:, item[1]]
            if arrays[-1].ndim == 1 and len(arrays) > 1:
                prv = arrays[-2]
                arrays[-2] = prv[item[1]] if prv.ndim == 1 else prv[:, item[1]]
        return ChainedArray(arrays, device=self.device)

    def to_dense(self):
        out = None
        for a in self._arrays:
            d = a.to_dense()
            out = d if out is None else out @ d
        return out

    @property
    def array(self):
        return DeviceArray(self.to_dense())

    def compute(self):
        return self.to_dense().cpu().numpy()

    def persist(self):
        return self

    def rechunk(self, *a, **k):
        return self

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a.astype(dtype) if dtype is not None else a
