"""Legacy lazily centred / scaled operator  B = (A - 1 mu^T) D  (reference lmdec/array/scaled_legacy.py:11-612).

API kept: ``ScaledCenterArray(scale, center, factor, std_dist, warn).fit(a)``, ``dot``, ``T``, ``sym_mat_mult``,
``array``, ``__getitem__`` (refits the moments on the selection, :577-612), ``fromScaledArray`` (:211-249),
``center_vector`` / ``scale_vector``, and ``ArrayMoment`` with the ``std <= 1e-6 -> 1`` guard (:60-85).

When the fitted matrix is a complete genotype matrix (only 0/1/2) it is held as a 2-bit tile store and the two big
products run on the tensor cores (``GenotypeOperator`` with no centring / scaling folded in: the legacy algebra is
applied around it exactly as ``_center_x`` / ``_scale_x`` do, :452-524).  Any other matrix is a dense device matrix.
"""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np
import torch

from ..device import F64, DeviceArray, as_tensor, require_cuda
from .matrix_ops import reshape_degenerate_2d_array
from .operators import DenseOperator


def _raw_operator(a, device):
    """Genotype tile store when `a` holds only 0/1/2, dense float64 otherwise.  Returns (operator, dense_or_None)."""
    from ..store import GenotypeOperator, GenotypeStore
    if isinstance(a, GenotypeStore):
        return GenotypeOperator(a, None, None), None
    if isinstance(a, DeviceArray):
        a = a.t
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(a)))
    if t.ndim != 2:
        raise ValueError("Cannot fit non-2D array.")
    t = t.to(device)
    is_geno = False
    if t.numel() > 0:
        if t.dtype in (torch.int8, torch.uint8, torch.int16, torch.int32, torch.int64, torch.bool):
            is_geno = bool(((t >= 0) & (t <= 2)).all().item())
        else:
            is_geno = bool(((t == 0) | (t == 1) | (t == 2)).all().item())
    if is_geno:
        store = GenotypeStore.from_dense(t.to(torch.int8), device=device)
        return GenotypeOperator(store, None, None), None
    dense = t.to(F64)
    return DenseOperator(dense, device=device), dense


class ArrayMoment:
    """Column mean and inverse standard deviation of a fitted array (scaled_legacy.py:11-165)."""

    def __init__(self, a, std_dist: str = "normal", warn: bool = True):
        if std_dist not in ["binom", "normal"]:
            raise ValueError("std_dist must be either `binom` or `normal`")
        self._array = a
        self._std_dist = std_dist
        self._warn = warn
        self._std_tol = 1e-6
        self._center_vector = self._scale_vector = self._sym_scale_vector = None
        self._vector_width = None

    def _std_inverter(self, std: torch.Tensor) -> torch.Tensor:
        bad = std <= self._std_tol
        if bool(bad.any().item()):
            if self._warn:
                warnings.warn("SNP Columns {} have low standard deviation. Setting STD of columns to 1".format(
                    torch.nonzero(bad).reshape(-1).cpu().numpy()))
            std = torch.where(bad, torch.ones_like(std), std)
        return 1.0 / std

    def fit(self):
        from ..store import GenotypeOperator
        a = self._array
        if isinstance(a, GenotypeOperator):
            mean, std = a.store.moments("norm")   # complete data: nan-aware == plain mean / std (:93-95)
        else:
            dense = a.to_dense() if hasattr(a, "to_dense") else as_tensor(a)
            mean = dense.mean(dim=0)
            std = dense.std(dim=0, unbiased=False)
        self._center_vector = mean
        if self._std_dist == "normal":
            self._scale_vector = self._std_inverter(std)
        else:
            p = mean / 2
            self._scale_vector = self._std_inverter(torch.sqrt(2 * p * (1 - p)))
        self._sym_scale_vector = self._scale_vector ** 2

    def fit_x(self, x):
        shape = tuple(x.shape)
        if len(shape) > 2:
            raise ValueError("Cannot fit on a {}-d array.".format(len(shape)))
        self._vector_width = 1 if len(shape) == 1 else shape[1]

    def _need(self, v, msg):
        if v is None:
            raise ValueError(msg)
        return v

    @property
    def center_vector(self):
        return DeviceArray(self._need(self._center_vector, "Must fit on array"))

    @property
    def scale_vector(self):
        return DeviceArray(self._need(self._scale_vector, "Must fit on array"))

    @property
    def sym_scale_vector(self):
        return DeviceArray(self._need(self._sym_scale_vector, "Must fit on array"))

    @property
    def vector_width(self):
        return self._need(self._vector_width, "Must fit on x for vector_width")

    @property
    def scale_matrix(self):
        w = self._need(self._vector_width, "Must fit on x for scale_matrix")
        return DeviceArray(self._scale_vector[:, None].expand(-1, w))

    @property
    def sym_scale_matrix(self):
        w = self._need(self._vector_width, "Must fit on x for sym_scale_matrix")
        return DeviceArray(self._sym_scale_vector[:, None].expand(-1, w))


class ScaledCenterArray:
    ndim = 2
    group = None

    def __init__(self, scale: bool = True, center: bool = True, factor: Optional[str] = None,
                 std_dist: Optional[str] = None, warn: bool = True):
        self.scale = scale
        self.center = center
        if factor not in [None, "n", "p"]:
            raise ValueError('factor, {}, must be in [None, "n", "p"]'.format(factor))
        self.factor = factor
        if std_dist not in [None, "normal", "binom"]:
            raise ValueError("std_dist must be in [None, 'normal', 'binom']")
        self._std_dist = std_dist or "normal"
        self._warn = warn
        self._array = None          # raw operator (genotype store or dense)
        self._dense = None          # dense float64 copy when the raw operator is dense
        self._array_moment = None
        self._factor_value = None
        self._t_flag = False
        self._t_cache = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def fromScaledArray(cls, array, scaled_array: "ScaledCenterArray", factor: Optional[str] = None):
        """New operator over `array` that reuses the moments of `scaled_array` (projection of new samples)."""
        shape = tuple(array.shape)
        if (scaled_array.scale and shape[1] != len(scaled_array.scale_vector)) or \
                (scaled_array.center and shape[1] != len(scaled_array.center_vector)):
            raise ValueError("expected array of shape {} to have matching second dimension with array_moment "
                             "shape".format(shape))
        sa = cls(scale=scaled_array.scale, center=scaled_array.center, factor=factor,
                 std_dist=scaled_array._std_dist, warn=scaled_array._warn)
        sa._array, sa._dense = _raw_operator(array, require_cuda())
        sa._array_moment = ArrayMoment(sa._array, std_dist=sa._std_dist, warn=sa._warn)
        if sa.factor:
            n, p = shape
            sa._factor_value = n if sa.factor == "n" else p
        m = scaled_array._array_moment
        sa._array_moment._center_vector = m._center_vector
        sa._array_moment._scale_vector = m._scale_vector
        sa._array_moment._sym_scale_vector = m._scale_vector ** 2
        return sa

    def fit(self, a, x=None):
        if self._t_flag:
            raise ValueError("Cannot fit on a Transposed Array. Use <instance>.T to recover base array")
        if len(tuple(a.shape)) != 2:
            raise ValueError("Cannot fit non-2D array.")
        self._array, self._dense = _raw_operator(a, require_cuda())
        self._array_moment = ArrayMoment(self._array, std_dist=self._std_dist, warn=self._warn)
        self._array_moment.fit()
        if self.factor:
            n, p = self._array.shape
            self._factor_value = n if self.factor == "n" else p
        if x is not None:
            self.fit_x(x)
        self._t_cache = None

    def fit_x(self, x):
        self._array_moment.fit_x(x)

    # ------------------------------------------------------------------ state
    def _moment(self):
        if self._array_moment is None:
            raise AttributeError("ScaledCenterArray has not been fit")
        return self._array_moment

    @property
    def center_vector(self):
        return self._moment().center_vector

    @property
    def scale_vector(self):
        return self._moment().scale_vector

    @property
    def factor_value(self):
        return self.factor if self._factor_value is None else self._factor_value

    @property
    def shape(self):
        s = tuple(self._array.shape)
        return tuple(reversed(s)) if self._t_flag else s

    @property
    def local_rows(self):
        return self.shape[0]

    @property
    def dtype(self):
        return np.dtype("float64")

    @property
    def T(self) -> "ScaledCenterArray":
        if self._t_cache is None:
            t = ScaledCenterArray.__new__(ScaledCenterArray)
            t.__dict__.update(self.__dict__)
            t._t_flag = not self._t_flag
            t._t_cache = self
            self._t_cache = t
        return self._t_cache

    def rechunk(self, rechunk_info=None):
        return self

    def persist(self):
        return self

    # ------------------------------------------------------------------ algebra (torch tensors)
    def _center_x(self, x, dx, transpose=False):
        mu = self._moment()._center_vector
        if transpose:
            s = dx.sum(dim=0)
            return x - (torch.outer(mu, s) if dx.ndim == 2 else mu * s)
        return x - mu @ dx

    def _raw_tdot(self, x):
        from ..store import GenotypeOperator
        if isinstance(self._array, GenotypeOperator):
            return self._array.tdot_t(x, out_dtype=F64)     # the legacy API is float64 end to end
        return self._array.tdot_t(x)

    def _sym_t(self, x):
        x_h = self._raw_tdot(x)
        if self.center:
            x_h = self._center_x(x_h, x, transpose=True)
        if self.scale:
            d2 = self._moment()._sym_scale_vector
            x_h = d2[:, None] * x_h if x_h.ndim == 2 else d2 * x_h
        y = self._array.dot_t(x_h)
        if self.center:
            y = self._center_x(y, x_h, transpose=False)
        if self.factor:
            y = y / self._factor_value
        return y

    def dot_t(self, x):
        d = self._moment()._scale_vector
        if not self._t_flag and self.scale:
            x = d[:, None] * x if x.ndim == 2 else d * x
        y = self._raw_tdot(x) if self._t_flag else self._array.dot_t(x)
        if self.center:
            y = self._center_x(y, x, transpose=self._t_flag)
        if self._t_flag and self.scale:
            y = d[:, None] * y if y.ndim == 2 else d * y
        return y

    def tdot_t(self, y):
        return self.T.dot_t(y)

    # ------------------------------------------------------------------ public products
    @reshape_degenerate_2d_array
    def sym_mat_mult(self, x):
        """factor^-1 . B B^T x  (scaled_legacy.py:379-399)."""
        if self._t_flag:
            raise NotImplementedError("sym_mat_mult is only implemented for non-transposed arrays")
        xt = as_tensor(x)
        if xt.shape[0] != self.shape[0]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(xt.shape)))
        return DeviceArray(self._sym_t(xt))

    @reshape_degenerate_2d_array
    def dot(self, x):
        xt = as_tensor(x)
        if xt.shape[0] != self.shape[1]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(xt.shape)))
        return DeviceArray(self.dot_t(xt))

    def to_dense(self):
        a = self._array.to_dense()
        if self.center:
            a = a - self._moment()._center_vector
        if self.scale:
            a = a * self._moment()._scale_vector
        return a.T if self._t_flag else a

    @property
    def array(self):
        n, p = self.shape
        if self._warn and n * p >= 1e9:
            warnings.warn("Array is Large, {}".format(self.shape))
        return DeviceArray(self.to_dense())

    def compute(self):
        return self.to_dense().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, item):
        if isinstance(item, tuple) and len(item) > 2:
            raise IndexError("too many indices for array")
        new = ScaledCenterArray(scale=self.scale, center=self.center, factor=self.factor, std_dist=self._std_dist,
                                warn=self._warn)
        base = self._array.to_dense()
        if self._t_flag:
            base = base.T
        if isinstance(item, tuple):
            item = tuple(torch.as_tensor(i, device=base.device) if isinstance(i, np.ndarray) else i for i in item)
        elif isinstance(item, np.ndarray):
            item = torch.as_tensor(item, device=base.device)
        sub = base[item]
        if sub.ndim == 1:
            sub = sub[None, :] if not isinstance(item, tuple) or not isinstance(item[0], slice) else sub[:, None]
        new.fit(sub)
        return new
