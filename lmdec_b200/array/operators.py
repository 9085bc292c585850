"""Device-side leaves of the operator protocol: dense float matrices, sparse matrices, vectors.

``as_operator`` is the ``da.array(data)`` coercion of the drivers (lmdec/decomp/iter_methods.py:357,543): Stacked /
Chained / ScaledCenter / genotype operators are kept, everything else becomes a dense device matrix whose products are
plain library GEMMs (cuBLAS through torch) in float64.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..device import F64, DeviceArray, as_tensor, require_cuda


class DenseOperator:
    """A [n, p] floating matrix resident in HBM."""

    ndim = 2
    group = None

    def __init__(self, a, device=None, keep_dtype: bool = False):
        device = device or require_cuda()
        if isinstance(a, DeviceArray):
            a = a.t
        if isinstance(a, torch.Tensor):
            t = a.to(device)
        else:
            arr = np.asarray(a)
            if arr.dtype == object:
                raise TypeError("object arrays are not supported")
            t = torch.as_tensor(np.ascontiguousarray(arr), device=device)
        if not (keep_dtype and t.dtype == torch.float32):
            t = t.to(F64)
        if t.ndim != 2:
            raise ValueError("expected a 2-D array")
        self.a = t

    @property
    def shape(self):
        return tuple(self.a.shape)

    @property
    def local_rows(self):
        return self.a.shape[0]

    @property
    def T(self):
        out = DenseOperator.__new__(DenseOperator)
        out.a = self.a.T
        return out

    # float32 matrices: "fp32" (default) = plain fp32 library GEMM; "tf32x3" = error-compensated split
    # A = A_hi + A_lo, x = x_hi + x_lo on the TF32 tensor cores (A_hi x_hi + A_lo x_hi + A_hi x_lo: fp32-level accuracy).
    # Measured on configs[4] (1M x 4k, K = 110) the three skinny cuBLAS TF32 GEMMs are slower than the fp32 one (58 vs
    # 47 ms per iteration), so it stays opt-in until the dense path gets its own tcgen05 kernel.
    f32_mode = os.environ.get("LMDEC_DENSE_F32", "fp32")

    def _split(self):
        if getattr(self, "_hi", None) is None:
            hi = (self.a.view(torch.int32) & -8192).view(torch.float32)      # keep the 10 mantissa bits TF32 keeps
            self._hi, self._lo = hi, self.a - hi
        return self._hi, self._lo

    def _mm32(self, x, transpose):
        xf = x.to(F64)
        if self.f32_mode != "tf32x3":
            a = self.a.T if transpose else self.a
            return (a @ xf.to(torch.float32)).to(F64)
        hi, lo = self._split()
        if transpose:
            hi, lo = hi.T, lo.T
        x32 = xf.to(torch.float32)
        x_hi = (x32.view(torch.int32) & -8192).view(torch.float32)
        x_lo = x32 - x_hi
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            out = (hi @ x_hi).to(F64) + (lo @ x_hi).to(F64) + (hi @ x_lo).to(F64)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return out

    def dot_t(self, x):
        if x.shape[0] != self.a.shape[1]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(x.shape)))
        if self.a.dtype == torch.float32:
            return self._mm32(x, False)
        return (self.a @ x.to(self.a.dtype)).to(F64)

    def tdot_t(self, y):
        if y.shape[0] != self.a.shape[0]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape[::-1], tuple(y.shape)))
        if self.a.dtype == torch.float32:
            return self._mm32(y, True)
        return (self.a.T @ y.to(self.a.dtype)).to(F64)

    def dot(self, x):
        return DeviceArray(self.dot_t(as_tensor(x, self.a.device)))

    def __getitem__(self, item):
        if isinstance(item, tuple):
            item = tuple(torch.as_tensor(i, device=self.a.device) if isinstance(i, np.ndarray) else i for i in item)
        elif isinstance(item, np.ndarray):
            item = torch.as_tensor(item, device=self.a.device)
        sub = self.a[item]
        out = DenseOperator.__new__(DenseOperator)
        out.a = sub if sub.ndim == 2 else sub.reshape(1, -1)
        return out

    def to_dense(self):
        return self.a.to(F64)

    def compute(self):
        return self.a.cpu().numpy()

    def persist(self):
        return self

    def rechunk(self, *a, **k):
        return self

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a.astype(dtype) if dtype is not None else a


class SparseOperator:
    """scipy.sparse / pydata-sparse matrix on the device (the COO mask member of the reference's StackedArray,
    lmdec/decomp/utils.py:185-188)."""

    ndim = 2
    group = None

    def __init__(self, a, device=None):
        import scipy.sparse as sp
        device = device or require_cuda()
        if not sp.issparse(a):
            if hasattr(a, "tocsr"):
                a = a.tocsr()
            elif hasattr(a, "to_scipy_sparse"):
                a = a.to_scipy_sparse()
            else:
                raise TypeError("not a sparse matrix")
        csr = a.tocsr().astype(np.float64)
        self.shape = tuple(csr.shape)
        self._host = csr
        self.a = torch.sparse_csr_tensor(torch.as_tensor(csr.indptr.astype(np.int64)),
                                         torch.as_tensor(csr.indices.astype(np.int64)),
                                         torch.as_tensor(csr.data), size=csr.shape, dtype=F64).to(device)
        csc_t = a.T.tocsr().astype(np.float64)
        self.at = torch.sparse_csr_tensor(torch.as_tensor(csc_t.indptr.astype(np.int64)),
                                          torch.as_tensor(csc_t.indices.astype(np.int64)),
                                          torch.as_tensor(csc_t.data), size=csc_t.shape, dtype=F64).to(device)

    @property
    def local_rows(self):
        return self.shape[0]

    @property
    def T(self):
        return SparseOperator(self._host.T, device=self.a.device)

    def dot_t(self, x):
        return self.a @ x if x.ndim == 2 else (self.a @ x[:, None])[:, 0]

    def tdot_t(self, y):
        return self.at @ y if y.ndim == 2 else (self.at @ y[:, None])[:, 0]

    def dot(self, x):
        return DeviceArray(self.dot_t(as_tensor(x, self.a.device)))

    def __getitem__(self, item):
        return SparseOperator(self._host[item], device=self.a.device)

    def to_dense(self):
        return self.a.to_dense()

    def compute(self):
        return np.asarray(self._host.todense())


def is_sparse(a) -> bool:
    """lmdec/array/utils.py:6-26"""
    try:
        import scipy.sparse as sp
        if sp.issparse(a):
            return True
    except ImportError:  # pragma: no cover
        pass
    return type(a).__module__.split(".")[0] == "sparse" or isinstance(a, SparseOperator)


def as_operator(data, device=None):
    from ..store import GenotypeOperator, GenotypeStore, _Transposed
    from .chained import ChainedArray
    from .scaled_legacy import ScaledCenterArray
    from .stacked import StackedArray
    if isinstance(data, (StackedArray, ChainedArray)):
        # the make_snp_array pattern spelled with the containers (lmdec/decomp/utils.py:278-289) goes to the 2-bit
        # tile store and the tensor-core products; any other tree keeps the generic container algebra
        if os.environ.get("LMDEC_RECOGNISE", "1") != "0":
            cached = getattr(data, "_geno_op", False)
            if cached is False:
                from .recognize import genotype_operator_from_containers
                cached = data._geno_op = genotype_operator_from_containers(data)
            if cached is not None:
                return cached
        return data
    if isinstance(data, GenotypeStore):
        # raw genotypes (what load_plink_array returns): usable as they are when nothing is missing; with missing
        # codes the reference's float matrix would carry NaN into the QR (tests/test_iter_methods.py:318-325)
        if data.has_missing:
            raise np.linalg.LinAlgError("SVD did not converge (the genotype store holds missing codes: wrap it with "
                                        "make_snp_array to mean-impute them)")
        return GenotypeOperator(data, None, None)
    if isinstance(data, (GenotypeOperator, _Transposed, ScaledCenterArray, DenseOperator, SparseOperator)):
        return data
    if is_sparse(data):
        return SparseOperator(data, device=device)
    op = DenseOperator(data, device=device, keep_dtype=True)
    return op
