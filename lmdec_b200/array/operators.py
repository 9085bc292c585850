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


class DenseTiles:
    """float32 tile stores of a dense [n, p] matrix (one per orientation, built on first use) and the tcgen05 product
    over them (`lmdec_dense_matmul`, csrc/dense_mma.cuh): error-compensated TF32, float32-level accuracy.

    The reference multiplies dense members chunk by chunk with `np.dot` (lmdec/array/chained.py:169-177); here the
    matrix is streamed once per product at HBM speed.  Like the genotype store it keeps both orientations, so `A.dot`
    and `A.T.dot` are the same kernel ("rows of a store times a narrow operand")."""

    def __init__(self, a: torch.Tensor):
        from .. import _capi
        self._capi = _capi
        self.lib = _capi.load()
        self.ctx = _capi.Context()
        self.a = a                      # [n, p] float32, row-major, on the device
        self.n, self.p = a.shape
        self._stores = {}
        self._img = {}

    def store(self, transpose: bool) -> torch.Tensor:
        st = self._stores.get(transpose)
        if st is None:
            M, Kd = (self.p, self.n) if transpose else (self.n, self.p)
            st = torch.empty(self.lib.lmdec_dense_store_bytes(M, Kd) // 4, dtype=torch.float32, device=self.a.device)
            self._capi.check(self.lib.lmdec_dense_pack(self._capi.ptr(self.a), 0, self.n, self.p, self.a.stride(0),
                                                       int(transpose), self._capi.ptr(st), self._capi.stream_ptr()),
                             "lmdec_dense_pack")
            self._stores[transpose] = st
        return st

    def matmul(self, b: torch.Tensor, transpose: bool, rows: int) -> torch.Tensor:
        """A[0:rows] b (transpose = False, b [p, K]) or A[0:rows]^T b (transpose = True, b [rows, K]); float64 out."""
        capi, lib = self._capi, self.lib
        if b.dtype not in (F64, torch.float32):
            b = b.to(F64)
        if b.stride(1) != 1:
            b = b.contiguous()
        M, Kd = (self.p, self.n) if transpose else (self.n, self.p)
        m_limit, kd_limit = (self.p, rows) if transpose else (rows, self.p)
        K = b.shape[1]
        out = torch.empty((m_limit, K), dtype=F64, device=b.device)
        if m_limit == 0 or K == 0:
            return out
        if kd_limit == 0:
            return out.zero_()
        st = self.store(transpose)
        for k0 in range(0, K, 128):             # one tensor-core N tile (<= 128 operand columns) per launch
            kb = min(128, K - k0)
            N = (kb + 15) // 16 * 16
            key = (transpose, N)
            img = self._img.get(key)
            if img is None:
                nb = lib.lmdec_dense_image_bytes(Kd, N) // 4
                img = self._img[key] = (torch.empty(nb, dtype=torch.float32, device=b.device),
                                        torch.empty(nb, dtype=torch.float32, device=b.device))
            bk, ok = b[:, k0:k0 + kb], out[:, k0:k0 + kb]
            capi.check(lib.lmdec_dense_matmul(
                self.ctx.handle, capi.ptr(st), M, Kd, m_limit, kd_limit, capi.ptr(bk), int(b.dtype == F64),
                b.stride(0), kb, capi.ptr(img[0]), capi.ptr(img[1]), capi.ptr(ok), out.stride(0), 0,
                capi.stream_ptr()), "lmdec_dense_matmul")
        return out


class DenseOperator:
    """A [n, p] floating matrix resident in HBM.

    float32 matrices of at least `native_min_elems` elements multiply through `DenseTiles` (hand-written tcgen05
    kernel); float64 matrices and small ones are plain library GEMMs."""

    ndim = 2
    group = None
    native_min_elems = int(os.environ.get("LMDEC_DENSE_NATIVE_MIN", str(1 << 22)))   # below: not worth two tile stores
    native = os.environ.get("LMDEC_DENSE_NATIVE", "1") != "0"

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
        self._init_views()

    def _init_views(self, root=None, rows=None, transposed=False):
        self._root = root                # the operator that owns the tile stores (None: this one)
        self._rows = rows                # row prefix of the root this view covers (None: all)
        self._transposed = transposed    # this view is root[0:rows].T
        self._tiles = None

    def _view(self, a, rows=None, transposed=False):
        out = DenseOperator.__new__(DenseOperator)
        out.a = a
        out._init_views(self._root or self, rows if rows is not None else self._rows, transposed)
        return out

    def _plain(self, a):
        out = DenseOperator.__new__(DenseOperator)
        out.a = a
        out._init_views()
        return out

    def _native_tiles(self):
        """DenseTiles of the root matrix when this view can use the tensor-core path, else None"""
        root = self._root or self
        a = root.a
        if not (DenseOperator.native and a.is_cuda and a.dtype == torch.float32 and a.ndim == 2
                and a.numel() >= DenseOperator.native_min_elems and a.stride(1) == 1):
            return None
        if root._tiles is None:
            root._tiles = DenseTiles(a)
        return root._tiles

    @property
    def path(self) -> str:
        """'tcgen05' or 'cublas': which product path this operator takes (tests assert on it)"""
        return "tcgen05" if self._native_tiles() is not None else "cublas"

    @property
    def shape(self):
        return tuple(self.a.shape)

    @property
    def local_rows(self):
        return self.a.shape[0]

    @property
    def T(self):
        return self._view(self.a.T, transposed=not self._transposed)

    # float32 matrices below the native threshold: "fp32" (default) = plain fp32 library GEMM; "tf32x3" = the same
    # error-compensated split on cuBLAS TF32 GEMMs (A/B reference for the tcgen05 kernel; slower than fp32 on cuBLAS).
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

    def _product(self, x, transpose: bool):
        """this view times x (transpose: its transpose times x)"""
        squeeze = x.ndim == 1
        if squeeze:
            x = x[:, None]
        tiles = self._native_tiles()
        if tiles is not None:
            rows = tiles.n if self._rows is None else self._rows
            out = tiles.matmul(x, transpose != self._transposed, rows)
        elif self.a.dtype == torch.float32:
            out = self._mm32(x, transpose)
        else:
            a = self.a.T if transpose else self.a
            out = (a @ x.to(self.a.dtype)).to(F64)
        return out[:, 0] if squeeze else out

    def dot_t(self, x):
        if x.shape[0] != self.a.shape[1]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(x.shape)))
        return self._product(x, False)

    def tdot_t(self, y):
        if y.shape[0] != self.a.shape[0]:
            raise ValueError("shapes {} and {} not aligned".format(self.shape[::-1], tuple(y.shape)))
        return self._product(y, True)

    def dot(self, x):
        return DeviceArray(self.dot_t(as_tensor(x, self.a.device)))

    def __getitem__(self, item):
        if isinstance(item, tuple):
            item = tuple(torch.as_tensor(i, device=self.a.device) if isinstance(i, np.ndarray) else i for i in item)
        elif isinstance(item, np.ndarray):
            item = torch.as_tensor(item, device=self.a.device)
        sub = self.a[item]
        sub = sub if sub.ndim == 2 else sub.reshape(1, -1)
        # a row prefix [0:m, :] of an untransposed matrix stays a view of the same tile stores (the cumulative
        # partitions of SuccessiveBatchedPowerMethod, iter_methods.py:557,573); anything else is a new plain matrix
        rows = None
        if isinstance(item, slice):
            rows = item
        elif isinstance(item, tuple) and len(item) == 2 and isinstance(item[0], slice) \
                and isinstance(item[1], slice) and item[1] == slice(None):
            rows = item[0]
        if rows is not None and not self._transposed:
            start, stop, step = rows.indices(self.a.shape[0])
            if start == 0 and step == 1:
                return self._view(sub, rows=stop)
        return self._plain(sub)

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
