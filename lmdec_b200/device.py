"""Device plumbing: torch owns HBM and streams, the C ABI does the arithmetic.

``DeviceArray`` is what the drop-in API returns where the reference returns a dask array: it exposes ``shape``,
``T``, slicing, ``compute()``/``persist()`` and ``__array__`` (reference usage: tests/test_iter_methods.py:81,
lmdec/decomp/iter_methods.py:408-417) while keeping the data resident on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi

F64 = torch.float64


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("lmdec_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class DeviceArray:
    """Thin array-like view of a CUDA tensor (float64 unless stated)."""

    __array_priority__ = 100

    def __init__(self, t: torch.Tensor):
        self.t = t

    # -- dask-like surface
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.ndim

    @property
    def dtype(self):
        return np.dtype(str(self.t.dtype).replace("torch.", ""))

    @property
    def T(self):
        return DeviceArray(self.t.T)

    def compute(self):
        return self.t.detach().cpu().numpy()

    def persist(self):
        return self

    def rechunk(self, *a, **k):
        return self

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, item):
        if isinstance(item, np.ndarray):
            item = torch.as_tensor(item, device=self.t.device)
        elif isinstance(item, tuple):
            item = tuple(torch.as_tensor(i, device=self.t.device) if isinstance(i, np.ndarray) else i for i in item)
        return DeviceArray(self.t[item])

    def dot(self, other):
        return DeviceArray(self.t @ as_tensor(other, self.t.device))

    # -- arithmetic used on S / U / V by callers
    def _bin(self, other, op):
        o = other.t if isinstance(other, DeviceArray) else other
        if isinstance(o, np.ndarray):
            o = torch.as_tensor(o, device=self.t.device)
        return DeviceArray(op(self.t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __neg__(self): return DeviceArray(-self.t)
    __radd__ = __add__
    __rmul__ = __mul__

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.t.dtype}, device={self.t.device})"


class LazyDeviceArray(DeviceArray):
    """DeviceArray whose tensor is produced on first use.

    The iteration stores (U_k, S_k, V_k) of every iterate in ``history.iter['last_value']`` (reference
    lmdec/decomp/iter_methods.py:399); materialising U_k = q U_R and uploading the K-sized host results each iteration
    would put pageable host->device copies -- which synchronise the stream -- into the hot loop.  They are deferred
    until somebody looks at them (finalisation, KeyboardInterrupt recovery, a user reading the history)."""

    def __init__(self, make, shape):
        self._make, self._t, self._shape = make, None, tuple(shape)

    @property
    def t(self):
        if self._t is None:
            self._t = self._make()
            self._make = None
        return self._t

    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return len(self._shape)


def as_tensor(x, device=None, dtype=F64) -> torch.Tensor:
    """Anything array-like -> CUDA tensor of ``dtype`` (no copy when already there)."""
    device = device or require_cuda()
    if isinstance(x, DeviceArray):
        x = x.t
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    a = np.asarray(x)
    if a.dtype == object:
        raise TypeError("cannot move an object array to the device")
    return torch.as_tensor(np.ascontiguousarray(a), device=device).to(dtype)


# ------------------------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI (all asynchronous on the current torch stream)
# ------------------------------------------------------------------------------------------------------------------
_WORK = {}


def _work(device, n: int) -> torch.Tensor:
    key = (device, "f64")
    w = _WORK.get(key)
    if w is None or w.numel() < n:
        w = torch.empty(max(n, 1 << 16), dtype=F64, device=device)
        _WORK[key] = w
    return w


def _rows(x: torch.Tensor) -> torch.Tensor:
    """row-major with unit inner stride (the kernels take a leading dimension)."""
    if x.ndim != 2:
        raise ValueError("expected a 2-D operand")
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x


GRAM_LIBRARY_MIN_K = 40


def gram(x: torch.Tensor, y: torch.Tensor = None, want_colsum: bool = False):
    """G = x^T x (or x^T y) [K,K] and optionally the column sums of x; deterministic."""
    lib = _capi.load()
    x = _rows(x)
    m, K = x.shape
    if K > GRAM_LIBRARY_MIN_K:
        # wide iterates (configs[4]: K = 110): the K x K Gram matrix is a plain library GEMM -- cuBLAS float64 runs it
        # 4.7x faster than the small-K reduction kernel (1.0 vs 4.8 ms at 1M x 110, tools/gram_time.py); the kernel's own
        # range ends at K = 128
        g = x.T @ (x if y is None else y)
        return (g, x.sum(0)) if want_colsum else g
    yy = None if y is None else _rows(y)
    G = torch.empty((K, K), dtype=F64, device=x.device)
    cs = torch.empty(K, dtype=F64, device=x.device) if want_colsum else None
    work = _work(x.device, lib.lmdec_gram_work_doubles(m, K))
    _capi.check(lib.lmdec_gram(_capi.ptr(x), x.stride(0), _capi.ptr(yy), 0 if yy is None else yy.stride(0), m, K,
                               _capi.ptr(G), _capi.ptr(cs), _capi.ptr(work), _capi.stream_ptr()), "lmdec_gram")
    return (G, cs) if want_colsum else G


def qvals(si: torch.Tensor, sj: torch.Tensor, scale: bool = True) -> torch.Tensor:
    lib = _capi.load()
    out = torch.empty(1, dtype=F64, device=si.device)
    si, sj = si.contiguous(), sj.contiguous()
    _capi.check(lib.lmdec_qvals(_capi.ptr(si), _capi.ptr(sj), si.numel(), int(bool(scale)), _capi.ptr(out),
                                _capi.stream_ptr()), "lmdec_qvals")
    return out


def sqdiff(a: torch.Tensor, b: torch.Tensor, s: torch.Tensor = None) -> torch.Tensor:
    """sum (a - b*s)^2, deterministic two-stage reduction."""
    lib = _capi.load()
    a, b = _rows(a), _rows(b)
    out = torch.empty(1, dtype=F64, device=a.device)
    work = _work(a.device, 1024)
    sc = None if s is None else s.contiguous()
    _capi.check(lib.lmdec_sqdiff(_capi.ptr(a), a.stride(0), _capi.ptr(b), b.stride(0), _capi.ptr(sc), a.shape[0],
                                 a.shape[1], _capi.ptr(out), _capi.ptr(work), _capi.stream_ptr()), "lmdec_sqdiff")
    return out


def chol_whiten(G: torch.Tensor, rank_tol: float = 1e-12):
    """Device K x K Cholesky whitening: returns (r_inv, r, flag) all on the device (flag int32[1])."""
    lib = _capi.load()
    K = G.shape[0]
    G = G.contiguous()
    r_inv = torch.empty((K, K), dtype=F64, device=G.device)
    r = torch.empty((K, K), dtype=F64, device=G.device)
    flag = torch.empty(1, dtype=torch.int32, device=G.device)
    _capi.check(lib.lmdec_chol_whiten(_capi.ptr(G), K, float(rank_tol), _capi.ptr(r_inv), _capi.ptr(r),
                                      _capi.ptr(flag), _capi.stream_ptr()), "lmdec_chol_whiten")
    return r_inv, r, flag


def gram_chol(x: torch.Tensor, rank_tol: float = 1e-12):
    """Gram matrix + Cholesky whitening of a LOCAL iterate in two launches: (G, r_inv, r, flag) on the device."""
    lib = _capi.load()
    x = _rows(x)
    m, K = x.shape
    dev = x.device
    G = torch.empty((K, K), dtype=F64, device=dev)
    r_inv = torch.empty((K, K), dtype=F64, device=dev)
    r = torch.empty((K, K), dtype=F64, device=dev)
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    work = _work(dev, lib.lmdec_gram_work_doubles(m, K))
    _capi.check(lib.lmdec_gram_chol(_capi.ptr(x), x.stride(0), m, K, float(rank_tol), _capi.ptr(G), _capi.ptr(r_inv),
                                    _capi.ptr(r), _capi.ptr(flag), _capi.ptr(work), _capi.stream_ptr()),
                "lmdec_gram_chol")
    return G, r_inv, r, flag


def qr_pack(r1, r2, g2, f1, f2) -> torch.Tensor:
    lib = _capi.load()
    K = r1.shape[0]
    packed = torch.empty(K * K + 3, dtype=F64, device=r1.device)
    _capi.check(lib.lmdec_qr_pack(_capi.ptr(r1), _capi.ptr(r2), _capi.ptr(g2), _capi.ptr(f1), _capi.ptr(f2), K,
                                  _capi.ptr(packed), _capi.stream_ptr()), "lmdec_qr_pack")
    return packed


def cholqr2_small_fits(n: int, K: int) -> bool:
    ctas = _capi.load().lmdec_cholqr2_small_max_elems() // 8192      # CTAs of the cluster, 8192 doubles of slice each
    return K <= 32 and ctas > 0 and -(-n // ctas) * K <= 8192


def cholqr2_small(x: torch.Tensor, rank_tol: float = 1e-12, stats_work: torch.Tensor = None, stats_lim: float = 0.0):
    """CholeskyQR2 of a small local iterate in one launch: (q, packed) with packed = [r | flag1 | flag2 | dev].
    `stats_work` / `stats_lim`: work buffer of the operator that multiplies A^T by q next (receives q's column
    statistics) and its fixed-point range 2^(8 planes - 2)."""
    lib = _capi.load()
    x = _rows(x)
    n, K = x.shape
    q = torch.empty((n, K), dtype=F64, device=x.device)
    packed = torch.empty(K * K + 3, dtype=F64, device=x.device)
    _capi.check(lib.lmdec_cholqr2_small(_capi.ptr(x), x.stride(0), n, K, float(rank_tol), _capi.ptr(q), K,
                                        _capi.ptr(packed), _capi.ptr(stats_work), float(stats_lim), _capi.stream_ptr()),
                "lmdec_cholqr2_small")
    return q, packed


def cholqr2_grid_fits(n: int, K: int) -> bool:
    lib = _capi.load()
    sms = lib.lmdec_cholqr2_grid_max_elems() // 16384
    if K > 32 or sms < 1:
        return False
    grid = min(sms, -(-n // 8))
    per = -(-(-(-n // grid)) // 8) * 8
    return per * K <= 16384


def cholqr2_grid(x: torch.Tensor, rank_tol: float = 1e-12, stats_work: torch.Tensor = None, stats_lim: float = 0.0):
    """CholeskyQR2 of a mid-size local iterate in one cooperative launch over all SMs: (q, packed) as cholqr2_small."""
    lib = _capi.load()
    x = _rows(x)
    n, K = x.shape
    q = torch.empty((n, K), dtype=F64, device=x.device)
    packed = torch.empty(K * K + 3, dtype=F64, device=x.device)
    gwork = _work(x.device, lib.lmdec_cholqr2_grid_work_doubles(K))
    _capi.check(lib.lmdec_cholqr2_grid(_capi.ptr(x), x.stride(0), n, K, float(rank_tol), _capi.ptr(q), K,
                                       _capi.ptr(packed), _capi.ptr(gwork), _capi.ptr(stats_work), float(stats_lim),
                                       _capi.stream_ptr()), "lmdec_cholqr2_grid")
    return q, packed
