"""Device-resident genotype storage and the lazily centred / scaled / mean-imputed operator over it.

Replaces the dask chunk graph the reference builds in ``make_snp_array`` (lmdec/decomp/utils.py:195-289):

    SNP = Chained[ Stacked[ Stacked[G(int8, NaN->0), M(sparse COO of column means at the NaNs)], -mean ], 1/std ]

Here G and the NaN positions live together in a 2-bit tile store (codes 0,1,2 and 3 = missing), held twice
(sample-major for ``A.dot``, SNP-major -- the .bed orientation -- for ``A.T.dot``), and the algebra of
``StackedArray.dot`` / ``ChainedArray.dot`` (lmdec/array/stacked.py:172-217, chained.py:169-177) is folded into the
small operands:

    A.dot(t)    = G (D t) + I (mu o D t) - 1 (mu^T D t)          I = missing indicator
    A.T.dot(y)  = D ( G^T y + mu o (I^T y) - mu (1^T y) )

The two big products run exactly (int64) on the tensor cores: the float64 operand is split into P signed-byte planes
(fixed point, 8P-2 bits below its column maximum), see ``csrc/geno_mma.cuh``.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import numpy as np
import torch

from . import _capi
from .device import F64, DeviceArray, as_tensor, require_cuda

_DTYPE_CODE = {torch.int8: 0, torch.uint8: 1, torch.float32: 2, torch.float64: 3}


_FUSED_STATS = os.environ.get("LMDEC_FUSED_STATS", "1") != "0"   # A/B switch: statistics of t = A^T q from the epilogue

# bench.py sets this to a list to time the dominant kernel with CUDA events on the launching stream
KERNEL_EVENTS = None


class _timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if KERNEL_EVENTS is not None:
            self.e1.record()
            KERNEL_EVENTS.append((self.tag, self.e0, self.e1))


def _round16(x: int) -> int:
    return (x + 15) // 16 * 16


class NotGenotypeError(ValueError):
    """the matrix holds values other than 0, 1, 2 and NaN (it cannot live in the 2-bit store)"""


class GenotypeStore:
    """n x p genotypes (rows = samples) as two 2-bit tile stores on one GPU.

    Two ways to shard over ranks (one process per GPU), never both:

    * ``group`` / ``n_global`` / ``row_offset``: samples (rows) are block-sharded; moments come from the all-reduced
      integer counts, so every rank holds identical mean / std; per iteration the p x K product A^T q is all-reduced.
    * ``snp_group`` / ``p_global`` / ``col_offset``: SNPs (columns) are block-sharded, every rank holds all samples.
      A^T q and everything per SNP stay local; per iteration only the n x K product A t is all-reduced -- the cheap
      exchange when n << p (2,504 x 1M, K = 20: 400 KB instead of 80 MB), and the iterate q is replicated, so its
      orthonormalisation needs no collective at all.
    """

    def __init__(self, n: int, p: int, device=None, group=None, n_global: Optional[int] = None, row_offset: int = 0,
                 snp_group=None, p_global: Optional[int] = None, col_offset: int = 0):
        if group is not None and snp_group is not None:
            raise ValueError("shard either the samples (group) or the SNPs (snp_group), not both")
        self.lib = _capi.load()
        self.ctx = _capi.Context()          # tunables + kernel timing of every product launched over this store
        self.device = device or require_cuda()
        self.n, self.p = int(n), int(p)
        self.group = group
        self.n_global = int(n_global) if n_global is not None else self.n
        self.row_offset = int(row_offset)
        self.snp_group = snp_group
        self.p_global = int(p_global) if p_global is not None else self.p
        self.col_offset = int(col_offset)
        self.s_store = torch.zeros(self.lib.lmdec_store_bytes(self.n, self.p), dtype=torch.uint8, device=self.device)
        self.v_store = torch.zeros(self.lib.lmdec_store_bytes(self.p, self.n), dtype=torch.uint8, device=self.device)
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._counts = None

    ndim = 2

    @property
    def shape(self):
        """global shape (samples over all row shards, SNPs over all SNP shards)"""
        return (self.n_global, self.p_global)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_dense(cls, array, device=None, rows_per_block: int = 8192, group=None, n_global=None, row_offset=0,
                   snp_group=None, p_global=None, col_offset=0):
        """Pack a dense {0,1,2,NaN} matrix (NumPy or torch; int8/uint8/float32/float64; integer missing = -1 or 3).

        Host arrays are streamed to the device in row blocks through a pinned staging buffer, so the dense matrix is
        never resident in HBM as a whole.
        """
        if isinstance(array, DeviceArray):
            array = array.t
        if not isinstance(array, torch.Tensor):
            array = np.asarray(array)
            if array.dtype not in (np.int8, np.uint8, np.float32, np.float64):
                array = array.astype(np.float64)
            array = torch.from_numpy(np.ascontiguousarray(array))
        if array.ndim != 2:
            raise ValueError("expected a 2-D genotype matrix")
        if array.dtype not in _DTYPE_CODE:
            array = array.to(torch.float64)
        n, p = array.shape
        self = cls(n, p, device=device, group=group, n_global=n_global, row_offset=row_offset, snp_group=snp_group,
                   p_global=p_global, col_offset=col_offset)
        rows_per_block = max(256, rows_per_block // 256 * 256)
        for r0 in range(0, n, rows_per_block):
            blk = array[r0:r0 + rows_per_block]
            if blk.device != self.device:
                blk = blk.contiguous().pin_memory().to(self.device, non_blocking=True) if blk.device.type == "cpu" \
                    else blk.to(self.device)
            self.pack_rows(blk, r0)
        self.finish()
        return self

    def pack_rows(self, block: torch.Tensor, r0: int):
        """Write rows [r0, r0+len(block)) (r0 a multiple of 256) of both stores from a device block."""
        if r0 % 256:
            raise ValueError("row blocks must start at a multiple of 256")
        if block.stride(1) != 1:
            block = block.contiguous()
        nb, p = block.shape
        code = _DTYPE_CODE[block.dtype]
        st = _capi.stream_ptr()
        ld = block.stride(0)
        _capi.check(self.lib.lmdec_pack_2bit(_capi.ptr(block), code, ld, 1, nb, p, r0, 0, _capi.ptr(self.s_store),
                                             self.n, self.p, _capi.ptr(self._bad), st), "pack (sample-major)")
        _capi.check(self.lib.lmdec_pack_2bit(_capi.ptr(block), code, 1, ld, p, nb, 0, r0, _capi.ptr(self.v_store),
                                             self.p, self.n, _capi.ptr(self._bad), st), "pack (SNP-major)")
        self._counts = None

    @classmethod
    def from_bed_bytes(cls, bed, n_samples: int, n_variants: int, device=None, variants_as_rows: bool = False):
        """PLINK .bed payload (variant-major, after the 3-byte magic) -> store, without a float round trip.
        Replaces pandas_plink.read_plink under load_plink_array (lmdec/array/io.py:16-73).  Code map (the one
        pandas_plink uses): 00 -> 0, 01 -> missing, 10 -> 1, 11 -> 2.  The store has samples as rows (n = n_samples,
        p = n_variants) unless `variants_as_rows` (the orientation ``read_plink`` itself returns)."""
        base = cls(n_variants, n_samples, device=device)          # rows = variants: the .bed orientation
        bed_t = bed if isinstance(bed, torch.Tensor) else torch.from_numpy(np.frombuffer(bed, dtype=np.uint8).copy())
        bed_t = bed_t.to(base.device).contiguous()
        if bed_t.numel() != n_variants * ((n_samples + 3) // 4):
            raise ValueError("bed payload size does not match n_variants x ceil(n_samples/4)")
        st = _capi.stream_ptr()
        _capi.check(base.lib.lmdec_recode_bed(_capi.ptr(bed_t), n_variants, n_samples, _capi.ptr(base.s_store), st))
        base.derive_snp_major()
        base.finish()
        return base if variants_as_rows else base.transposed()

    def derive_snp_major(self):
        """column-major store (M = p) from the row-major one (M = n): lmdec_store_transpose, codes untouched"""
        _capi.check(self.lib.lmdec_store_transpose(_capi.ptr(self.s_store), self.n, self.p, _capi.ptr(self.v_store),
                                                   _capi.stream_ptr()), "store_transpose")
        self._counts = None
        return self

    def transposed(self) -> "GenotypeStore":
        """the p x n matrix over the SAME two tile stores (they swap roles; nothing is copied)"""
        if self.group is not None or self.snp_group is not None:
            raise NotImplementedError("transposed() of a sharded store")
        t = GenotypeStore.__new__(GenotypeStore)
        t.__dict__.update(self.__dict__)
        t.n, t.p, t.n_global, t.p_global = self.p, self.n, self.p, self.n
        t.s_store, t.v_store = self.v_store, self.s_store
        t._counts = None
        t._bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        return t

    # ------------------------------------------------------------------ persistence (reference array/io.py:129-273)
    def save(self, path: str):
        """Raw tile stores + shape (no moments): see ``lmdec_b200.array.io.save_array`` for the operator tree that also
        persists mean, 1/std, the imputation vector and the integer code counts."""
        import os
        os.makedirs(path, exist_ok=True)
        torch.save({"n": self.n, "p": self.p, "n_global": self.n_global, "row_offset": self.row_offset,
                    "format": "lmdec_b200 tile store v1 (2-bit dosage codes, 3 = missing; 128 x 256 tiles)"},
                   os.path.join(path, "meta.pt"))
        np.save(os.path.join(path, "sample_major.npy"), self.s_store.cpu().numpy())
        np.save(os.path.join(path, "snp_major.npy"), self.v_store.cpu().numpy())

    @classmethod
    def load(cls, path: str, device=None, group=None):
        import os
        meta = torch.load(os.path.join(path, "meta.pt"))
        self = cls(meta["n"], meta["p"], device=device, group=group, n_global=meta["n_global"],
                   row_offset=meta["row_offset"])
        for name, dst in (("sample_major.npy", self.s_store), ("snp_major.npy", self.v_store)):
            src = torch.from_numpy(np.load(os.path.join(path, name), mmap_mode="r")[:])
            if src.numel() != dst.numel():
                raise ValueError("stored tile store does not match its recorded shape")
            dst.copy_(src)
        return self

    def finish(self):
        bad = int(self._bad.item())
        if bad:
            raise NotGenotypeError("genotype matrix holds values other than 0, 1, 2 and NaN")
        return self

    # ------------------------------------------------------------------ exact integer content
    def code_counts(self, rows: Optional[int] = None) -> torch.Tensor:
        """int64 [p, 4]: per SNP the number of samples with dosage 0, 1, 2 and missing (all ranks summed)."""
        full = rows is None
        if full and self._counts is not None:
            return self._counts
        lim = self.n if full else int(rows)
        cnt = torch.empty((self.p, 4), dtype=torch.int64, device=self.device)
        _capi.check(self.lib.lmdec_code_counts(_capi.ptr(self.v_store), self.p, self.n, lim, _capi.ptr(cnt),
                                               _capi.stream_ptr()), "code_counts")
        if self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(cnt, group=self.group)
        if full:
            self._counts = cnt
        return cnt

    @property
    def has_missing(self) -> bool:
        return bool((self.code_counts()[:, 3].sum() > 0).item())

    def to_dense(self, transpose: bool = False, rows: Optional[int] = None) -> torch.Tensor:
        """int8 [n, p] (or [p, n]) with -1 at missing: the bit-exact decode used by the parity tests.
        `rows` limits the decode to a leading block of store rows."""
        M, Kd, st = (self.p, self.n, self.v_store) if transpose else (self.n, self.p, self.s_store)
        if rows is not None:
            M = min(M, int(rows))
        out = torch.empty((M, Kd), dtype=torch.int8, device=self.device)
        _capi.check(self.lib.lmdec_unpack_2bit(_capi.ptr(st), M, Kd, _capi.ptr(out), Kd, _capi.stream_ptr()))
        return out

    def moments(self, std_method: str = "binom"):
        """(mean, std) float64 [p] from the integer counts: get_array_moments (lmdec/decomp/utils.py:64-82)."""
        c = self.code_counts().to(F64)
        cnt = c[:, 0] + c[:, 1] + c[:, 2]
        mean = (c[:, 1] + 2.0 * c[:, 2]) / cnt
        if std_method == "binom":
            u = mean / 2
            std = torch.sqrt(2 * u * (1 - u))
        elif std_method in ("norm", "normal"):
            var = (c[:, 0] * mean ** 2 + c[:, 1] * (1 - mean) ** 2 + c[:, 2] * (2 - mean) ** 2) / cnt
            std = torch.sqrt(var)
        else:
            raise NotImplementedError(f"std_method, {std_method}, is not implemented ")
        return mean, std


class GenotypeOperator:
    """The operator ``make_snp_array`` returns: ((G + M) - 1 mean^T) diag(1/std), evaluated on the tensor cores.

    Operator protocol used by the decomposition drivers: ``shape``, ``dot``, ``T.dot``, row-prefix
    slicing ``[0:m, :]`` (moments are NOT refit on a prefix, as in the reference: lmdec/array/chained.py:235-241,
    stacked.py:254-265), fancy row indexing raises ``NotImplementedError`` like ``ChainedArray.__getitem__``.
    """

    def __init__(self, store: GenotypeStore, mean: Optional[torch.Tensor], inv_std: Optional[torch.Tensor],
                 planes: int = 3, rows: Optional[int] = None, impute: bool = True):
        self.store = store
        self.mean = mean            # centring vector (None = no centring)
        self.inv_std = inv_std      # scaling vector (None = no scaling)
        self.planes = int(planes)
        self.rows = store.n if rows is None else int(rows)   # local row prefix
        # global number of rows of this (view of the) operator: identical on every rank by construction, never derived
        # from rank-local quantities (a prefix view clamps differently on every rank of a row-sharded store)
        self.global_rows = store.n_global if rows is None or store.group is None else None
        if rows is not None and store.group is None:
            self.global_rows = self.rows
        self.has_missing = store.has_missing
        if self.has_missing and not impute:
            raise ValueError("the store holds missing genotypes; they are always mean-imputed")
        self.impute_mean = store.moments()[0] if self.has_missing else None
        self._t = None
        self._buf = {}

    # ------------------------------------------------------------------ protocol
    @property
    def shape(self):
        """global shape: samples over all row shards x SNPs over all SNP shards"""
        if self.store.group is None:
            return (self.rows, self.store.p_global)
        if self.global_rows is None:
            raise ValueError("row-sharded operator built with an explicit local prefix: set `global_rows`")
        return (self.global_rows, self.store.p_global)

    @property
    def local_cols(self):
        return self.store.p

    @property
    def local_rows(self):
        return self.rows

    ndim = 2

    @property
    def T(self):
        if self._t is None:
            self._t = _Transposed(self)
        return self._t

    def dot(self, x):
        return DeviceArray(self.dot_t(as_tensor(x, self.store.device)))

    def persist(self):
        return self

    def rechunk(self, *a, **k):
        return self

    def __getitem__(self, item):
        if not isinstance(item, tuple):
            item = (item, slice(None))
        if len(item) != 2 or item[1] != slice(None, None, None):
            raise NotImplementedError("only row selection [rows, :] is supported")
        r = item[0]
        if not isinstance(r, slice):
            raise NotImplementedError("Only implements slice indexing")
        start, stop, step = r.indices(self.shape[0])
        if start != 0 or step != 1:
            raise NotImplementedError("only row prefixes [0:m, :] are views of the store")
        global_stop = stop
        if self.store.group is not None:
            # block-contiguous shards: the global prefix [0, stop) is this rank's [0, clamp(stop - offset)]
            stop = min(max(stop - self.store.row_offset, 0), self.store.n)
        op = GenotypeOperator.__new__(GenotypeOperator)
        op.__dict__.update(self.__dict__)
        op.rows, op._t, op._buf, op.global_rows = stop, None, self._buf, global_stop
        return op

    # ------------------------------------------------------------------ buffers
    def _buffers(self, K):
        P = self.planes
        N = _round16(K * P)
        if N > 128:
            raise ValueError(f"{K} operand columns with {P} planes need N = {N} > 128 tensor-core columns: "
                             "dot_t / tdot_t split wider operands into column blocks")
        key = (K, P)
        b = self._buf.get(key)
        if b is None:
            st, dev = self.store, self.store.device
            lib = st.lib
            miss = self.has_missing
            nb_dot, nb_tdot = lib.lmdec_image_bytes(st.p, N), lib.lmdec_image_bytes(st.n, N)
            b = {
                "N": N,
                "work": torch.zeros(lib.lmdec_apply_work_doubles(K), dtype=F64, device=dev),   # (ticket slot = 0)
                "im1": torch.empty(max(nb_dot, nb_tdot), dtype=torch.int8, device=dev),
                "im2": torch.empty(nb_dot, dtype=torch.int8, device=dev) if miss else None,
                "acc0": torch.empty((max(st.n, st.p), K), dtype=torch.int64, device=dev),
                "acc1": torch.empty((st.p, K), dtype=torch.int64, device=dev) if miss else None,
            }
            self._buf[key] = b
        return b

    def _max_cols(self) -> int:
        """operand columns one product launch takes: K * planes <= 128 tensor-core columns (N of the MMA); wider
        operands (any k + buffer <= min(n, p) is legal in the reference, iter_methods.py:351-355) run as column blocks"""
        return 128 // self.planes

    def _center_w(self):
        if self.mean is None:
            return None
        if getattr(self, "_cw", None) is None:
            self._cw = (self.mean if self.inv_std is None else self.mean * self.inv_std).contiguous()
        return self._cw

    # ------------------------------------------------------------------ the two products (torch tensors in/out)
    def _apply(self, b_in: torch.Tensor, transpose: bool, out: torch.Tensor = None, m_range=None,
               reuse_operand: bool = False, out_dtype=F64, emit_stats: bool = False) -> torch.Tensor:
        """One lmdec_operator_apply call; `m_range` = (m0, m1) restricts the OUTPUT rows (m0 a multiple of 128).
        b_in may be float64 or float32; the result is `out_dtype` (float64 or float32).  `emit_stats` (A^T y): the
        product also leaves the column statistics of its result in the work buffer; an A.dot of exactly that tensor
        (same object, not modified since) then skips the statistics pass over its p x K operand."""
        st, lib = self.store, self.store.lib
        K = b_in.shape[1]
        buf = self._buffers(K)
        m = self.rows
        if transpose:
            store, M, Kd, m_lim, kd_lim = st.v_store, st.p, st.n, st.p, m
        else:
            store, M, Kd, m_lim, kd_lim = st.s_store, st.n, st.p, m, st.p
        if out is None:
            out = torch.empty((m_lim, K), dtype=out_dtype, device=st.device)
        flags = int(transpose) | (2 if reuse_operand else 0) | (4 if b_in.dtype == torch.float32 else 0) | \
            (8 if out.dtype == torch.float32 else 0)
        emitted = None if reuse_operand else buf.pop("emitted", None)   # any other product reuses the work buffer
        if m == 0:
            return out.zero_() if transpose else out
        if emitted is not None and m_range is None:
            kind, ref, version = emitted         # statistics of b_in already in the work buffer?
            if kind == ("q" if transpose else "t") and ref() is b_in and b_in._version == version \
                    and b_in.shape[0] == kd_lim:
                flags |= 32
        if transpose and emit_stats and m_range is None and _FUSED_STATS:
            flags |= 16
        m0, m1 = (0, m_lim) if m_range is None else m_range
        if m0 % 128 or not (0 <= m0 < m1 <= m_lim):
            raise ValueError("bad output row range")
        miss = self.has_missing

        def off(t, per_row_bytes=8):
            return None if t is None else C.c_void_p(t.data_ptr() + m0 * per_row_bytes)

        row_vec = transpose  # scale / impute / centre vectors are per OUTPUT row only for A^T y
        store_ptr = C.c_void_p(store.data_ptr() + (m0 // 128) * ((Kd + 255) // 256) * 8192)
        with _timed("tdot" if transpose else "dot"):
            _capi.check(lib.lmdec_operator_apply(
                st.ctx.handle, store_ptr, M - m0, Kd, m1 - m0, kd_lim, int(miss), flags, _capi.ptr(b_in), b_in.stride(0),
                K, self.planes,
                off(self.inv_std) if row_vec else _capi.ptr(self.inv_std),
                (off(self.impute_mean) if row_vec else _capi.ptr(self.impute_mean)) if miss else None,
                off(self._center_w()) if row_vec else _capi.ptr(self._center_w()),
                _capi.ptr(buf["work"]), _capi.ptr(buf["im1"]), _capi.ptr(buf["im2"]),
                _capi.ptr(buf["acc0"]), _capi.ptr(buf["acc1"]), off(out, K * out.element_size()), K,
                _capi.stream_ptr()),
                "lmdec_operator_apply")
        if flags & 16:
            buf["emitted"] = ("t", weakref.ref(out), out._version)
        return out

    def qr_stats_slot(self, rows: int, K: int):
        """(work buffer, fixed-point range) for the column statistics of the iterate q [rows, K] that the
        orthonormalisation kernel (lmdec_cholqr2_small) can leave for the following A^T q; (None, 0) when it cannot."""
        if not _FUSED_STATS or self.store.group is not None or rows != self.rows or _round16(K * self.planes) > 128:
            return None, 0.0
        buf = self._buffers(K)
        buf.pop("emitted", None)
        return buf["work"], float(2 ** (8 * self.planes - 2))

    def qr_stats_emitted(self, q: torch.Tensor):
        self._buffers(q.shape[1])["emitted"] = ("q", weakref.ref(q), q._version)

    def apass_t(self, q: torch.Tensor, inv_factor: float = 1.0):
        """(t, x) = (A^T q, A t * inv_factor) in ONE C call (`lmdec_apass`): the step of the power iteration.
        Returns None when this operator cannot take the fused call (row-sharded samples, a row-range view, wide K)."""
        st, lib = self.store, self.store.lib
        if st.group is not None or q.ndim != 2 or q.shape[0] != self.rows or q.shape[1] > self._max_cols() \
                or self.rows == 0 or q.dtype not in (F64, torch.float32) or not _FUSED_STATS:
            return None
        if q.stride(1) != 1 or q.stride(0) != q.shape[1]:
            q = q.contiguous()
        K = q.shape[1]
        buf = self._buffers(K)
        emitted = buf.pop("emitted", None)
        q_flags = 2 if q.dtype == torch.float32 else 0
        if emitted is not None:
            kind, ref, version = emitted
            if kind == "q" and ref() is q and q._version == version:
                q_flags |= 1              # the orthonormalisation kernel left q's column statistics in the work buffer
        t = torch.empty((st.p, K), dtype=torch.float32, device=st.device)
        x = torch.empty((self.rows, K), dtype=F64, device=st.device)
        miss = self.has_missing
        with _timed("apass"):
            _capi.check(lib.lmdec_apass(
                st.ctx.handle, _capi.ptr(st.s_store), _capi.ptr(st.v_store), st.n, st.p, self.rows, int(miss),
                _capi.ptr(q), K, q_flags, K, self.planes, _capi.ptr(self.inv_std),
                _capi.ptr(self.impute_mean) if miss else None, _capi.ptr(self._center_w()),
                _capi.ptr(buf["work"]), _capi.ptr(buf["im1"]), _capi.ptr(buf["im1"]), _capi.ptr(buf["im2"]),
                _capi.ptr(buf["acc0"]), _capi.ptr(buf["acc1"]), _capi.ptr(t), _capi.ptr(x), K, float(inv_factor),
                _capi.stream_ptr()), "lmdec_apass")
        if st.snp_group is not None:      # SNP shards: the n x K partial products are the only exchange
            import torch.distributed as dist
            dist.all_reduce(x, group=st.snp_group)
        return t, x

    def dot_t(self, t: torch.Tensor) -> torch.Tensor:
        """x[n_local, K] = A t  for t [p, K] float64 on the device."""
        squeeze = t.ndim == 1
        if squeeze:
            t = t[:, None]
        if t.shape[0] != self.store.p:
            raise ValueError("shapes {} and {} not aligned".format(self.shape, tuple(t.shape)))
        if t.dtype not in (F64, torch.float32):
            t = t.to(F64)
        kc = self._max_cols()
        if t.shape[1] > kc:
            # more operand columns than one tensor-core N tile holds (K * planes > 128): column blocks, one product each
            x = torch.cat([self.dot_t(t[:, k0:k0 + kc].contiguous()) for k0 in range(0, t.shape[1], kc)], dim=1)
            return x
        if t.stride(1) != 1 or t.stride(0) != t.shape[1]:
            t = t.contiguous()
        x = self._apply(t, False)
        if self.store.snp_group is not None:      # SNP shards: the n x K partial products are the only exchange
            import torch.distributed as dist
            dist.all_reduce(x, group=self.store.snp_group)
        return x[:, 0] if squeeze else x

    @property
    def wide_dtype(self):
        """dtype of the p x K side of the iteration (t = A^T q) on the internal fast path; the public `.T.dot` is
        always float64.  float32: its 24-bit mantissa is finer than the 22-bit fixed point t is quantised to on its next
        use, the product's epilogue writes half the bytes and the operand kernel of the next product reads half (single
        GPU 0.752 -> 0.735 ms per iteration), and when row-sharded the all-reduce moves half the NVLink bytes (2 x B200:
        1.41 -> 1.28 ms).  `LMDEC_WIDE_DTYPE=f64` keeps the wide side in float64."""
        env = os.environ.get("LMDEC_WIDE_DTYPE")
        if env:
            return {"f32": torch.float32, "f64": torch.float64}[env]
        return torch.float32

    def tdot_t(self, y: torch.Tensor, out_dtype=None) -> torch.Tensor:
        """w[p, K] = A^T y  for y [n_local, K] on the device (all-reduced over ranks when sharded); `out_dtype`
        defaults to `wide_dtype`."""
        out_dtype = out_dtype or self.wide_dtype
        squeeze = y.ndim == 1
        if squeeze:
            y = y[:, None]
        if y.shape[0] != self.rows:
            raise ValueError("shapes {} and {} not aligned".format(self.shape[::-1], tuple(y.shape)))
        if y.dtype not in (F64, torch.float32):
            y = y.to(F64)
        kc = self._max_cols()
        if y.shape[1] > kc:
            return torch.cat([self.tdot_t(y[:, k0:k0 + kc].contiguous(), out_dtype)
                              for k0 in range(0, y.shape[1], kc)], dim=1)
        if y.stride(1) != 1 or y.stride(0) != y.shape[1]:
            y = y.contiguous()
        if self.store.group is None:
            w = self._apply(y, True, out_dtype=out_dtype, emit_stats=True)
        else:
            w = self._tdot_sharded(y, out_dtype)
        return w[:, 0] if squeeze else w

    # ------------------------------------------------------------------ row-sharded A^T y
    def _symm_ring(self, K: int):
        """Three p x K float32 buffers in ONE symmetric allocation shared with the other ranks of the row group (torch
        symmetric memory: peer-mapped, with an NVSwitch multicast alias), or None when that is not available.  The
        buffers rotate: a product adds into one while the previous result is still being read and the next is zeroed."""
        ring = self._buf.get(("symm", K), False)
        if ring is not False:
            return ring
        ring = None
        st = self.store
        if os.environ.get("LMDEC_FUSED_REDUCE", "1") != "0" and K % 4 == 0:
            import torch.distributed as dist
            buf = hdl = None
            n_el = ((st.p * K + 31) // 32) * 32
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty((3 * n_el,), dtype=torch.float32, device=st.device)
                hdl = symm_mem.rendezvous(buf, st.group)
                mine = 1 if hdl.multicast_ptr else 0
            except Exception:           # no symmetric memory / multicast on this box: NCCL all-reduce
                if os.environ.get("LMDEC_FUSED_REDUCE") == "require":
                    raise
                mine = 0
            ok = torch.tensor([mine], device=st.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=st.group)      # every rank takes the same path
            if int(ok.item()):
                buf.zero_()
                hdl.barrier(channel=0)
                ring = {"buf": buf, "hdl": hdl, "n_el": n_el, "pos": 0}
        self._buf[("symm", K)] = ring
        return ring

    def _tdot_fused_reduce(self, y: torch.Tensor):
        """A^T y with the all-reduce over the row shards fused into the kernel's epilogue (lmdec_operator_apply_reduce);
        None when the fused path is not available.  The result is a view of the symmetric ring: it stays valid until the
        next-but-one product of this operator (the iteration consumes t before that; anything kept longer is copied)."""
        st, lib = self.store, self.store.lib
        K = y.shape[1]
        if self._buf.get(("symm_off", K)):
            return None
        ring = self._symm_ring(K)
        if ring is None:
            return None
        buf, hdl, n_el = ring["buf"], ring["hdl"], ring["n_el"]
        cur = ring["pos"]
        nxt = (cur + 1) % 3
        out = buf[cur * n_el: cur * n_el + st.p * K].view(st.p, K)
        b = self._buffers(K)
        b.pop("emitted", None)
        miss = self.has_missing
        m = self.rows
        if m > 0:
            flags = 1 | 8 | (4 if y.dtype == torch.float32 else 0)
            mc = C.c_void_p(hdl.multicast_ptr + cur * n_el * 4)
            with _timed("tdot"):
                rc = lib.lmdec_operator_apply_reduce(
                    st.ctx.handle, _capi.ptr(st.v_store), st.p, st.n, st.p, m, int(miss), flags, _capi.ptr(y), y.stride(0),
                    K, self.planes, _capi.ptr(self.inv_std), _capi.ptr(self.impute_mean) if miss else None,
                    _capi.ptr(self._center_w()), _capi.ptr(b["work"]), _capi.ptr(b["im1"]), _capi.ptr(b["im2"]),
                    _capi.ptr(b["acc0"]), _capi.ptr(b["acc1"]), _capi.ptr(out), mc, _capi.stream_ptr())
            if rc == 3:          # cannot be fused for this shape: the same decision on every rank (same shapes), NCCL path
                self._buf[("symm_off", K)] = True
                return None
            _capi.check(rc, "lmdec_operator_apply_reduce")
        # zero the buffer the NEXT product adds into, then meet the other ranks: behind the barrier every rank's
        # additions to `out` have landed here and every rank's next buffer is clean
        buf[nxt * n_el:(nxt + 1) * n_el].zero_()
        hdl.barrier(channel=0)
        ring["pos"] = nxt
        return out

    def _tdot_sharded(self, y: torch.Tensor, out_dtype) -> torch.Tensor:
        """A^T y over row-sharded samples.  float32 results take the fused path when the box offers NVSwitch multicast:
        the kernel's epilogue adds each finished tile into every rank's copy of the result (`_tdot_fused_reduce`), one
        signal-pad barrier follows, no collective.  Otherwise the p x K partial products are all-reduced over NCCL in
        place.  `reduce_parts` > 1 (NCCL path) splits the product into SNP ranges and launches each range's all-reduce
        asynchronously; measured on 2 and 8 x B200 the extra launches cost as much as the overlap wins
        (profiles/r01/multi_gpu_notes.txt, profiles/r02/target_8gpu_parts2.json), so the default is 1."""
        import torch.distributed as dist
        st = self.store
        K = y.shape[1]
        if out_dtype == torch.float32:
            w = self._tdot_fused_reduce(y)
            if w is not None:
                return w
        w = torch.empty((st.p, K), dtype=out_dtype, device=st.device)
        n_tiles = (st.p + 127) // 128
        parts = max(1, min(int(getattr(st, "reduce_parts", 0) or os.environ.get("LMDEC_REDUCE_PARTS", "1")),
                           n_tiles))
        bounds = [min(st.p, (n_tiles * i // parts) * 128) for i in range(parts)] + [st.p]
        pending = []
        if parts > 1:
            st.ctx.set(_capi.KNOB_SM_RESERVE, int(os.environ.get("LMDEC_SM_RESERVE", "16")))
        for i in range(parts):
            m0, m1 = bounds[i], bounds[i + 1]
            if m1 <= m0:
                continue
            self._apply(y, True, out=w, m_range=(m0, m1), reuse_operand=bool(pending))
            pending.append(dist.all_reduce(w[m0:m1], group=st.group, async_op=True))
        st.ctx.set(_capi.KNOB_SM_RESERVE, 0)
        for work in pending:
            work.wait()
        return w

    # ------------------------------------------------------------------ materialisation (tests / small inputs)
    def to_dense(self) -> torch.Tensor:
        g = self.store.to_dense(rows=self.rows).to(F64)
        if self.has_missing:
            g = torch.where(g < 0, self.impute_mean.expand_as(g), g)
        if self.mean is not None:
            g = g - self.mean
        if self.inv_std is not None:
            g = g * self.inv_std
        return g

    def compute(self):
        return self.to_dense().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.compute()
        return a.astype(dtype) if dtype is not None else a


class _Transposed:
    def __init__(self, op: GenotypeOperator):
        self.op = op

    ndim = 2

    @property
    def shape(self):
        return self.op.shape[::-1]

    @property
    def T(self):
        return self.op

    def dot(self, y):
        return DeviceArray(self.op.tdot_t(as_tensor(y, self.op.store.device), out_dtype=F64))

    def dot_t(self, y):
        return self.op.tdot_t(y)

    def tdot_t(self, t):
        return self.op.dot_t(t)

    def compute(self):
        return self.op.compute().T

    def __array__(self, dtype=None, copy=None):
        return self.op.__array__(dtype).T
