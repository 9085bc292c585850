"""Block power iteration drivers with the reference's API, defaults, history layout and tolerance semantics
(reference lmdec/decomp/iter_methods.py:21-600).

Host control flow only: the loop, tolerances, history, KeyboardInterrupt recovery and time limit are Python as in the
reference; every product with the data matrix is a C-ABI kernel call (``GenotypeOperator``) or a library GEMM (dense
inputs), the orthonormalisation is CholeskyQR2 (``array/linalg.py``), and iterates stay on the GPU.

One step of the reference,  q = tsqr(x).Q ;  x <- A (A^T q) / factor  (iter_methods.py:381-386), is evaluated as
``t = A^T q`` (all-reduced over ranks when samples are row-sharded) and ``x = A t``.  ``t`` and the next iterate are
cached per iterate so that the 'v-subspace' and 'rmse' metrics reuse them instead of making the extra passes over A the
reference makes (matrix_ops.py:93, metrics.py:157): A^T x = t r, and A A^T u = (A t) u_r for u = q u_r.
"""
from __future__ import annotations

import time
import warnings
from typing import List, Optional, Union

import numpy as np
import torch
from tqdm import tqdm

from ..array.linalg import orthonormalize, orthonormalize_begin, tsqr_svd
from ..array.matrix_ops import subspace_to_V
from ..array.metrics import _rmse_from_products, q_value_converge, subspace_dist
from ..array.operators import as_operator
from ..array.random import array_constant_partition, cumulative_partition
from ..device import F64, DeviceArray, LazyDeviceArray, as_tensor
from .init_methods import rnormal_start, sub_svd_init, sym_mat_mult_start
from .utils import PowerMethodSummary


def recover_last_value(f):
    """KeyboardInterrupt inside svd() returns history.iter['last_value'] when one exists
    (lmdec/array/wrappers/recover_iter.py:5-18)."""
    from functools import wraps

    @wraps(f)
    def wrapped(self, *args, **kwargs):
        try:
            return f(self, *args, **kwargs)
        except KeyboardInterrupt:
            if self.history is not None and self.history.iter["last_value"]:
                return self.history.iter["last_value"]
            raise
    return wrapped


def _group_of(op):
    return getattr(getattr(op, "store", None), "group", None) or getattr(op, "group", None)


def _snp_group_of(op):
    """process group over which the SNP (p) axis is sharded, i.e. over which p x K matrices are row-sharded"""
    return getattr(getattr(op, "store", None), "snp_group", None)


def _tdot(op, y):
    return op.tdot_t(y) if hasattr(op, "tdot_t") else op.T.dot_t(y)


class _IterAlgo:
    def __init__(self, max_iter: Optional[int] = None, scoring_method: Optional[Union[List[str], str]] = None,
                 tol=None, factor=None, time_limit: Optional[int] = None, warn: Optional[bool] = None):
        if max_iter is None:
            self.max_iter = 50
        else:
            if max_iter <= 0:
                raise ValueError("max_iter must be a postitive integer")
            if int(max_iter) != max_iter:
                raise ValueError("max_iter must be a integer")
            self.max_iter = max_iter
        if scoring_method is None:
            scoring_method = "q-vals"
        if time_limit is None:
            self.time_limit = 1000
        else:
            if time_limit <= 0:
                raise ValueError("time_limit must be a positive amount")
            self.time_limit = time_limit
        if tol is None:
            tol = .1e-3
        if not isinstance(tol, list):
            tol = [tol]
        if not isinstance(scoring_method, list):
            scoring_method = [scoring_method]
        if len(scoring_method) != len(tol):
            raise ValueError("tolerance, tol, specification must match with convergence criteria, scoring_method,"
                             " specification. There is no one to one mapping between \n"
                             " {} and {}".format(tol, scoring_method))
        if any(x not in ["q-vals", "rmse", "v-subspace"] for x in scoring_method):
            raise ValueError("Must use scoring method in {}".format(["q-vals", "rmse", "v-subspace"]))
        self.factor = factor
        self.warn = False if warn is None else warn
        self.tol = tol
        self.scoring_method = scoring_method
        self.history = None
        self.num_iter = None
        self.array = None

    def _reset_history(self):
        self.history = PowerMethodSummary(time={"start": None, "stop": None, "iter": [], "step": [], "acc": []},
                                          acc={"q-vals": [], "rmse": [], "v-subspace": []},
                                          iter={"U": [], "S": [], "V": [], "last_value": []})

    @property
    def time(self):
        return self.history.time["stop"] - self.history.time["start"]

    def _parameter_step(self, x, **kwargs):
        pass

    @recover_last_value
    def svd(self, array, verbose: bool = True, return_history: bool = False, **kwargs):
        """k-truncated SVD of `array`: returns (U [n,k], S [k], V [k,p]) as device arrays, or the history namedtuple."""
        self._reset_history()
        self.history.time["start"] = time.time()
        x_k = self._initialization(array, **kwargs)
        converged = False
        for i in tqdm(range(1, self.max_iter + 1), disable=not verbose):
            self.num_iter = i
            iter_start = time.time()
            acc_list = self._solution_accuracy(x_k, **kwargs)
            self.history.time["acc"].append(time.time() - iter_start)
            for method, acc, tol in zip(self.scoring_method, acc_list, self.tol):
                self.history.acc[method].append(acc)
                if acc <= tol:
                    converged = True
            if converged:
                break
            self._parameter_step(x_k, **kwargs)
            step_start = time.time()
            x_k = self._solution_step(x_k, **kwargs)
            self.history.time["step"].append(time.time() - step_start)
            self.history.time["iter"].append(time.time() - iter_start)
            if time.time() - self.history.time["start"] > self.time_limit:
                break
        result = self._finalization(x_k, return_history=return_history, **kwargs)
        self.history.time["stop"] = time.time()
        if self.warn and not converged:
            warnings.warn("Did not converge. \n"
                          "Time Usage : {0:.2f}s of {1}s (Time Limit) \n"
                          "Iteration Usage : {2} of {3} (Iteration Limit)"
                          .format(self.time, self.time_limit, self.num_iter, self.max_iter))
        return result


class PowerMethod(_IterAlgo):
    """Truncated SVD by block power iteration on A A^T (iter_methods.py:221-433).

    Extra keyword arguments of this build: ``svd(..., x0=)`` injects the start matrix (parity runs; the reference's own
    starts depend on dask's RNG), ``seed=`` seeds the ``lmbd`` noise.  The legacy ``scale=`` / ``center=`` constructor
    arguments wrap a plain matrix in a ``ScaledCenterArray`` (notebook API); ``svd`` accepts and ignores the legacy
    ``persist= dtype= transpose= mask_nan= mask_fill=``.
    """

    def __init__(self, k: int = 10, max_iter: int = 50, scoring_method="q-vals", tol=.1, factor: Optional[str] = "n",
                 lmbd=.01, buffer: int = 10, sub_svd_start=True, init_row_sampling_factor: int = 5,
                 time_limit: int = 1000, warn: bool = False, scale=None, center=None, std_dist=None):
        super().__init__(max_iter=max_iter, scoring_method=scoring_method, tol=tol, factor=factor,
                         time_limit=time_limit, warn=warn)
        if int(k) <= 0:
            raise ValueError("k must be a positive integer")
        if int(buffer) < 0:
            raise ValueError("buffer must be a non-negative integer")
        self.k = int(k)
        self.buffer = int(buffer)
        self.lmbd = lmbd
        self.sub_svd_start = sub_svd_start
        if init_row_sampling_factor <= 0:
            raise ValueError("init_row_sampling_factor must be a positive value")
        self.init_row_sampling_factor = init_row_sampling_factor
        self.compute = True
        self.scale, self.center, self.std_dist = scale, center, std_dist
        self.scaled_array = None
        self.speculate = True
        self._cache = {}

    def project(self, x, onto, scale_center_x: bool = True):
        """Project new samples `x` [N, P] onto `onto` ([P, M], or [M, P] as the reference's docstring has it):
        returns x' onto [N, M], where x' is `x` centred / scaled / mean-imputed with the moments of the matrix the last
        ``svd`` ran on (``scale_center_x``), i.e. the principal-component scores of unseen samples.

        The reference documents this method and then raises NotImplementedError (iter_methods.py:321-348); its xfail
        test (tests/test_iter_methods.py:448-474) states the contract implemented here:
        ``PM.project(y, onto=V_k.T) == ((y - mu) / std) @ V_k.T``.  A genotype `x` ({0,1,2,NaN}) is packed into a
        2-bit store and the product is ONE ``lmdec_operator_apply`` with the stored moments (NaN -> the TRAINING mean of
        the SNP, which centres to zero); any other `x` is a dense float64 product."""
        from ..array.scaled_legacy import ScaledCenterArray
        from ..store import GenotypeOperator, GenotypeStore, NotGenotypeError
        if self.array is None:
            raise ValueError("project needs the moments of a fitted matrix: call svd() first")
        src = self.array
        p_dim = tuple(x.shape)[1] if len(tuple(x.shape)) == 2 else None
        if p_dim is None:
            raise ValueError("expected a 2-D array of new samples, got shape {}".format(tuple(x.shape)))
        onto_t = as_tensor(onto).to(F64)
        if onto_t.ndim != 2 or p_dim not in tuple(onto_t.shape):
            raise ValueError("P in shape of `x` {} and shape of `onto` {} must match".format(tuple(x.shape),
                                                                                           tuple(onto_t.shape)))
        if onto_t.shape[0] != p_dim:
            onto_t = onto_t.T
        onto_t = onto_t.contiguous()
        mean = inv_std = impute = None
        if isinstance(src, GenotypeOperator):
            impute = src.impute_mean if src.impute_mean is not None else src.store.moments()[0]
            if scale_center_x:
                mean, inv_std = src.mean, src.inv_std
        elif isinstance(src, ScaledCenterArray) and scale_center_x:
            mean = src.center_vector.t if src.center else None
            inv_std = src.scale_vector.t if src.scale else None
        for v in (mean, inv_std):
            if v is not None and v.shape[0] != p_dim:
                raise ValueError("x has {} columns, the fitted moments {}".format(p_dim, v.shape[0]))
        if isinstance(x, GenotypeOperator):
            x = x.store
        store = None
        if isinstance(x, GenotypeStore):
            store = x
        elif getattr(src, "store", None) is not None or isinstance(src, ScaledCenterArray):
            try:
                store = GenotypeStore.from_dense(x)
            except NotGenotypeError:
                store = None
        if store is not None:
            op = GenotypeOperator(store, mean, inv_std, planes=getattr(src, "planes", 3))
            if op.has_missing:
                if impute is None:
                    impute = mean if mean is not None else store.moments()[0]
                op.impute_mean = impute.contiguous()
            return DeviceArray(op.dot_t(onto_t))
        xt = as_tensor(x).to(F64)
        if bool(torch.isnan(xt).any().item()):
            fill = impute if impute is not None else (mean if mean is not None else torch.nanmean(xt, dim=0))
            xt = torch.where(torch.isnan(xt), fill.expand_as(xt), xt)
        if mean is not None:
            xt = xt - mean
        if inv_std is not None:
            xt = xt * inv_std
        return DeviceArray(xt @ onto_t)

    # ------------------------------------------------------------------ per-iterate products
    def _state(self, x: torch.Tensor) -> dict:
        """q, r of the iterate (CholeskyQR2) and, lazily, t = A^T q and the next iterate A t / factor.

        With ``speculate`` (default) the two products of the coming step are queued on the device right behind the
        orthonormalisation, BEFORE the host asks for r: the K x K SVD, the q-vals check and the Python bookkeeping
        then run while the GPU works, instead of the GPU idling behind them.  The only cost is one unused step when the
        iterate turns out to be converged."""
        st = self._cache
        if st.get("x") is not x:
            group = _group_of(self.array)
            K = x.shape[1]
            if x.is_cuda and K <= 128:
                pend = orthonormalize_begin(x, group, stats_for=self.array)
                st = self._cache = {"x": x, "q": pend.q}
                if self.speculate:
                    self._next(st)
                r = pend.finish()
                if r is not None:
                    st["r"] = r
                    return st
            q, r = orthonormalize(x, group)      # rank-revealing host path (also raises LinAlgError on NaN/Inf)
            st = self._cache = {"x": x, "q": q, "r": r}
        return st

    def _t(self, st):
        if "t" not in st:
            st["t"] = _tdot(self.array, st["q"])
        return st["t"]

    def _next(self, st):
        if "next_raw" not in st:
            fused = None
            if "t" not in st and hasattr(self.array, "apass_t"):
                fused = self.array.apass_t(st["q"])          # one C call: t = A^T q and A t (lmdec_apass)
            if fused is not None:
                st["t"], st["next_raw"] = fused
            else:
                st["next_raw"] = self.array.dot_t(self._t(st))   # A A^T q, before the factor
        return st["next_raw"]

    def _wrap_input(self, data):
        if self.scale is None and self.center is None:
            return as_operator(data)
        from ..array.scaled_legacy import ScaledCenterArray
        if isinstance(data, ScaledCenterArray):
            return data
        sa = ScaledCenterArray(scale=bool(self.scale), center=bool(self.center), std_dist=self.std_dist)
        sa.fit(data)
        self.scaled_array = sa
        return sa

    def _noise(self, x: torch.Tensor, lmbd, seed=None, sharded: bool = True) -> torch.Tensor:
        """x <- (1-lmbd) x + lmbd ||x_col|| / sqrt(n) N(0,1)   (iter_methods.py:372-375, 577-580)."""
        gen = torch.Generator(device=x.device)
        # x is REPLICATED over the ranks that shard the other axis (the n x K iterate over SNP shards; the p x K
        # hand-over of the batched method over sample-row shards): every rank must then draw the same noise, or the
        # ranks leave the loop at different iterations and the next collective hangs.  An unseeded draw is agreed
        # on through one 8-byte all-reduce.
        replicated_over = _snp_group_of(self.array) if sharded else _group_of(self.array)
        if seed is None:
            if replicated_over is not None:
                import torch.distributed as dist
                s_ = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(x.device)
                dist.all_reduce(s_, op=dist.ReduceOp.MAX, group=replicated_over)
                gen.manual_seed(int(s_.item()))
            else:
                gen.seed()
        else:
            gen.manual_seed(int(seed))
        c2 = (x * x).sum(dim=0)
        g = _group_of(self.array) if sharded else _snp_group_of(self.array)
        n = self.array.shape[0] if sharded else self.array.shape[1]
        if g is not None:
            import torch.distributed as dist
            dist.all_reduce(c2, group=g)
        c_norms = torch.sqrt(c2)
        z = torch.randn(x.shape, generator=gen, dtype=F64, device=x.device)
        return x * (1 - lmbd) + (lmbd * c_norms / np.sqrt(n)) * z

    def _initialization(self, data, x0=None, seed=None, **kwargs):
        vec_t = self.k + self.buffer
        if len(tuple(data.shape)) != 2:
            raise ValueError("expected a 2-D array, got shape {}".format(tuple(data.shape)))
        if vec_t > min(data.shape):
            raise ValueError("Cannot find more than min(n,p) singular values of array function."
                             "Currently k = {}, buffer = {}. k + b > min(n,p)".format(self.k, self.buffer))
        self.array = self._wrap_input(data)
        self._cache = {}
        if self.factor == "n":
            self.factor = self.array.shape[0]
        elif self.factor == "p":
            self.factor = self.array.shape[1]
        elif self.factor is None:
            self.factor = False
        if x0 is not None:
            x = as_tensor(x0).clone()
            if x.shape != (getattr(self.array, "local_rows", self.array.shape[0]), vec_t):
                raise ValueError("x0 must have shape (n, k + buffer)")
            return x
        if self.sub_svd_start == "sym_mat_mult":
            return sym_mat_mult_start(self.array, vec_t).t
        if self.sub_svd_start:
            x = sub_svd_init(self.array, k=vec_t, warm_start_row_factor=self.init_row_sampling_factor, log=0).t
            if self.lmbd:
                x = self._noise(x, self.lmbd, seed)
            return x
        return rnormal_start(self.array, vec_t, log=0).t

    def _solution_step(self, x, **kwargs):
        st = self._state(x)
        nxt = self._next(st)
        if self.factor:
            nxt = nxt / self.factor
        return nxt

    def _solution_accuracy(self, x, **kwargs):
        st = self._state(x)
        q, r = st["q"], st["r"]
        u_r, s, vh = np.linalg.svd(r, full_matrices=False)   # tsqr(x, compute_svd=True): matrix_ops.py:72
        k = self.k
        dev = x.device
        u_rk_np = np.ascontiguousarray(u_r[:, :k])
        _dev_cache = {}

        def u_rk():           # uploaded at most once, and only if somebody needs it
            if "u" not in _dev_cache:
                _dev_cache["u"] = torch.as_tensor(u_rk_np, device=dev)
            return _dev_cache["u"]

        U_k = LazyDeviceArray(lambda: q @ u_rk(), (x.shape[0], u_rk_np.shape[1]))
        S_np = np.sqrt(s)[:k]
        S_k = LazyDeviceArray(lambda: torch.as_tensor(S_np, device=dev), S_np.shape)
        full_v = any(m in self.scoring_method for m in ["rmse", "v-subspace"])
        if full_v:
            # subspace_to_V (matrix_ops.py:89-99): left singular vectors of A^T x = t r; t is reused by the step
            w = self._t(st).to(F64) @ torch.as_tensor(r, device=dev)
            v, _, _ = tsqr_svd(w, _snp_group_of(self.array))
            V_k = DeviceArray(v.T[:k, :])
        else:
            vh_k = np.ascontiguousarray(vh[:k, :])
            V_k = LazyDeviceArray(lambda: torch.as_tensor(vh_k, device=dev), vh_k.shape)
        self.history.iter["last_value"] = (U_k, S_k, V_k)
        acc_list = []
        for method in self.scoring_method:
            if method == "q-vals":
                try:
                    prev_S_k = self.history.iter["S"][-1]
                    acc = q_value_converge(S_np, prev_S_k)
                except IndexError:
                    acc = float("INF")
                self.history.iter["S"].append(S_np)
            elif method == "rmse":
                # rmse_k(array, U_k, S_k**2, factor) (iter_methods.py:410): A A^T U_k = (A t) u_r, shared with the step
                n, p = self.array.shape
                aatu = self._next(st) @ u_rk()
                acc = _rmse_from_products(aatu, U_k.t, S_k.t ** 2, n, k, p, self.factor, _group_of(self.array))
            else:  # 'v-subspace'
                try:
                    prev_V_k = self.history.iter["V"][-1]
                    acc = subspace_dist(V_k.T, prev_V_k.T, S_np, group=_snp_group_of(self.array))
                except IndexError:
                    acc = float("INF")
                self.history.iter["V"].append(V_k)
            acc_list.append(acc)
        return acc_list

    def _finalization(self, x, return_history, **kwargs):
        if self.history.iter["last_value"][2].shape[1] != getattr(self.array, "local_cols", self.array.shape[1]):
            U, S, _ = self.history.iter["last_value"]
            st = self._cache if self._cache.get("x") is x else None
            if st is not None and "t" in st:
                w = st["t"].to(F64) @ torch.as_tensor(st["r"], device=x.device)
                v, _, _ = tsqr_svd(w, _snp_group_of(self.array))
                V = DeviceArray(v.T[:self.k, :])
            else:
                V = subspace_to_V(x, self.array, self.k)
            self.history.iter["last_value"] = U, S, V
        if return_history:
            return self.history
        return self.history.iter["last_value"]


class _vPowerMethod(PowerMethod):
    """PowerMethod seeded with a right-side start v (iter_methods.py:436-480)."""

    def __init__(self, v_start, k=None, buffer=None, max_iter=None, scoring_method=None, tol=None, full_svd=False,
                 factor=None, warn=False):
        _IterAlgo.__init__(self, max_iter=max_iter, scoring_method=scoring_method, tol=tol, factor=factor, warn=warn)
        self.lmbd = 1
        self.v_start = as_tensor(v_start)
        self.k, self.buffer, self.full_svd = k, buffer, full_svd
        self.scale = self.center = self.std_dist = None
        self.speculate = True
        self._cache = {}

    def _initialization(self, data, **kwargs):
        self.array = data
        self._cache = {}
        return self.array.dot_t(self.v_start)

    def _finalization(self, x, return_history, **kwargs):
        if not self.full_svd:
            st = self._cache if self._cache.get("x") is x else None
            if st is not None and "t" in st:
                return st["t"].to(F64) @ torch.as_tensor(st["r"], device=x.device)    # A^T x = (A^T q) r
            return _tdot(self.array, x).to(F64)
        return PowerMethod._finalization(self, x, return_history=False)


class SuccessiveBatchedPowerMethod(PowerMethod):
    """Power method over cumulative row batches, each seeded by the previous batch's V (iter_methods.py:483-600)."""

    def __init__(self, k: int = 10, max_sub_iter: int = 50, factor: Optional[str] = "n", scoring_method="q-vals",
                 f=.1, lmbd=.1, tol=.1, buffer: int = 10, sub_svd_start=True, init_row_sampling_factor: int = 5,
                 max_sub_time=1000, warn: bool = False):
        super().__init__(k=k, max_iter=max_sub_iter, scoring_method=scoring_method, tol=tol, factor=factor,
                         lmbd=lmbd, buffer=buffer, sub_svd_start=sub_svd_start,
                         init_row_sampling_factor=init_row_sampling_factor, time_limit=max_sub_time, warn=warn)
        self.f = f
        self.sub_svds = None

    def _reset_history(self):
        super()._reset_history()
        self.history.acc["sub_svd_acc"] = []
        self.history.iter["sub_svd"] = []

    @recover_last_value
    def svd(self, array, verbose: bool = False, return_history: bool = False, x0=None, seed=None, **kwargs):
        self._reset_history()
        self.history.time["start"] = time.time()
        self.array = self._wrap_input(array)
        if self.factor == "n":
            self.factor = self.array.shape[0]
        elif self.factor == "p":
            self.factor = self.array.shape[1]
        elif self.factor is None:
            self.factor = False
        vec_t = self.k + self.buffer
        partitions = array_constant_partition(self.array.shape, f=self.f, min_size=vec_t)
        partitions = cumulative_partition(partitions)
        sub_array = self.array[partitions[0], :]
        if x0 is not None:
            x = as_tensor(x0).clone()
        elif self.sub_svd_start == "warm":
            x = sub_svd_init(sub_array, k=vec_t, warm_start_row_factor=self.init_row_sampling_factor, log=0).t
        else:
            x = rnormal_start(sub_array, k=vec_t, log=0).t
        x = _tdot(sub_array, x).to(F64)
        for i, part in enumerate(tqdm(partitions[:-1], disable=not verbose)):
            _PM = _vPowerMethod(v_start=x, k=self.k, buffer=self.buffer, max_iter=self.max_iter, factor=self.factor,
                                scoring_method=self.scoring_method, tol=self.tol, warn=self.warn)
            x = _PM.svd(self.array[part, :], verbose=False)
            self.history.iter["last_value"] = _PM.history.iter["last_value"]
            self.history.iter["sub_svd"].append(self.history.iter["last_value"])
            if self.lmbd:
                x = self._noise(x, self.lmbd, None if seed is None else seed + i, sharded=False)  # x is [p, K] here
            if "v-subspace" in self.scoring_method:
                self.history.iter["V"].append(_PM.history.iter["V"][-1])
            self.history.iter["S"].append(_PM.history.iter["S"][-1])
            self.history.acc["sub_svd_acc"].append(_PM.history.acc)
            self.history.time["iter"].append(_PM.history.time)
        _PM = _vPowerMethod(v_start=x, k=self.k, buffer=self.buffer, max_iter=self.max_iter, factor=self.factor,
                            scoring_method=self.scoring_method, tol=self.tol, full_svd=True, warn=self.warn)
        _PM.svd(self.array, verbose=False)
        self.history.iter["last_value"] = _PM.history.iter["last_value"]
        self.history.iter["sub_svd"].append(self.history.iter["last_value"])
        self.history.time["stop"] = time.time()
        if return_history:
            return self.history
        return self.history.iter["last_value"]
