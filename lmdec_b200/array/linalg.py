"""Orthonormalisation and small-matrix algebra.

The reference orthonormalises the iterate with dask's ``tsqr`` (lmdec/decomp/iter_methods.py:382; with
``compute_svd=True`` at lmdec/array/matrix_ops.py:72,94 and lmdec/decomp/init_methods.py:35,144).  Here the tall
factor never leaves the GPU: a Gram matrix per rank (``lmdec_gram``), one K x K all-reduce, and the K x K Cholesky /
SVD on the host in float64 (CholeskyQR2).  ``x = q r`` holds with ``q`` orthonormal; ``r`` is upper triangular unless
the rank-revealing fallback had to be taken.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..device import (F64, GRAM_LIBRARY_MIN_K, chol_whiten, cholqr2_grid, cholqr2_grid_fits, cholqr2_small, cholqr2_small_fits, gram,
                      gram_chol, qr_pack)


def _allreduce(t: torch.Tensor, group):
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, group=group)
    return t


_GRAM_RINGS = {}


def _gram_allreduce(x: torch.Tensor, group):
    """sum over the row shards of x^T x.  With NVSwitch multicast the per-rank Gram kernel adds its result straight into
    every rank's copy of a symmetric K x K buffer (`lmdec_gram_reduce`) and one signal-pad barrier follows -- no NCCL
    call on the latency-critical K x K reductions of the sharded CholeskyQR2; otherwise Gram kernel + all-reduce."""
    K = x.shape[1]
    if group is None or not x.is_cuda or K > GRAM_LIBRARY_MIN_K or x.dtype != F64 \
            or os.environ.get("LMDEC_FUSED_REDUCE", "1") == "0":
        return _allreduce(gram(x), group)
    key = (id(group), K, x.device.index)
    ring = _GRAM_RINGS.get(key, False)
    if ring is False:
        ring = None
        import torch.distributed as dist
        buf = hdl = None
        n_el = ((K * K + 31) // 32) * 32
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty((4 * n_el,), dtype=F64, device=x.device)
            hdl = symm_mem.rendezvous(buf, group)
            mine = 1 if hdl.multicast_ptr else 0
        except Exception:
            if os.environ.get("LMDEC_FUSED_REDUCE") == "require":
                raise
            mine = 0
        ok = torch.tensor([mine], device=x.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)          # every rank takes the same path
        if int(ok.item()):
            buf.zero_()
            hdl.barrier(channel=0)
            ring = {"buf": buf, "hdl": hdl, "n_el": n_el, "pos": 0}
        _GRAM_RINGS[key] = ring
    if ring is None:
        return _allreduce(gram(x), group)
    from .. import _capi
    from ..device import _rows, _work
    lib = _capi.load()
    buf, hdl, n_el, cur = ring["buf"], ring["hdl"], ring["n_el"], ring["pos"]
    nxt = (cur + 1) % 4
    xr = _rows(x)
    m = xr.shape[0]
    work = _work(x.device, lib.lmdec_gram_work_doubles(max(m, 1), K))
    _capi.check(lib.lmdec_gram_reduce(_capi.ptr(xr), xr.stride(0), m, K,
                                      _capi.C.c_void_p(hdl.multicast_ptr + cur * n_el * 8), _capi.ptr(work),
                                      _capi.stream_ptr()), "lmdec_gram_reduce")
    buf[nxt * n_el:(nxt + 1) * n_el].zero_()
    hdl.barrier(channel=0)
    ring["pos"] = nxt
    return buf[cur * n_el: cur * n_el + K * K].view(K, K)


def _whiten_factor(G: np.ndarray, rank_tol: float = 1e-12):
    """Return (r_inv, r, deficient) with r_inv^T G r_inv = I on the numerical range of G.

    G is first scaled to unit diagonal (column scales of the iterate can differ by many orders of magnitude), then
    factored by Cholesky; a pivot below `rank_tol` switches to an eigen-decomposition that drops the null directions
    (their columns of r_inv are zero and `deficient` is True)."""
    if not np.all(np.isfinite(G)):
        # what LAPACK reports through dask/numpy when NaN/Inf reach the QR: tests/test_iter_methods.py:318-325
        raise np.linalg.LinAlgError("SVD did not converge (non-finite values in the iterate)")
    K = G.shape[0]
    G = (G + G.T) / 2
    d = np.sqrt(np.clip(np.diag(G), 0, None))
    zero = d == 0
    d_safe = np.where(zero, 1.0, d)
    Gs = G / np.outer(d_safe, d_safe)
    try:
        if zero.any():
            raise np.linalg.LinAlgError("zero column")
        L = np.linalg.cholesky(Gs)
        if np.min(np.diag(L)) ** 2 < rank_tol:
            raise np.linalg.LinAlgError("numerically rank deficient")
        rs = L.T
        rs_inv = np.linalg.solve(rs, np.eye(K))
        deficient = False
    except np.linalg.LinAlgError:
        lam, V = np.linalg.eigh(Gs)
        keep = lam > rank_tol
        inv_sqrt = np.where(keep, 1.0 / np.sqrt(np.where(keep, lam, 1.0)), 0.0)
        rs_inv = V * inv_sqrt
        rs = (V * np.sqrt(np.clip(lam, 0, None))).T
        deficient = not bool(np.all(keep))
    return rs_inv / d_safe[:, None], rs * d_safe[None, :], deficient


def orthonormalize(x: torch.Tensor, group=None, seed: int = 0, gram_fn=None):
    """CholeskyQR2: returns (q [n,K] orthonormal, r [K,K] float64 on the host) with x = q r.

    Rank-deficient iterates (the reference's Householder QR completes them with arbitrary orthonormal columns) get
    their null directions refilled with seeded Gaussian columns before the second pass.
    """
    if x.ndim != 2:
        raise ValueError("expected a 2-D iterate")
    K = x.shape[1]
    if gram_fn is None and x.is_cuda and K <= 128:
        fast = _orthonormalize_device(x, group)
        if fast is not None:
            return fast
    r_total = np.eye(K)
    q = x
    gram_fn = gram_fn or gram      # the device kernel; tests of the host logic pass a CPU stand-in
    for it in range(4):
        G = _allreduce(gram_fn(q), group).cpu().numpy()
        if it >= 2 and np.max(np.abs(G - np.eye(K))) < 1e-9:
            break
        r_inv, r, deficient = _whiten_factor(G)
        q = q @ torch.as_tensor(r_inv, device=x.device)
        r_total = r @ r_total
        if deficient:
            dead = np.where(np.abs(r_inv).sum(axis=0) == 0)[0]
            gen = torch.Generator(device="cpu").manual_seed(seed + 17 * it)
            fill = torch.randn((x.shape[0], len(dead)), generator=gen, dtype=F64).to(x.device)
            q = q.clone()
            q[:, torch.as_tensor(dead, device=x.device)] = fill
            # the refilled columns carry no part of x: zero rows of r
            r_total[dead, :] = 0.0
    return q, r_total


_SIDE = {}
_SMALL_QR = os.environ.get("LMDEC_SMALL_QR", "1") != "0"   # single-launch CholeskyQR2 for small iterates (A/B switch)


def _side_stream(device):
    st = _SIDE.get(device)
    if st is None:
        st = _SIDE[device] = (torch.cuda.Stream(device=device), torch.empty(128 * 128 + 8, dtype=F64).pin_memory())
    return st


class _PendingQR:
    """Device half of CholeskyQR2 already enqueued; ``finish()`` makes the single device->host copy (r + health).

    The copy runs on a side stream that waits only for the orthonormalisation (an event recorded right behind it), so
    work queued on the main stream afterwards -- the products of the coming step -- does not delay it."""

    def __init__(self, q, packed, K):
        self.q, self._packed, self._K = q, packed, K
        self._ev = torch.cuda.Event()
        self._ev.record()

    def finish(self):
        """r [K,K] on the host, or None when the iterate was not numerically full rank / finite."""
        K = self._K
        side, pinned = _side_stream(self._packed.device)
        side.wait_event(self._ev)
        n = self._packed.numel()
        with torch.cuda.stream(side):
            pinned[:n].copy_(self._packed, non_blocking=True)
        self._packed.record_stream(side)
        side.synchronize()
        host = pinned[:n].numpy().copy()
        if host[K * K] != 0 or host[K * K + 1] != 0 or not (host[K * K + 2] < 1e-2):
            return None
        return host[:K * K].reshape(K, K)


def orthonormalize_begin(x: torch.Tensor, group=None, stats_for=None) -> _PendingQR:
    """Enqueue CholeskyQR2 entirely on the device (both K x K factorizations included) and return without syncing, so
    the caller can queue the next products behind it before asking for r.  `stats_for`: the operator whose A^T product
    consumes q next; when it offers `qr_stats_slot`, the single-launch kernel leaves q's column statistics there."""
    K = x.shape[1]
    if group is None and x.dtype == F64 and _SMALL_QR and K <= 32:
        small = cholqr2_small_fits(x.shape[0], K)
        if small or cholqr2_grid_fits(x.shape[0], K):
            slot = getattr(stats_for, "qr_stats_slot", None)
            work, lim = slot(x.shape[0], K) if slot is not None else (None, 0.0)
            # one launch for the whole CholeskyQR2: one 8-CTA cluster, or all SMs cooperatively for mid-size iterates
            q, packed = (cholqr2_small if small else cholqr2_grid)(x, stats_work=work, stats_lim=lim)
            if work is not None:
                stats_for.qr_stats_emitted(q)
            return _PendingQR(q, packed, K)
    if group is None and K <= GRAM_LIBRARY_MIN_K:
        _, ri1, r1, f1 = gram_chol(x)                  # 2 launches per pass, nothing else
        q1 = x @ ri1
        G2, ri2, r2, f2 = gram_chol(q1)
    else:
        G1 = _gram_allreduce(x, group)
        ri1, r1, f1 = chol_whiten(G1)
        q1 = x @ ri1
        G2 = _gram_allreduce(q1, group)
        ri2, r2, f2 = chol_whiten(G2)
    q = q1 @ ri2
    packed = qr_pack(r1, r2, G2, f1, f2)
    return _PendingQR(q, packed, K)


def _orthonormalize_device(x: torch.Tensor, group=None):
    pend = orthonormalize_begin(x, group)
    r = pend.finish()
    return None if r is None else (pend.q, r)


def tsqr_svd(x: torch.Tensor, group=None):
    """``tsqr(x, compute_svd=True)``: (u [n,K] on the device, s [K] host, vh [K,K] host)."""
    q, r = orthonormalize(x, group)
    u_r, s, vh = np.linalg.svd(r, full_matrices=False)
    u = q @ torch.as_tensor(u_r, device=x.device)
    return u, s, vh
