"""Start matrices (reference lmdec/decomp/init_methods.py:13-189)."""
from __future__ import annotations

import numpy as np
import torch

from ..array.linalg import tsqr_svd
from ..array.operators import as_operator
from ..array.random import array_split
from ..device import F64, DeviceArray, as_tensor


def _group(a):
    return getattr(getattr(a, "store", None), "group", None) or getattr(a, "group", None)


def _local_slice(array, n_global):
    """rows of a global [n, K] host matrix owned by this rank (block-contiguous shards)."""
    store = getattr(array, "store", None)
    if store is not None and store.group is not None:
        return slice(store.row_offset, store.row_offset + array.local_rows)
    return slice(0, n_global)


DEVICE_RNG_MIN_ELEMS = 1 << 24


def rnormal_start(array, k: int, seed: int = 42, log: int = 0):
    """N(0,1) start [n, k].  The reference draws from dask's chunked RandomState(seed) (init_methods.py:180-184), whose
    stream depends on the dask chunking; here it is NumPy's RandomState(seed) drawn on the host (global matrix, so every
    sharding sees the same start) and moved to the device.  Unsharded starts of DEVICE_RNG_MIN_ELEMS values or more
    (configs[4]: 1M x 110, 1.6 s of host RNG -- three quarters of the whole time-to-converged) are drawn by the device
    generator with the same seed instead: a different, equally seeded stream (the RNG stream is unpinned either way)."""
    array = as_operator(array)
    n = array.shape[0]
    store = getattr(array, "store", None)
    sharded = store is not None and (store.group is not None or store.snp_group is not None)
    if n * k >= DEVICE_RNG_MIN_ELEMS and not sharded:
        import torch
        from ..device import F64, require_cuda
        dev = require_cuda()
        gen = torch.Generator(device=dev).manual_seed(int(seed))
        out = DeviceArray(torch.randn((n, k), generator=gen, dtype=F64, device=dev))
        return (out, {}) if log else out
    omega = np.random.RandomState(seed).standard_normal(size=(n, k))
    out = DeviceArray(as_tensor(omega[_local_slice(array, n)]))
    return (out, {}) if log else out


def v_init(array, v, log: int = 0):
    """x0 = left singular vectors of A v (init_methods.py:33-35)."""
    array = as_operator(array)
    x = array.dot_t(as_tensor(v))
    u, _, _ = tsqr_svd(x, _group(array))
    out = DeviceArray(u)
    return (out, {}) if log else out


def _sub_svd(array, sub_array_t: torch.Tensor, k: int, log: int = 0):
    """SVD of the [p, rows] sample, then v_init with its top-k left vectors (init_methods.py:144-147)."""
    v, _, _ = tsqr_svd(sub_array_t, getattr(getattr(array, "store", None), "snp_group", None))
    out = v_init(array, v[:, :k])
    return (out, {}) if log else out


def sub_svd_init(array, k: int, warm_start_row_factor: float = 5, seed: int = 42, log: int = 0):
    """Warm start from a seeded sample of warm_start_row_factor*k rows (init_methods.py:90-115)."""
    array = as_operator(array)
    rows = int(warm_start_row_factor * k)
    n, p = array.shape
    idx, _ = array_split(array_shape=array.shape, f=rows / n, axis=0, seed=seed)
    try:
        sub = array[idx, :]
    except NotImplementedError:
        sub = array[0:rows, :]
    sub_t = sub.to_dense().T.contiguous()
    if _group(array) is not None:
        import torch.distributed as dist
        # the sample is a prefix held by the leading ranks: gather it by summing zero-padded copies
        full = torch.zeros((p, rows), dtype=F64, device=sub_t.device)
        off = array.store.row_offset
        if off < rows:
            full[:, off:off + sub_t.shape[1]] = sub_t[:, :max(0, rows - off)]
        dist.all_reduce(full, group=_group(array))
        sub_t = full
    out = _sub_svd(array, sub_t, k)
    return (out, {}) if log else out


def sym_mat_mult_start(array, k: int, seed: int = 42, log: int = 0):
    """Randomised range-finder start  x0 = A A^T omega  for a Gaussian omega [n, k].

    NOT in the reference checkout (no such symbol); named by the project brief.  Defined here as one application of the
    Gram operator (``ScaledCenterArray.sym_mat_mult``, scaled_legacy.py:292-399) to ``rnormal_start``.  Parity unpinned.
    """
    array = as_operator(array)
    omega = rnormal_start(array, k, seed=seed).t
    t = array.tdot_t(omega) if hasattr(array, "tdot_t") else array.T.dot_t(omega)
    out = DeviceArray(array.dot_t(t).to(F64))
    return (out, {}) if log else out
