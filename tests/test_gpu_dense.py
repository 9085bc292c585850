"""Dense float32 operands on the tcgen05 path (csrc/dense_mma.cuh): store round trip (bit-exact), the 3xTF32 product
against a float64 reference of the same float32 matrix (float32-level accuracy), row prefixes, split contractions, and
the public API (ChainedArray of a dense member, BASELINE configs[4] in miniature) against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from lmdec_b200 import _capi
    return _capi, _capi.load()


@pytest.mark.parametrize("shape", [(128, 32), (300, 77), (1, 5), (1000, 1031)])
@pytest.mark.parametrize("transpose", [False, True])
def test_dense_store_round_trip_is_bit_exact(shape, transpose):
    capi, lib = _lib()
    dev = torch.device("cuda", 0)
    n, p = shape
    a = torch.randn((n, p), generator=torch.Generator(device=dev).manual_seed(1), device=dev, dtype=torch.float32)
    M, Kd = (p, n) if transpose else (n, p)
    st = torch.full((lib.lmdec_dense_store_bytes(M, Kd) // 4,), float("nan"), dtype=torch.float32, device=dev)
    capi.check(lib.lmdec_dense_pack(capi.ptr(a), 0, n, p, p, int(transpose), capi.ptr(st), capi.stream_ptr()))
    assert not bool(torch.isnan(st).any())          # every byte of every tile written (padding = 0)
    back = torch.empty((M, Kd), dtype=torch.float32, device=dev)
    capi.check(lib.lmdec_dense_unpack(capi.ptr(st), M, Kd, capi.ptr(back), Kd, capi.stream_ptr()))
    assert torch.equal(back, a.T if transpose else a)
    # float64 source: rounded to float32 once
    a64 = a.double() * (1 + 2.0 ** -30)
    capi.check(lib.lmdec_dense_pack(capi.ptr(a64), 1, n, p, p, int(transpose), capi.ptr(st), capi.stream_ptr()))
    capi.check(lib.lmdec_dense_unpack(capi.ptr(st), M, Kd, capi.ptr(back), Kd, capi.stream_ptr()))
    assert torch.equal(back, (a64.T if transpose else a64).float())


def _rel(out, ref):
    return float((out - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("n,p,K", [(1000, 777, 20), (300, 5000, 110), (129, 33, 1), (4096, 640, 128), (700, 300, 150)])
def test_dense_products_match_float64(n, p, K, monkeypatch):
    from lmdec_b200.array.operators import DenseOperator
    monkeypatch.setattr(DenseOperator, "native_min_elems", 0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(n + p + K)
    a = torch.randn((n, p), generator=g, device=dev, dtype=torch.float32) * 3 + 0.5
    op = DenseOperator(a, keep_dtype=True)
    assert op.path == "tcgen05"
    t = torch.randn((p, K), generator=g, device=dev, dtype=torch.float64)
    y = torch.randn((n, K), generator=g, device=dev, dtype=torch.float64)
    a64 = a.double()
    # 3xTF32 with float32 accumulation: a few float32 ulps of the largest entry
    assert _rel(op.dot_t(t), a64 @ t) < 4e-6
    assert _rel(op.tdot_t(y), a64.T @ y) < 4e-6
    assert _rel(op.T.dot_t(y), a64.T @ y) < 4e-6
    assert _rel(op.T.tdot_t(t), a64 @ t) < 4e-6
    assert _rel(op.dot_t(t.float()), a64 @ t.float().double()) < 4e-6       # float32 operand
    assert _rel(op.dot_t(t[:, 0]), a64 @ t[:, 0]) < 4e-6                     # vector operand


def test_dense_long_contraction_splits(monkeypatch):
    """A^T y of a tall matrix: few m-tiles, the contraction is split over the SMs and accumulated atomically"""
    from lmdec_b200.array.operators import DenseOperator
    monkeypatch.setattr(DenseOperator, "native_min_elems", 0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    n, p, K = 200_000, 256, 30
    a = torch.randn((n, p), generator=g, device=dev, dtype=torch.float32)
    op = DenseOperator(a, keep_dtype=True)
    y = torch.randn((n, K), generator=g, device=dev, dtype=torch.float64)
    ref = a.double().T @ y
    out = op.tdot_t(y)
    assert _rel(out, ref) < 4e-6
    t = torch.randn((p, K), generator=g, device=dev, dtype=torch.float64)
    assert _rel(op.dot_t(t), a.double() @ t) < 4e-6


def test_dense_row_prefix_views(monkeypatch):
    from lmdec_b200.array.operators import DenseOperator
    monkeypatch.setattr(DenseOperator, "native_min_elems", 0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(6)
    n, p, K = 1500, 900, 12
    a = torch.randn((n, p), generator=g, device=dev, dtype=torch.float32)
    op = DenseOperator(a, keep_dtype=True)
    for m in (1, 127, 128, 1000):
        sub = op[0:m, :]
        assert sub.shape == (m, p) and sub.path == "tcgen05" and sub._root is op
        t = torch.randn((p, K), generator=g, device=dev, dtype=torch.float64)
        y = torch.randn((m, K), generator=g, device=dev, dtype=torch.float64)
        assert _rel(sub.dot_t(t), a[:m].double() @ t) < 4e-6
        assert _rel(sub.T.dot_t(y), a[:m].double().T @ y) < 4e-6
    assert op[np.array([3, 5, 9]), :]._root is None          # fancy rows: a new plain matrix


def test_dense_chained_power_method_matches_oracle(monkeypatch):
    """BASELINE configs[4] in miniature: dense float ChainedArray, k oversampled, sym_mat_mult start"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import lmdec_oracle as oracle
    from lmdec_b200 import PowerMethod
    from lmdec_b200.array.chained import ChainedArray
    from lmdec_b200.array.metrics import subspace_dist
    from lmdec_b200.array.operators import DenseOperator
    monkeypatch.setattr(DenseOperator, "native_min_elems", 0)
    rng = np.random.default_rng(3)
    n, p, k, buffer = 6000, 400, 10, 10
    # planted spectrum sigma_i = 100 * 0.8^i over a noise floor
    u, _ = np.linalg.qr(rng.standard_normal((n, 40)))
    v, _ = np.linalg.qr(rng.standard_normal((p, 40)))
    a = ((u * (100 * 0.8 ** np.arange(40))) @ v.T + 0.01 * rng.standard_normal((n, p))).astype(np.float32)
    x0 = rng.standard_normal((n, k + buffer))
    kw = dict(k=k, buffer=buffer, max_iter=12, tol=1e-12, scoring_method="q-vals", sub_svd_start=False, lmbd=0,
              factor=None)
    arr = ChainedArray([a])
    pm = PowerMethod(**kw)
    U, S, V = pm.svd(arr, x0=x0, verbose=False)
    member = arr.arrays[0] if hasattr(arr, "arrays") else None
    if member is not None and hasattr(member, "path"):
        assert member.path == "tcgen05"
    Uo, So, Vo = oracle.PowerMethod(**kw).svd(a.astype(np.float64), x0=x0)
    S, So = np.asarray(S), np.asarray(So)
    assert np.max(np.abs(S - So) / So) < 1e-4
    assert subspace_dist(np.asarray(U), np.asarray(Uo), So) < 1e-4
    assert subspace_dist(np.asarray(V).T, np.asarray(Vo).T, So) < 1e-4
    # against LAPACK on the float32 matrix
    s_true = np.linalg.svd(a.astype(np.float64), compute_uv=False)[:k]
    assert np.max(np.abs(S - s_true) / s_true) < 1e-4
