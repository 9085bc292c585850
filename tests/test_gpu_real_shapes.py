"""Oracle parity at the REAL shapes of BASELINE configs[2], [3] and [4] (configs[1] is in test_gpu_round2.py).

The oracle cannot hold these matrices, but a column slab of A^T y and a product A t whose operand vanishes outside the
slab are computable on the host exactly from the decoded codes; the dense configuration is compared with float64 products
of the same float32 matrix.  Each test frees its stores: the three together need < 60 GB of the 180 GB of a B200."""
import gc

import numpy as np
import pytest

from helpers import oracle

pytestmark = pytest.mark.gpu


def _slab_parity(cuda, n, p, K, slab, seed, miss=0.0):
    import torch
    from bench import make_shard
    from lmdec_b200.store import GenotypeOperator
    store = make_shard(torch, cuda, n, p, miss, seed=seed)
    mean, std = store.moments("binom")
    op = GenotypeOperator(store, mean, 1.0 / std)
    codes = store.to_dense(transpose=True, rows=slab).cpu().numpy()            # [slab, n] int8, -1 = missing
    g = codes.T.astype(np.float64)
    g[g < 0] = np.nan
    ref = oracle.make_snp_array(g, std_method="binom")
    rng = np.random.default_rng(seed)
    y = np.linalg.qr(rng.standard_normal((n, K)))[0]
    t_slab = rng.standard_normal((slab, K))
    w_ref = np.asarray(ref.T.dot(y))
    x_ref = np.asarray(ref.dot(t_slab))
    w = op.tdot_t(torch.as_tensor(y, device=cuda))[:slab].double().cpu().numpy()
    t_full = torch.zeros((p, K), dtype=torch.float64, device=cuda)
    t_full[:slab] = torch.as_tensor(t_slab, device=cuda)
    x = op.dot_t(t_full).cpu().numpy()
    assert np.max(np.abs(w - w_ref)) <= 3e-6 * np.max(np.abs(w_ref))
    assert np.max(np.abs(x - x_ref)) <= 3e-6 * np.max(np.abs(x_ref))
    return store, op, g, ref


def _free(*objs):
    import torch
    del objs
    gc.collect()
    torch.cuda.empty_cache()


def test_config3_ukb_shard_products_match_oracle_on_a_slab(cuda):
    """configs[3]: the per-GPU shard of the UKB-shaped 488k x 800k matrix at 8 GPUs, 61,000 x 800,000, K = 30: every
    SNP tile / contraction chunk of both products runs; rows of A^T y on a 256-SNP slab and A t for an operand supported
    on it against oracle.make_snp_array on the decoded codes."""
    out = _slab_parity(cuda, 61_000, 800_000, 30, 256, seed=33)
    _free(*out)


def test_config2_sbpm_input_products_and_prefixes_match_oracle_on_a_slab(cuda):
    """configs[2]: 50,000 x 2,000,000, K = 30, the SuccessiveBatchedPowerMethod input.  Full-matrix slab parity, then
    the cumulative row prefixes the batched method multiplies by ([0:m, :] views of the same stores, moments NOT refit:
    lmdec/array/chained.py:235-241) against the oracle's prefix of the same operator."""
    import torch
    n, p, K, slab = 50_000, 2_000_000, 30, 256
    store, op, g, ref = _slab_parity(cuda, n, p, K, slab, seed=22)
    rng = np.random.default_rng(7)
    for m in (5_000, 25_000):               # first and fifth cumulative batch of f = .1
        sub, sub_ref = op[0:m, :], ref[0:m, :]
        assert sub.shape == (m, p)
        y = rng.standard_normal((m, K))
        w_ref = np.asarray(sub_ref.T.dot(y))
        w = sub.tdot_t(torch.as_tensor(y, device=cuda))[:slab].double().cpu().numpy()
        assert np.max(np.abs(w - w_ref)) <= 3e-6 * np.max(np.abs(w_ref))
        t_slab = rng.standard_normal((slab, K))
        t_full = torch.zeros((p, K), dtype=torch.float64, device=cuda)
        t_full[:slab] = torch.as_tensor(t_slab, device=cuda)
        x = sub.dot_t(t_full).cpu().numpy()
        x_ref = np.asarray(sub_ref.dot(t_slab))
        assert x.shape == (m, K) and np.max(np.abs(x - x_ref)) <= 3e-6 * np.max(np.abs(x_ref))
    _free(store, op, g, ref)


def test_config4_dense_full_size_products_match_float64(cuda):
    """configs[4]: dense float32 1,000,000 x 4,000, K = 110 on the tcgen05 3xTF32 kernel against float64 products of the
    same float32 matrix (computed on the device in row blocks -- what the oracle's numpy matmul computes)."""
    import torch
    from lmdec_b200.array.operators import DenseOperator
    n, p, K = 1_000_000, 4000, 110
    g = torch.Generator(device=cuda).manual_seed(3)
    a = torch.empty((n, p), dtype=torch.float32, device=cuda)
    for r0 in range(0, n, 65536):
        a[r0:r0 + 65536].normal_(generator=g)
    a += 0.25                                                  # a non-centred matrix: same-sign partial sums
    t = torch.randn((p, K), generator=g, dtype=torch.float64, device=cuda)
    y = torch.randn((n, K), generator=g, dtype=torch.float64, device=cuda)
    op = DenseOperator(a, keep_dtype=True)
    assert op.path == "tcgen05"
    x, w = op.dot_t(t), op.tdot_t(y)
    w_ref = torch.zeros((p, K), dtype=torch.float64, device=cuda)
    x_err = x_max = 0.0
    for r0 in range(0, n, 32768):
        blk = a[r0:r0 + 32768].double()
        x_ref = blk @ t
        x_err = max(x_err, float((x[r0:r0 + 32768] - x_ref).abs().max()))
        x_max = max(x_max, float(x_ref.abs().max()))
        w_ref += blk.T @ y[r0:r0 + 32768]
    assert x_err <= 4e-6 * x_max
    assert float((w - w_ref).abs().max()) <= 4e-6 * float(w_ref.abs().max())
    _free(a, op, x, w)
