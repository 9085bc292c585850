"""GPU parity of the C-ABI kernels against exact integer / float64 NumPy arithmetic (bit-exact where integer)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _capi():
    from lmdec_b200 import _capi
    return _capi, _capi.load()


def _geno(rng, n, p, miss=0.0):
    maf = rng.uniform(0.05, 0.5, size=p)
    g = rng.binomial(2, maf, size=(n, p)).astype(np.int8)
    if miss:
        g[rng.random((n, p)) < miss] = -1
    return g


def _pack(torch, capi, lib, g_dev, M, Kd, stride_m, stride_kd):
    store = torch.empty(lib.lmdec_store_bytes(M, Kd), dtype=torch.uint8, device=g_dev.device)
    bad = torch.zeros(1, dtype=torch.int32, device=g_dev.device)
    capi.check(lib.lmdec_pack_2bit(capi.ptr(g_dev), 0, stride_m, stride_kd, M, Kd, 0, 0, capi.ptr(store), M, Kd,
                                   capi.ptr(bad), capi.stream_ptr()), "pack")
    assert int(bad.item()) == 0
    return store


@pytest.mark.parametrize("n,p,miss", [(5, 7, 0.0), (130, 300, 0.1), (257, 1025, 0.05), (1000, 513, 0.0)])
def test_pack_unpack_counts_transpose_bit_exact(cuda, n, p, miss):
    import torch
    capi, lib = _capi()
    rng = np.random.default_rng(n * 1000 + p)
    g = _geno(rng, n, p, miss)
    gd = torch.from_numpy(g).to(cuda)
    s_store = _pack(torch, capi, lib, gd, n, p, p, 1)      # sample-major
    v_store = _pack(torch, capi, lib, gd, p, n, 1, p)      # SNP-major
    out = torch.full((n, p), 9, dtype=torch.int8, device=cuda)
    capi.check(lib.lmdec_unpack_2bit(capi.ptr(s_store), n, p, capi.ptr(out), p, capi.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), g)
    out_t = torch.full((p, n), 9, dtype=torch.int8, device=cuda)
    capi.check(lib.lmdec_unpack_2bit(capi.ptr(v_store), p, n, capi.ptr(out_t), n, capi.stream_ptr()))
    assert np.array_equal(out_t.cpu().numpy(), g.T)
    # transpose of the sample-major store == SNP-major store, byte for byte
    v2 = torch.empty_like(v_store)
    capi.check(lib.lmdec_store_transpose(capi.ptr(s_store), n, p, capi.ptr(v2), capi.stream_ptr()))
    assert torch.equal(v2, v_store)
    # counts (full and a ragged prefix)
    for lim in (n, max(1, n - 3), max(1, n // 2)):
        cnt = torch.empty((p, 4), dtype=torch.int64, device=cuda)
        capi.check(lib.lmdec_code_counts(capi.ptr(v_store), p, n, lim, capi.ptr(cnt), capi.stream_ptr()))
        sub = g[:lim]
        exp = np.stack([(sub == 0).sum(0), (sub == 1).sum(0), (sub == 2).sum(0), (sub == -1).sum(0)], axis=1)
        assert np.array_equal(cnt.cpu().numpy(), exp)


def _quantize(torch, capi, lib, B, kd_limit, K, P, N, mult, row1=None, row2=None):
    Kd = B.shape[0]
    nbytes = lib.lmdec_image_bytes(Kd, N)
    im1 = torch.empty(nbytes, dtype=torch.int8, device=B.device)
    im2 = torch.empty(nbytes, dtype=torch.int8, device=B.device) if row2 is not None else None
    capi.check(lib.lmdec_quantize_planes(capi.ptr(B), B.stride(0), Kd, kd_limit, K, P, N, capi.ptr(mult),
                                         capi.ptr(row1), capi.ptr(row2), capi.ptr(im1), capi.ptr(im2),
                                         capi.stream_ptr()), "quantize")
    return im1, im2


@pytest.mark.parametrize("M,Kd,K,P,miss,mode,splits", [
    (128, 256, 20, 3, 0.0, 0, 1),
    (100, 300, 20, 3, 0.0, 0, 1),
    (300, 1000, 20, 4, 0.0, 0, 0),
    (300, 5000, 20, 3, 0.05, 0, 0),
    (513, 2049, 30, 3, 0.05, 1, 1),
    (2000, 700, 10, 2, 0.1, 1, 0),
    (260, 40000, 20, 3, 0.05, 0, 7),
    (1, 1, 1, 1, 0.0, 0, 1),
    (40000, 260, 30, 4, 0.02, 1, 0),
    # long contractions: two m-tiles per staged operand chunk (pair modes M / E / T), odd tile counts, split and unsplit
    (300, 20000, 20, 3, 0.05, 1, 0),
    (300, 20000, 20, 3, 0.05, 1, 1),
    (700, 17000, 20, 3, 0.0, 0, 0),
    (640, 17000, 30, 3, 0.0, 0, 1),
    (129, 33000, 20, 4, 0.0, 0, 3),
    (256, 16384, 10, 2, 0.03, 0, 2),
])
def test_geno_matmul_exact(cuda, M, Kd, K, P, miss, mode, splits):
    import torch
    capi, lib = _capi()
    rng = np.random.default_rng(M + 7 * Kd + K)
    g = _geno(rng, M, Kd, miss)
    gd = torch.from_numpy(g).to(cuda)
    store = _pack(torch, capi, lib, gd, M, Kd, Kd, 1)
    N = ((K * P + 15) // 16) * 16
    B = rng.standard_normal((Kd, K))
    r1 = rng.uniform(0.5, 2.0, size=Kd)
    r2 = rng.uniform(0.0, 2.0, size=Kd)
    Bd, r1d, r2d = (torch.from_numpy(a).to(cuda) for a in (B, r1, r2))
    amax = torch.empty(K, dtype=torch.float64, device=cuda)
    amax2 = torch.empty(K, dtype=torch.float64, device=cuda)
    capi.check(lib.lmdec_colabsmax(capi.ptr(Bd), K, Kd, K, capi.ptr(r1d), capi.ptr(r2d), capi.ptr(amax),
                                   capi.ptr(amax2), capi.stream_ptr()))
    np.testing.assert_array_equal(amax.cpu().numpy(), np.abs(B * r1[:, None]).max(0))
    np.testing.assert_array_equal(amax2.cpu().numpy(), np.abs(B * r1[:, None] * r2[:, None]).max(0))
    scale = np.maximum(amax.cpu().numpy(), amax2.cpu().numpy())
    has_missing = int(miss > 0)
    two_images = has_missing and mode == 0
    mult = (2.0 ** (8 * P - (3 if two_images else 2))) / scale
    multd = torch.from_numpy(mult).to(cuda)
    im1, im2 = _quantize(torch, capi, lib, Bd, Kd, K, P, N, multd, r1d, r2d if two_images else None)
    out0 = torch.full((M, K), -7, dtype=torch.int64, device=cuda)
    out1 = torch.full((M, K), -7, dtype=torch.int64, device=cuda)
    capi.check(lib.lmdec_geno_matmul(None, capi.ptr(store), M, Kd, M, Kd, has_missing, mode, capi.ptr(im1), capi.ptr(im2),
                                     K, P, N, capi.ptr(out0), capi.ptr(out1), splits, capi.stream_ptr()), "matmul")
    torch.cuda.synchronize()
    q1 = np.rint((B * r1[:, None]) * mult).astype(np.int64)
    q2 = np.rint((B * r1[:, None]) * mult * r2[:, None]).astype(np.int64)
    val = np.where(g < 0, 0, g).astype(np.int64)
    ind = (g < 0).astype(np.int64)
    if mode == 0:
        exp0 = val @ q1 + (ind @ q2 if has_missing else 0)
        assert np.array_equal(out0.cpu().numpy(), exp0)
    else:
        assert np.array_equal(out0.cpu().numpy(), val @ q1)
        assert np.array_equal(out1.cpu().numpy(), ind @ q1)


def test_geno_matmul_prefix_limits(cuda):
    """row-prefix (m_limit) and contraction-prefix (kd_limit) arguments used by SuccessiveBatchedPowerMethod."""
    import torch
    capi, lib = _capi()
    rng = np.random.default_rng(5)
    M, Kd, K, P = 700, 1500, 20, 3
    N = 64
    g = _geno(rng, M, Kd, 0.05)
    store = _pack(torch, capi, lib, torch.from_numpy(g).to(cuda), M, Kd, Kd, 1)
    B = rng.standard_normal((Kd, K))
    Bd = torch.from_numpy(B).to(cuda)
    for m_lim, kd_lim in [(700, 1500), (301, 1500), (700, 777), (129, 33)]:
        mult = np.full(K, (2.0 ** (8 * P - 2)) / np.abs(B[:kd_lim]).max())
        multd = torch.from_numpy(mult).to(cuda)
        im1, _ = _quantize(torch, capi, lib, Bd, kd_lim, K, P, N, multd)
        out0 = torch.full((M, K), -7, dtype=torch.int64, device=cuda)
        out1 = torch.full((M, K), -7, dtype=torch.int64, device=cuda)
        capi.check(lib.lmdec_geno_matmul(None, capi.ptr(store), M, Kd, m_lim, kd_lim, 1, 1, capi.ptr(im1), None, K, P, N,
                                         capi.ptr(out0), capi.ptr(out1), 0, capi.stream_ptr()))
        q1 = np.rint(B[:kd_lim] * mult).astype(np.int64)
        sub = g[:m_lim, :kd_lim]
        assert np.array_equal(out0.cpu().numpy()[:m_lim], np.where(sub < 0, 0, sub).astype(np.int64) @ q1)
        assert np.array_equal(out1.cpu().numpy()[:m_lim], (sub < 0).astype(np.int64) @ q1)
        assert (out0.cpu().numpy()[m_lim:] == -7).all()


def test_finalize_gram_qvals_sqdiff(cuda):
    import torch
    capi, lib = _capi()
    rng = np.random.default_rng(11)
    M, K = 1000, 20
    a0 = rng.integers(-2 ** 40, 2 ** 40, size=(M, K))
    a1 = rng.integers(-2 ** 30, 2 ** 30, size=(M, K))
    cs, ra, rs, rsh, csh = (rng.standard_normal(s) for s in (K, M, M, M, K))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    out = torch.empty((M, K), dtype=torch.float64, device=cuda)
    d = [t(a) for a in (a0, a1, cs, ra, rs, rsh, csh)]  # keep the device copies alive across the async call
    capi.check(lib.lmdec_finalize(capi.ptr(d[0]), capi.ptr(d[1]), M, K, capi.ptr(d[2]), capi.ptr(d[3]),
                                  capi.ptr(d[4]), capi.ptr(d[5]), capi.ptr(d[6]), capi.ptr(out), K,
                                  capi.stream_ptr()))
    exp = (a0 + ra[:, None] * a1) * cs * rs[:, None] + rsh[:, None] * csh
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=1e-14, atol=0)
    for m in (1, 127, 5000, 200000):
        X, Y = rng.standard_normal((m, K)), rng.standard_normal((m, K))
        Xd, Yd = t(X), t(Y)
        work = torch.empty(lib.lmdec_gram_work_doubles(m, K), dtype=torch.float64, device=cuda)
        G = torch.empty((K, K), dtype=torch.float64, device=cuda)
        cs_ = torch.empty(K, dtype=torch.float64, device=cuda)
        capi.check(lib.lmdec_gram(capi.ptr(Xd), K, None, 0, m, K, capi.ptr(G), capi.ptr(cs_), capi.ptr(work),
                                  capi.stream_ptr()))
        np.testing.assert_allclose(G.cpu().numpy(), X.T @ X, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(cs_.cpu().numpy(), X.sum(0), rtol=1e-12, atol=1e-9)
        capi.check(lib.lmdec_gram(capi.ptr(Xd), K, capi.ptr(Yd), K, m, K, capi.ptr(G), None, capi.ptr(work),
                                  capi.stream_ptr()))
        np.testing.assert_allclose(G.cpu().numpy(), X.T @ Y, rtol=1e-12, atol=1e-9)
        s = rng.uniform(0.5, 2, K)
        sd = t(s)
        o = torch.empty(1, dtype=torch.float64, device=cuda)
        w2 = torch.empty(600, dtype=torch.float64, device=cuda)
        capi.check(lib.lmdec_sqdiff(capi.ptr(Xd), K, capi.ptr(Yd), K, capi.ptr(sd), m, K, capi.ptr(o), capi.ptr(w2),
                                    capi.stream_ptr()))
        np.testing.assert_allclose(o.item(), ((X - Y * s) ** 2).sum(), rtol=1e-12)
    si, sj = rng.uniform(1, 2, 10), rng.uniform(1, 2, 10)
    o = torch.empty(1, dtype=torch.float64, device=cuda)
    sid, sjd = t(si), t(sj)
    capi.check(lib.lmdec_qvals(capi.ptr(sid), capi.ptr(sjd), 10, 1, capi.ptr(o), capi.stream_ptr()))
    exp = np.linalg.norm(si - sj) / ((np.linalg.norm(si) + np.linalg.norm(sj)) / 2)
    np.testing.assert_allclose(o.item(), exp, rtol=1e-14)


def test_recode_bed_matches_plink_code_map(cuda):
    import torch
    capi, lib = _capi()
    rng = np.random.default_rng(3)
    n_var, n_samp = 300, 1003
    codes = rng.integers(0, 4, size=(n_var, n_samp)).astype(np.uint8)       # PLINK 2-bit codes
    bpv = (n_samp + 3) // 4
    padded = np.zeros((n_var, bpv * 4), dtype=np.uint8)
    padded[:, :n_samp] = codes
    padded[:, n_samp:] = 1  # garbage in the byte padding must be ignored
    bed = (padded[:, 0::4] | (padded[:, 1::4] << 2) | (padded[:, 2::4] << 4) | (padded[:, 3::4] << 6)).astype(np.uint8)
    store = torch.empty(lib.lmdec_store_bytes(n_var, n_samp), dtype=torch.uint8, device=cuda)
    bed_d = torch.from_numpy(bed).to(cuda)
    capi.check(lib.lmdec_recode_bed(capi.ptr(bed_d), n_var, n_samp, capi.ptr(store), capi.stream_ptr()))
    out = torch.empty((n_var, n_samp), dtype=torch.int8, device=cuda)
    capi.check(lib.lmdec_unpack_2bit(capi.ptr(store), n_var, n_samp, capi.ptr(out), n_samp, capi.stream_ptr()))
    plink_to_dosage = np.array([0, -1, 1, 2], dtype=np.int8)   # 00 hom-A1, 01 missing, 10 het, 11 hom-A2
    assert np.array_equal(out.cpu().numpy(), plink_to_dosage[codes])
    cnt = torch.empty((n_var, 4), dtype=torch.int64, device=cuda)
    capi.check(lib.lmdec_code_counts(capi.ptr(store), n_var, n_samp, n_samp, capi.ptr(cnt), capi.stream_ptr()))
    assert int(cnt.sum().item()) == n_var * n_samp


@pytest.mark.parametrize("n,K", [(2504, 20), (3, 2), (37, 1), (256, 32), (2048, 32), (5000, 13)])
def test_cholqr2_small_single_launch(cuda, n, K):
    """One-cluster CholeskyQR2 == the general multi-launch path: Q^T Q = I, Q R = X, R upper with positive diagonal."""
    import torch
    from lmdec_b200.device import cholqr2_small, cholqr2_small_fits
    assert cholqr2_small_fits(n, K)
    rng = np.random.default_rng(n * 31 + K)
    x = rng.standard_normal((n, K)) @ (np.eye(K) + 0.3 * rng.standard_normal((K, K))) * np.logspace(0, 3, K)
    if n < K:
        pytest.skip("wide iterate")
    xd = torch.from_numpy(np.pad(x, ((0, 0), (0, 3)))).to(cuda)[:, :K]      # strided rows (ldx = K + 3)
    q, packed = cholqr2_small(xd)
    q, packed = q.cpu().numpy(), packed.cpu().numpy()
    r = packed[:K * K].reshape(K, K)
    assert packed[K * K] == 0 and packed[K * K + 1] == 0 and packed[K * K + 2] < 1e-6
    np.testing.assert_allclose(q.T @ q, np.eye(K), atol=1e-13)
    np.testing.assert_allclose(q @ r, x, rtol=0, atol=1e-12 * np.abs(x).max())
    assert np.allclose(r, np.triu(r)) and (np.diag(r) > 0).all()
    qn, rn = np.linalg.qr(x)
    sgn = np.sign(np.diag(rn))
    np.testing.assert_allclose(r, rn * sgn[:, None], rtol=1e-9, atol=1e-10 * np.abs(rn).max())


def test_cholqr2_small_flags_rank_deficiency(cuda):
    import torch
    from lmdec_b200.device import cholqr2_small
    from lmdec_b200.array.linalg import orthonormalize
    rng = np.random.default_rng(5)
    x = rng.standard_normal((500, 6))
    x[:, 4] = x[:, 1]
    xd = torch.from_numpy(x).to(cuda)
    _, packed = cholqr2_small(xd)
    packed = packed.cpu().numpy()
    assert packed[36] != 0 or packed[37] != 0 or not packed[38] < 1e-2
    q, r = orthonormalize(xd)                       # the host fallback takes over
    q = q.cpu().numpy()
    np.testing.assert_allclose(q.T @ q, np.eye(6), atol=1e-10)
    np.testing.assert_allclose(q @ r, x, atol=1e-10)
    xd[7, 2] = float("nan")
    _, packed = cholqr2_small(xd)
    packed = packed.cpu().numpy()
    assert packed[36] != 0 or packed[37] != 0 or not packed[38] < 1e-2


@pytest.mark.parametrize("n,K", [(10000, 20), (12500, 20), (61000, 30), (50, 3), (2504, 20), (70001, 32)])
def test_cholqr2_grid_cooperative_launch(cuda, n, K):
    """Mid-size iterates: the all-SM cooperative CholeskyQR2 gives Q^T Q = I, Q R = X, the R of Householder QR, and the
    column statistics of Q for the product that follows."""
    import torch
    from lmdec_b200.device import cholqr2_grid, cholqr2_grid_fits
    capi, lib = _capi()
    assert cholqr2_grid_fits(n, K)
    rng = np.random.default_rng(n * 13 + K)
    x = rng.standard_normal((n, K)) @ (np.eye(K) + 0.3 * rng.standard_normal((K, K))) * np.logspace(0, 3, K)
    xd = torch.from_numpy(np.pad(x, ((0, 0), (0, 5)))).to(cuda)[:, :K]      # strided rows
    work = torch.zeros(lib.lmdec_apply_work_doubles(K), dtype=torch.float64, device=cuda)
    lim = float(2 ** 22)
    q, packed = cholqr2_grid(xd, stats_work=work, stats_lim=lim)
    q, packed, stats = q.cpu().numpy(), packed.cpu().numpy(), work[:5 * K].cpu().numpy()
    r = packed[:K * K].reshape(K, K)
    assert packed[K * K] == 0 and packed[K * K + 1] == 0 and packed[K * K + 2] < 1e-6
    np.testing.assert_allclose(q.T @ q, np.eye(K), atol=1e-13)
    np.testing.assert_allclose(q @ r, x, rtol=0, atol=1e-12 * np.abs(x).max())
    rn = np.linalg.qr(x)[1]
    np.testing.assert_allclose(r, rn * np.sign(np.diag(rn))[:, None], rtol=1e-8, atol=1e-10 * np.abs(rn).max())
    amax = np.abs(q).max(axis=0)
    assert np.array_equal(stats[3 * K:4 * K], amax) and np.array_equal(stats[:K], lim / amax)
    np.testing.assert_allclose(stats[2 * K:3 * K], -q.sum(axis=0), rtol=0, atol=1e-12)
    # non-finite input is flagged, not silently factorised
    xd2 = xd.clone()
    xd2[n // 2, K // 2] = float("inf")
    _, packed2 = cholqr2_grid(xd2)
    packed2 = packed2.cpu().numpy()
    assert packed2[K * K] != 0 or packed2[K * K + 1] != 0 or not packed2[K * K + 2] < 1e-2
