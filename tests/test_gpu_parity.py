"""GPU parity of the drop-in API against the CPU oracle (NumPy restatement of the reference) on the same inputs.

Tolerances (BASELINE.json north_star): singular values within 1e-4 relative; subspace distance of U and V <= 1e-4
under lmdec's own metric; genotype decode / counts bit-exact (tests/test_gpu_kernels.py).
"""
import numpy as np
import pytest

from helpers import oracle, structured_genotypes, synth_genotypes

pytestmark = pytest.mark.gpu

S_RTOL = 1e-4
SUBSPACE_TOL = 1e-4


def _np(a):
    return np.asarray(a)


@pytest.mark.parametrize("miss", [0.0, 0.05])
@pytest.mark.parametrize("mean,std", [(True, True), (True, False), (False, True), (False, False)])
def test_genotype_operator_products_match_oracle(cuda, miss, mean, std):
    from lmdec_b200.decomp import make_snp_array
    n, p, K = 300, 700, 12
    g = synth_genotypes(n, p, miss, seed=3)
    ref = oracle.make_snp_array(g, mean=mean, std=std, std_method="binom")
    op = make_snp_array(g, mean=mean, std=std, std_method="binom")
    assert op.shape == (n, p)
    rng = np.random.default_rng(0)
    t, y = rng.standard_normal((p, K)), rng.standard_normal((n, K))
    x_ref, w_ref = np.asarray(ref.dot(t)), np.asarray(ref.T.dot(y))
    x, w = _np(op.dot(t)), _np(op.T.dot(y))
    # fixed point with 3 planes: 22 bits below each operand column's maximum
    assert np.max(np.abs(x - x_ref)) <= 3e-6 * np.max(np.abs(x_ref))
    assert np.max(np.abs(w - w_ref)) <= 3e-6 * np.max(np.abs(w_ref))
    # 1-D operands
    assert np.max(np.abs(_np(op.dot(t[:, 0])) - x_ref[:, 0])) <= 3e-6 * np.max(np.abs(x_ref))
    # dense materialisation equals the oracle's
    np.testing.assert_allclose(_np(op), ref.toarray().astype(np.float64), rtol=1e-12, atol=1e-12)
    # 4 planes: 30 bits
    op4 = make_snp_array(g, mean=mean, std=std, std_method="binom", planes=4)
    assert np.max(np.abs(_np(op4.dot(t)) - x_ref)) <= 1e-7 * np.max(np.abs(x_ref))


def test_moments_bit_exact_counts_and_means(cuda):
    from lmdec_b200.decomp import get_array_moments
    from lmdec_b200.store import GenotypeStore
    g = synth_genotypes(513, 300, 0.07, seed=9)
    st = GenotypeStore.from_dense(g)
    cnt = st.code_counts().cpu().numpy()
    exp = np.stack([(g == 0).sum(0), (g == 1).sum(0), (g == 2).sum(0), np.isnan(g).sum(0)], axis=1)
    assert np.array_equal(cnt, exp)
    dec = st.to_dense().cpu().numpy().astype(np.float64)
    dec[dec < 0] = np.nan
    assert np.array_equal(np.isnan(dec), np.isnan(g)) and np.array_equal(np.nan_to_num(dec), np.nan_to_num(g))
    for method in ("binom", "norm"):
        m_ref, s_ref = oracle.get_array_moments(g, std_method=method)
        m, s = get_array_moments(st, std_method=method)
        assert np.array_equal(_np(m), m_ref)               # mean = exact integer ratio
        np.testing.assert_allclose(_np(s), s_ref, rtol=1e-13)


def _run_pair(g, k, buffer, tol, miss_kw=None, scoring="q-vals", max_iter=50, planes=3):
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    n = g.shape[0]
    x0 = np.random.default_rng(42).standard_normal((n, k + buffer))
    ref_op = oracle.make_snp_array(g, std_method="binom")
    ref = oracle.PowerMethod(k=k, buffer=buffer, tol=tol, scoring_method=scoring, sub_svd_start=False, lmbd=0,
                             max_iter=max_iter)
    Ur, Sr, Vr = ref.svd(ref_op, x0=x0)
    op = make_snp_array(g, std_method="binom", planes=planes)
    pm = PowerMethod(k=k, buffer=buffer, tol=tol, scoring_method=scoring, sub_svd_start=False, lmbd=0,
                     max_iter=max_iter)
    U, S, V = pm.svd(op, x0=x0, verbose=False)
    return (ref, Ur, Sr, Vr), (pm, _np(U), _np(S), _np(V))


@pytest.mark.parametrize("miss", [0.0, 0.05])
def test_powermethod_genotypes_matches_oracle(cuda, miss):
    g = structured_genotypes(600, 3000, miss=miss, seed=1)
    (ref, Ur, Sr, Vr), (pm, U, S, V) = _run_pair(g, k=10, buffer=10, tol=1e-6)
    assert pm.num_iter == ref.num_iter
    assert U.shape == Ur.shape and V.shape == Vr.shape and S.shape == Sr.shape
    np.testing.assert_allclose(S, Sr, rtol=S_RTOL)
    assert oracle.subspace_dist(U, Ur, Sr) <= SUBSPACE_TOL
    assert oracle.subspace_dist(V.T, Vr.T, Sr) <= SUBSPACE_TOL
    np.testing.assert_allclose(pm.history.acc["q-vals"][1:], ref.history.acc["q-vals"][1:], rtol=5e-2, atol=1e-7)
    # history contract (reference tests/test_iter_methods.py:477-497)
    assert len(pm.history.iter["S"]) == pm.num_iter and pm.history.iter["U"] == []
    converged = pm.history.acc["q-vals"][-1] <= 1e-6
    assert len(pm.history.time["iter"]) == (pm.num_iter - 1 if converged else pm.num_iter)


def test_powermethod_three_metrics_agree_with_oracle(cuda):
    g = structured_genotypes(400, 1500, miss=0.03, seed=5)
    scoring = ["q-vals", "rmse", "v-subspace"]
    (ref, Ur, Sr, Vr), (pm, U, S, V) = _run_pair(g, k=5, buffer=10, tol=[1e-9, 1e-9, 1e-9], scoring=scoring,
                                                 max_iter=6)
    assert pm.num_iter == ref.num_iter == 6
    np.testing.assert_allclose(S, Sr, rtol=S_RTOL)
    for m in scoring:
        a, b = np.array(pm.history.acc[m][1:]), np.array(ref.history.acc[m][1:])
        np.testing.assert_allclose(a, b, rtol=2e-2, atol=2e-7)
    assert oracle.subspace_dist(V.T, Vr.T, Sr) <= SUBSPACE_TOL


@pytest.mark.parametrize("k", [1, 3, 9])
def test_powermethod_dense_exact_recovery(cuda, k):
    """reference tests/test_iter_methods.py:12-33: PowerMethod recovers the exact SVD of a dense matrix."""
    from lmdec_b200 import PowerMethod
    rng = np.random.default_rng(k)
    n, p = 100, 60
    u, _ = np.linalg.qr(rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((p, p)))
    s = np.zeros(p)
    s[:30] = 1.5 ** -np.arange(30) * 10
    a = (u[:, :p] * s) @ v.T
    pm = PowerMethod(k=k, tol=1e-12, scoring_method="q-vals", max_iter=200, factor=None, buffer=10, lmbd=0)
    U, S, V = pm.svd(a, verbose=False)
    np.testing.assert_array_almost_equal(_np(S), s[:k], decimal=6)
    assert oracle.subspace_dist(_np(U), u[:, :k], s[:k], epsilon=0) <= 1e-8
    assert oracle.subspace_dist(_np(V).T, v[:, :k], s[:k], epsilon=0) <= 1e-8


def test_powermethod_validation_and_nan(cuda):
    from lmdec_b200 import PowerMethod
    for kw in (dict(max_iter=0), dict(max_iter=1.5), dict(time_limit=-1), dict(k=0), dict(buffer=-1),
               dict(init_row_sampling_factor=0), dict(tol=[1e-3, 1e-3]), dict(scoring_method="nope")):
        with pytest.raises(ValueError):
            PowerMethod(**kw)
    with pytest.raises(ValueError):
        PowerMethod(k=10, buffer=10).svd(np.random.rand(15, 15), verbose=False)
    with pytest.raises(ValueError):                      # project is implemented here (tests/test_gpu_round2.py)
        PowerMethod().project(np.zeros((2, 2)), np.zeros((2, 2)))
    a = np.random.default_rng(0).random((50, 40))
    a[3, 4] = np.nan
    with pytest.raises(np.linalg.LinAlgError):          # reference tests/test_iter_methods.py:318-325
        PowerMethod(k=2, buffer=2, sub_svd_start=False).svd(a, verbose=False)


def test_sbpm_matches_oracle(cuda):
    from lmdec_b200 import SuccessiveBatchedPowerMethod
    from lmdec_b200.decomp import make_snp_array
    g = structured_genotypes(1000, 800, miss=0.02, seed=7)
    k, buffer = 5, 5
    kw = dict(k=k, buffer=buffer, f=.2, lmbd=0, tol=1e-7, scoring_method="q-vals", max_sub_iter=40)
    ref_op = oracle.make_snp_array(g, std_method="binom")
    x0 = np.random.default_rng(1).standard_normal((200, k + buffer))
    ref = oracle.SuccessiveBatchedPowerMethod(**kw)
    Ur, Sr, Vr = ref.svd(ref_op, x0=x0)
    sb = SuccessiveBatchedPowerMethod(**kw)
    U, S, V = sb.svd(make_snp_array(g, std_method="binom"), x0=x0)
    np.testing.assert_allclose(_np(S), Sr, rtol=S_RTOL)
    assert oracle.subspace_dist(_np(U), Ur, Sr) <= SUBSPACE_TOL
    assert oracle.subspace_dist(_np(V).T, Vr.T, Sr) <= SUBSPACE_TOL
    assert len(sb.history.iter["sub_svd"]) == len(ref.history.iter["sub_svd"]) == 5
    assert len(sb.history.acc["sub_svd_acc"]) == 4


@pytest.mark.parametrize("scale,center,factor", [(True, True, None), (True, False, "n"), (False, True, "p"),
                                                 (False, False, None)])
@pytest.mark.parametrize("kind", ["geno", "float"])
def test_scaled_center_array_matches_oracle(cuda, scale, center, factor, kind):
    from lmdec_b200.array import ScaledCenterArray
    rng = np.random.default_rng(4)
    n, p, K = 40, 30, 3
    a = rng.integers(0, 3, size=(n, p)).astype(np.float64) if kind == "geno" else rng.standard_normal((n, p))
    ref = oracle.ScaledCenterArray(scale=scale, center=center, factor=factor)
    ref.fit(a)
    sa = ScaledCenterArray(scale=scale, center=center, factor=factor)
    sa.fit(a)
    x, y = rng.standard_normal((n, K)), rng.standard_normal((p, K))
    tol = dict(rtol=1e-4, atol=1e-4) if kind == "geno" else dict(rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(_np(sa.sym_mat_mult(x)), ref.sym_mat_mult(x), **tol)
    np.testing.assert_allclose(_np(sa.dot(y)), ref.dot(y), **tol)
    np.testing.assert_allclose(_np(sa.T.dot(x)), ref.T.dot(x), **tol)
    np.testing.assert_allclose(_np(sa.dot(y[:, :1])), ref.dot(y[:, :1]), **tol)        # (p,1) squeeze path
    np.testing.assert_allclose(_np(sa.array), ref.array, rtol=1e-12, atol=1e-12)
    assert sa.T.T is sa and sa.shape == (n, p) and sa.T.shape == (p, n)
    sub, sub_ref = sa[0:10, :], ref[0:10, :]                                          # refit on the selection
    np.testing.assert_allclose(_np(sub.center_vector), sub_ref.center_vector, rtol=1e-12, atol=1e-12)
    with pytest.raises(NotImplementedError):
        sa.T.sym_mat_mult(y)
    with pytest.raises(ValueError):
        sa.T.fit(a)


def test_stacked_chained_inputs(cuda):
    """the make_snp_array pattern spelled with the reference's own containers (decomp/utils.py:280-287)."""
    import scipy.sparse as sp
    from lmdec_b200 import PowerMethod
    from lmdec_b200.array import ChainedArray, StackedArray
    g = structured_genotypes(200, 300, miss=0.05, seed=11)
    mean, std = oracle.get_array_moments(g)
    filled, mask = oracle.mask_imputation(g, mean)
    arr = ChainedArray((StackedArray((StackedArray((filled.astype(np.int8), sp.coo_matrix(mask))), -mean)), 1 / std))
    ref = oracle.make_snp_array(g)
    rng = np.random.default_rng(2)
    t, y = rng.standard_normal((300, 4)), rng.standard_normal((200, 4))
    np.testing.assert_allclose(_np(arr.dot(t)), ref.dot(t), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(_np(arr.T.dot(y)), ref.T.dot(y), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(_np(arr[0:50, :].dot(t)), ref[0:50, :].dot(t), rtol=1e-10, atol=1e-10)
    x0 = rng.standard_normal((200, 8))
    Ur, Sr, Vr = oracle.PowerMethod(k=4, buffer=4, tol=1e-8, sub_svd_start=False, lmbd=0).svd(ref, x0=x0)
    # the drivers recognise the pattern and run it on the 2-bit tile store (tests/test_gpu_round2.py): fixed-point
    # products, north-star tolerances; the container's own .dot above stays the generic float64 algebra
    U, S, V = PowerMethod(k=4, buffer=4, tol=1e-8, sub_svd_start=False, lmbd=0).svd(arr, x0=x0, verbose=False)
    np.testing.assert_allclose(_np(S), Sr, rtol=S_RTOL)
    assert oracle.subspace_dist(_np(V).T, Vr.T, Sr) <= SUBSPACE_TOL


def test_metrics_closed_forms(cuda):
    """reference tests/test_metrics.py:13-162."""
    from lmdec_b200.array import q_value_converge, rmse_k, subspace_dist
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((50, 5)))
    s = np.arange(5, 0, -1.0)
    assert subspace_dist(q, q, s) <= 1e-6
    assert subspace_dist(q, q[:, ::-1].copy(), s[::-1].copy()) <= 1e-6
    with pytest.raises(ValueError):
        subspace_dist(2 * q, q, s)
    for a in (0.1, 0.5, 0.9):
        v1 = np.eye(4)[:, :2]
        v2 = np.stack([np.eye(4)[:, 0], np.sqrt(1 - a ** 2) * np.eye(4)[:, 1] + a * np.eye(4)[:, 2]], axis=1)
        got = subspace_dist(v1, v2, np.ones(2))
        assert abs(got - oracle.subspace_dist(v1, v2, np.ones(2))) < 1e-12
    si, sj = rng.uniform(1, 2, 7), rng.uniform(1, 2, 7)
    assert abs(q_value_converge(si, sj) - oracle.q_value_converge(si, sj)) < 1e-14
    assert abs(q_value_converge(si, sj, scale=False, norm=1) - oracle.q_value_converge(si, sj, False, 1)) < 1e-13
    with pytest.raises(ValueError):
        q_value_converge(si, sj[:3])
    N = 20
    eye = np.eye(N)
    for sv in (0.0, 0.5, 1.0):
        got = rmse_k(eye, eye, np.full(N, sv))
        assert abs(got - oracle.rmse_k(eye, eye, np.full(N, sv))) < 1e-14


def test_config0_powermethod_10k_x_50k_matches_oracle(cuda):
    """BASELINE.json configs[0]: PowerMethod k=10 on synthetic 10k x 50k binomial genotypes, centre + scale.
    The oracle (reference's CPU path restated) runs on the host cores with the same x0 and tolerance."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    n, p, k, buffer = 10_000, 50_000, 10, 10
    g = synth_genotypes(n, p, 0.0, seed=0, dtype=np.float32)
    x0 = np.random.default_rng(42).standard_normal((n, k + buffer))
    kw = dict(k=k, buffer=buffer, tol=1e-3, scoring_method="q-vals", sub_svd_start=False, lmbd=0, max_iter=12)
    ref = oracle.PowerMethod(**kw)
    Ur, Sr, Vr = ref.svd(oracle.make_snp_array(g, std_method="binom"), x0=x0)
    op = make_snp_array(g, std_method="binom")
    assert not op.has_missing
    pm = PowerMethod(**kw)
    U, S, V = pm.svd(op, x0=x0, verbose=False)
    assert pm.num_iter == ref.num_iter
    np.testing.assert_allclose(_np(S), Sr, rtol=S_RTOL)
    assert oracle.subspace_dist(_np(U), Ur, Sr) <= SUBSPACE_TOL
    assert oracle.subspace_dist(_np(V).T, Vr.T, Sr) <= SUBSPACE_TOL


def test_full_size_properties_config1(cuda):
    """configs[1] size (2,504 x 1,000,000, 5% missing): size-independent properties instead of an oracle run.
    (a) decode round trip and count checksum, (b) adjointness <A t, y> = <t, A^T y>, (c) linearity,
    (d) column means of the operator are zero (centring) on a column sample."""
    import torch
    from lmdec_b200.store import GenotypeOperator, GenotypeStore
    n, p = 2504, 1_000_000
    gen = torch.Generator(device=cuda).manual_seed(7)
    maf = torch.rand(p, generator=gen, device=cuda) * 0.45 + 0.05
    store = GenotypeStore(n, p, device=cuda)
    checksum = torch.zeros(4, dtype=torch.int64, device=cuda)
    keep = {}
    for r0 in range(0, n, 256):
        nb = min(256, n - r0)
        g = (torch.rand((nb, p), generator=gen, device=cuda) < maf).to(torch.int8)
        g += (torch.rand((nb, p), generator=gen, device=cuda) < maf).to(torch.int8)
        g[torch.rand((nb, p), generator=gen, device=cuda) < 0.05] = -1
        for c, v in enumerate((0, 1, 2, -1)):
            checksum[c] += (g == v).sum()
        if r0 in (0, 2304):
            keep[r0] = g.clone()
        store.pack_rows(g, r0)
    store.finish()
    cnt = store.code_counts()
    assert torch.equal(cnt.sum(0), checksum)                      # checksum of checksums, bit exact
    assert int(cnt.sum().item()) == n * p
    dec = store.to_dense(rows=256)
    assert torch.equal(dec, keep[0])                              # decode round trip, first panel
    dec_t = store.to_dense(transpose=True, rows=128)              # SNP-major store, first 128 SNPs, all samples
    assert torch.equal(dec_t[:, 2304:2504], keep[2304][:, :128].T.contiguous())
    mean, std = store.moments("binom")
    op = GenotypeOperator(store, mean, 1.0 / std)
    K = 20
    t = torch.randn((p, K), generator=gen, device=cuda, dtype=torch.float64)
    y = torch.randn((n, K), generator=gen, device=cuda, dtype=torch.float64)
    at, aty = op.dot_t(t), op.tdot_t(y)
    lhs, rhs = (at * y).sum().item(), (t * aty).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), torch.linalg.norm(at).item() * torch.linalg.norm(y).item())
    t2 = torch.randn((p, K), generator=gen, device=cuda, dtype=torch.float64)
    lin = op.dot_t(t + 2.0 * t2) - (at + 2.0 * op.dot_t(t2))
    assert torch.linalg.norm(lin).item() <= 1e-5 * torch.linalg.norm(at).item()
    ones = torch.ones((n, 1), dtype=torch.float64, device=cuda)
    col_sums = op.tdot_t(ones)[:, 0]                              # 1^T A = 0 for a centred, mean-imputed operator
    assert col_sums.abs().max().item() <= 1e-4 * np.sqrt(n)


def test_store_save_load_and_bed_roundtrip(cuda, tmp_path):
    """persistence of the packed operator (reference array/io.py save_array / load_array_from_disk) and the .bed loader."""
    import torch
    from lmdec_b200.store import GenotypeOperator, GenotypeStore
    g = synth_genotypes(300, 517, 0.06, seed=12)
    st = GenotypeStore.from_dense(g)
    st.save(str(tmp_path / "store"))
    st2 = GenotypeStore.load(str(tmp_path / "store"))
    assert torch.equal(st.s_store, st2.s_store) and torch.equal(st.v_store, st2.v_store)
    assert torch.equal(st.code_counts(), st2.code_counts())
    # .bed payload: variant-major PLINK codes (00 -> 0, 01 -> missing, 10 -> 1, 11 -> 2)
    dosage_to_plink = {0: 0, 1: 2, 2: 3}
    codes = np.vectorize(lambda v: 1 if np.isnan(v) else dosage_to_plink[int(v)])(g.T).astype(np.uint8)   # [p, n]
    bpv = (300 + 3) // 4
    padded = np.zeros((517, bpv * 4), dtype=np.uint8)
    padded[:, :300] = codes
    bed = (padded[:, 0::4] | (padded[:, 1::4] << 2) | (padded[:, 2::4] << 4) | (padded[:, 3::4] << 6)).astype(np.uint8)
    st3 = GenotypeStore.from_bed_bytes(bed.tobytes(), n_samples=300, n_variants=517)
    assert torch.equal(st3.s_store, st.s_store) and torch.equal(st3.v_store, st.v_store)


def test_dense_float32_tf32x3_matches_float64(cuda):
    """float32 dense inputs (configs[4] path): the error-compensated TF32 split keeps fp32-level accuracy."""
    from lmdec_b200.array import ChainedArray
    from lmdec_b200.array.operators import DenseOperator
    rng = np.random.default_rng(3)
    a = rng.standard_normal((3000, 500)).astype(np.float32)
    x, y = rng.standard_normal((500, 12)), rng.standard_normal((3000, 12))
    op = DenseOperator(a, keep_dtype=True)
    assert op.a.dtype.is_floating_point and op.a.element_size() == 4
    op.f32_mode = "tf32x3"
    import torch
    ref_ax, ref_aty = a.astype(np.float64) @ x, a.astype(np.float64).T @ y
    ax, aty = op.dot_t(torch.as_tensor(x, device=cuda)).cpu().numpy(), op.tdot_t(torch.as_tensor(y, device=cuda)).cpu().numpy()
    assert np.max(np.abs(ax - ref_ax)) <= 2e-5 * np.max(np.abs(ref_ax))
    assert np.max(np.abs(aty - ref_aty)) <= 2e-5 * np.max(np.abs(ref_aty))
    from lmdec_b200 import PowerMethod
    u, s, v = np.linalg.svd(a.astype(np.float64), full_matrices=False)
    U, S, V = PowerMethod(k=3, buffer=10, tol=1e-10, factor=None, lmbd=0, max_iter=300).svd(ChainedArray([a]), verbose=False)
    np.testing.assert_allclose(_np(S), s[:3], rtol=1e-5)


@pytest.mark.parametrize("n,p,K,miss", [(300, 2000, 20, 0.05), (130, 700, 5, 0.0), (257, 1500, 32, 0.1)])
def test_epilogue_column_statistics_match_standalone_pass(cuda, n, p, K, miss):
    """A^T y leaves the statistics of its result for the following A t (lmdec_operator_apply flags 16 / 32): same
    fixed-point multiplier bit for bit, same rank-1 shift to rounding, and the same product."""
    import torch
    from lmdec_b200.decomp import make_snp_array
    rng = np.random.default_rng(n + p)
    g = rng.binomial(2, rng.uniform(0.05, 0.5, size=p), size=(n, p)).astype(float)
    if miss:
        g[rng.random((n, p)) < miss] = np.nan
    op = make_snp_array(g, std_method="binom")
    y = torch.from_numpy(rng.standard_normal((n, K))).to(cuda)
    t = op.tdot_t(y)
    buf = op._buffers(K)
    assert "emitted" in buf
    x_fused = op.dot_t(t)
    stats_fused = buf["work"][:5 * K].cpu().numpy().copy()
    assert "emitted" not in buf
    x_plain = op.dot_t(t.clone())                      # another tensor object: the stand-alone statistics pass runs
    stats_plain = buf["work"][:5 * K].cpu().numpy().copy()
    # mult, 1/mult and the maximum that sets the scale: the float32 epilogue takes the maximum on the fp32 pipe (two
    # roundings of 2^-24 against the float64 pass); the fixed-point range keeps a factor ~2 of headroom above `lim`
    np.testing.assert_allclose(stats_fused[:2 * K], stats_plain[:2 * K], rtol=5e-7)
    np.testing.assert_allclose(np.maximum(stats_fused[3 * K:4 * K], stats_fused[4 * K:]),
                               np.maximum(stats_plain[3 * K:4 * K], stats_plain[4 * K:]), rtol=5e-7)
    bound = 2.0 * stats_plain[3 * K:].max() * p               # sum_r |w_r t_rk| <= p max|mu| max|d t|
    np.testing.assert_allclose(stats_fused[2 * K:3 * K], stats_plain[2 * K:3 * K], rtol=0, atol=1e-14 * bound)
    np.testing.assert_allclose(x_fused.cpu().numpy(), x_plain.cpu().numpy(), rtol=0,
                               atol=2e-6 * np.abs(x_plain.cpu().numpy()).max())   # two fixed-point scales 1e-7 apart
    t2 = op.tdot_t(y)
    t2 += 1.0                                           # modified in place: the emitted statistics no longer apply
    x_mod = op.dot_t(t2)
    x_ref = op.dot_t(t2.clone())
    assert np.array_equal(x_mod.cpu().numpy(), x_ref.cpu().numpy())


@pytest.mark.parametrize("n,p,K,miss", [(300, 2000, 20, 0.05), (130, 700, 5, 0.0)])
def test_qr_kernel_leaves_statistics_for_tdot(cuda, n, p, K, miss):
    """lmdec_cholqr2_small leaves q's column statistics for the A^T q that follows (flag 32 with transpose)."""
    import torch
    from lmdec_b200.array.linalg import orthonormalize_begin
    from lmdec_b200.decomp import make_snp_array
    rng = np.random.default_rng(n * 7 + p)
    g = rng.binomial(2, rng.uniform(0.05, 0.5, size=p), size=(n, p)).astype(float)
    if miss:
        g[rng.random((n, p)) < miss] = np.nan
    op = make_snp_array(g, std_method="binom")
    x = torch.from_numpy(rng.standard_normal((n, K))).to(cuda)
    pend = orthonormalize_begin(x, None, stats_for=op)
    q = pend.q
    assert pend.finish() is not None
    buf = op._buffers(K)
    assert buf["emitted"][0] == "q"
    t_fused = op.tdot_t(q)
    stats_fused = buf["work"][:5 * K].cpu().numpy().copy()
    t_plain = op.tdot_t(q.clone())
    stats_plain = buf["work"][:5 * K].cpu().numpy().copy()
    assert np.array_equal(stats_fused[:2 * K], stats_plain[:2 * K])
    assert np.array_equal(stats_fused[3 * K:4 * K], stats_plain[3 * K:4 * K])
    np.testing.assert_allclose(stats_fused[2 * K:3 * K], stats_plain[2 * K:3 * K], rtol=0, atol=1e-14 * n)
    np.testing.assert_allclose(t_fused.cpu().numpy(), t_plain.cpu().numpy(), rtol=0,
                               atol=1e-9 * np.abs(t_plain.cpu().numpy()).max())
