"""Pins the CPU oracle (oracle/lmdec_oracle.py) against the reference's own golden vectors and known-answer tests.

Every case restates a test of the reference (file:line given) with the oracle in place of lmdec; the stored golden
SVD comes from lmdec/examples/ex1_{Uk,Sk,Vk}.npy via tests/golden/make_golden.py.  CPU only.
"""
import os

import numpy as np
import pytest

from helpers import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ex1_golden.npz")


def _ex1_matrix():
    n, p = 20000, 8000                       # lmdec/examples/array.py:50-54
    np.random.seed(10)
    return np.random.randn(n, p) - 0.05 * np.random.rand(n, p)


def test_golden_ex1_known_answer_svd():
    """lmdec/examples/array.py:13-62 stores the exact top-10 SVD; the oracle's PowerMethod must reproduce it."""
    gold = np.load(GOLDEN)
    np.testing.assert_allclose(gold["uu"], np.eye(10), atol=1e-10)
    np.testing.assert_allclose(gold["vv"], np.eye(10), atol=1e-10)
    a = _ex1_matrix()
    pm = oracle.PowerMethod(k=10, tol=1e-9, scoring_method="q-vals", max_iter=14, factor=None, lmbd=0,
                            sub_svd_start=False)
    U, S, V = pm.svd(a)
    Sk = gold["Sk"]
    assert abs(S[0] - Sk[0]) / Sk[0] < 1e-8                      # isolated leading value: converged
    np.testing.assert_allclose(S[1:], Sk[1:], rtol=3e-2)         # tight cluster 229.4..230.7: Ritz values, 14 its
    assert 1 - abs(U[:, 0] @ gold["u1"]) < 1e-8
    assert 1 - abs(V[0, :] @ gold["v1"]) < 1e-8
    # golden vectors are singular vectors of the regenerated matrix (checks the fixture against the generator)
    np.testing.assert_allclose(a.T @ gold["u1"], Sk[0] * gold["v1"], atol=1e-6)


def test_powermethod_case1_exact_recovery():
    """tests/test_iter_methods.py:12-33 (float matrix through make_snp_array, std_method='norm', rmse criterion)."""
    rng = np.random.RandomState(0)
    n, p = 100, 80
    array = rng.rand(n, p)
    sc = (array - array.mean(axis=0)).dot(np.diag(1 / array.std(axis=0)))
    U, S, V = np.linalg.svd(sc, full_matrices=False)
    op = oracle.make_snp_array(array, mean=True, std=True, std_method="norm", mask_nan=False, dtype="float64")
    for k in range(1, 10):
        pm = oracle.PowerMethod(k=k, tol=1e-9, scoring_method="rmse", max_iter=100, sub_svd_start=False,
                                init_row_sampling_factor=1, factor=None, lmbd=0)
        U_pm, S_pm, V_pm = pm.svd(op)
        np.testing.assert_array_almost_equal(S[:k], S_pm)
        assert V_pm.shape == (k, p) and U_pm.shape == (n, k)
        np.testing.assert_almost_equal(oracle.subspace_dist(V[:k], V_pm, S_pm), 0)
        np.testing.assert_almost_equal(oracle.subspace_dist(U[:, :k], U_pm, S_pm), 0)


def test_powermethod_geometric_spectrum():
    """tests/test_iter_methods.py:69-86: diag(1.01^i), default sub-svd start with noise."""
    N, k = 1000, 10
    s_orig = np.array([1.01 ** i for i in range(1, N + 1)])
    pm = oracle.PowerMethod(tol=1e-16, factor=None, lmbd=.1, max_iter=100)
    U, S, V = pm.svd(np.diag(s_orig), rng=np.random.default_rng(0))
    np.testing.assert_array_almost_equal(s_orig[-k:][::-1], S, decimal=0)
    v = np.zeros_like(V)
    for j in range(k):
        v[j, N - k + j] = 1
    np.testing.assert_almost_equal(oracle.subspace_dist(V, v, S), 0, decimal=5)


def test_powermethod_factor_and_history_contract():
    """tests/test_iter_methods.py:128-155 (factor scales S by sqrt(factor)), :477-497 (history list lengths)."""
    rng = np.random.RandomState(1)
    a = rng.rand(50, 40)
    s_true = np.linalg.svd(a, compute_uv=False)
    for factor, div in (("n", 50), ("p", 40), (None, 1)):
        pm = oracle.PowerMethod(k=3, tol=1e-12, factor=factor, lmbd=0, max_iter=200, sub_svd_start=False)
        _, S, _ = pm.svd(a)
        np.testing.assert_array_almost_equal(S * np.sqrt(div), s_true[:3])
    pm = oracle.PowerMethod(k=3, tol=[1e-16, 1e-16, 1e-16], scoring_method=["q-vals", "rmse", "v-subspace"],
                            max_iter=7, lmbd=0, sub_svd_start=False)
    pm.svd(a)
    assert pm.num_iter == 7
    for key in ("q-vals", "rmse", "v-subspace"):
        assert len(pm.history.acc[key]) == 7
    assert len(pm.history.iter["S"]) == 7 and len(pm.history.iter["V"]) == 7 and pm.history.iter["U"] == []
    assert pm.history.acc["q-vals"][0] == float("inf") and pm.history.acc["v-subspace"][0] == float("inf")


def test_powermethod_nan_raises_linalgerror_unless_snp_array():
    """tests/test_iter_methods.py:318-328."""
    rng = np.random.RandomState(2)
    a = rng.randint(0, 3, size=(60, 50)).astype(float)
    a[0, 0] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        oracle.PowerMethod(k=2, buffer=2, sub_svd_start=False).svd(a)
    oracle.PowerMethod(k=2, buffer=2, sub_svd_start=False).svd(oracle.make_snp_array(a))


def test_sspm_cases():
    """tests/test_iter_methods.py:500-566."""
    N, P, k = 20000, 100, 10
    array = np.zeros((N, P))
    s_orig = np.linspace(1, 2, P)
    array[:P, :] = np.diag(s_orig)
    array[N - 1, :] = 1
    U, S, V = np.linalg.svd(array, full_matrices=False)
    sb = oracle.SuccessiveBatchedPowerMethod(k=k, sub_svd_start=True, tol=1e-12, factor=None)
    U_pm, S_pm, V_pm = sb.svd(array, rng=np.random.default_rng(0))
    np.testing.assert_almost_equal(oracle.subspace_dist(U_pm, U[:, :k], S[:k]), 0)
    np.testing.assert_almost_equal(oracle.subspace_dist(V_pm, V[:k], S[:k]), 0)
    np.testing.assert_array_almost_equal(S[:k], S_pm)
    for sub_S in sb.history.iter["S"][:-1]:
        np.testing.assert_array_almost_equal(s_orig[::-1][:k], sub_S, decimal=5)
    # case 3: geometric spectrum, two criteria, per-batch known answers
    N, f = 500, .1
    s_orig = np.array([1.01 ** i for i in range(1, N + 1)])
    sb = oracle.SuccessiveBatchedPowerMethod(k=k, sub_svd_start=True, tol=[1e-14, 1e-14], f=f,
                                             scoring_method=["q-vals", "v-subspace"], factor=None)
    U_pm, S_pm, V_pm = sb.svd(np.diag(s_orig), rng=np.random.default_rng(0))
    np.testing.assert_array_almost_equal(s_orig[-k:][::-1], S_pm)
    for i, (sub_S, sub_V) in enumerate(zip(sb.history.iter["S"], sb.history.iter["V"])):
        s = sorted(s_orig[:int(f * (i + 1) * N)], reverse=True)[:k]
        np.testing.assert_array_almost_equal(s, sub_S)
        v = np.zeros_like(sub_V)
        for j in range(k):
            v[j, int(f * (i + 1) * N) - k + j] = 1
        np.testing.assert_almost_equal(oracle.subspace_dist(v, sub_V, sub_S), 0)


def test_metrics_known_answers():
    """tests/test_metrics.py:13-173."""
    rng = np.random.RandomState(3)
    for N in (1, 101, 501):
        for K in range(1, min(12, N), 5):
            V1, _ = np.linalg.qr(rng.rand(N, K))
            S = rng.randn(K) + 10
            for P in (1, 2, 3):
                np.testing.assert_almost_equal(oracle.subspace_dist(V1, V1.copy(), S, P), 0, decimal=5)
    for N in range(2, 10):
        I = np.eye(N)
        np.testing.assert_almost_equal(oracle.subspace_dist(I, rng.permutation(I.T).T, rng.randn(N) + 10), 0, decimal=5)
    V1 = np.array([[1, 0], [0, 1], [0, 0]])
    V2 = np.array([[1., 0], [0, 0], [0, 1]])
    for p in np.logspace(0, 2, 5):
        for a in np.linspace(0, 1, 11):
            np.testing.assert_almost_equal(oracle.subspace_dist(V1, V2, np.array([1, a]), power=p),
                                           (a ** p) / (1 + a ** p), decimal=5)
    a = rng.randn(100, 50)
    U, S, V = np.linalg.svd(a, full_matrices=False)
    with pytest.raises(ValueError):
        oracle.subspace_dist(U, a.dot(V), S)
    for a_ in np.linspace(0, 1, 11):
        si, sj = np.array([1, a_]), np.array([1, 1])
        np.testing.assert_almost_equal(oracle.q_value_converge(si, sj, norm=2, scale=False), 1 - a_)
        np.testing.assert_almost_equal(oracle.q_value_converge(si, sj, norm=2, scale=True),
                                       2 * (1 - a_) / (np.sqrt(2) + np.sqrt(1 + a_ ** 2)))
        np.testing.assert_almost_equal(oracle.q_value_converge(si, sj, norm=1, scale=True), 2 * (1 - a_) / (3 + a_))
    with pytest.raises(ValueError):
        oracle.q_value_converge(np.array([1, 1, 1]), np.array([1, 1]))
    for s in np.linspace(0.1, 1, 4):
        for N in range(2, 8):
            for K in range(1, N):
                np.testing.assert_almost_equal(oracle.rmse_k(np.eye(N), np.eye(N)[:, :K], s * np.ones(K)),
                                               (1 - s) / np.sqrt(N * N))
    n, p = 10, 7
    a = rng.random_sample((n, p))
    U, S, _ = np.linalg.svd(a.dot(a.T), full_matrices=False)
    np.testing.assert_almost_equal(oracle.rmse_k(a, U, S), 0)


def test_matrix_ops_known_answers():
    """tests/test_matrix_ops.py:14-144."""
    rng = np.random.RandomState(4)
    for N in range(2, 6):
        d, x = rng.rand(N), rng.rand(N, 3)
        np.testing.assert_array_almost_equal(oracle.diag_dot(d, x), np.diag(d).dot(x))
        np.testing.assert_array_almost_equal(oracle.diag_dot(d, x[:, 0]), np.diag(d).dot(x[:, 0]))
    with pytest.raises(ValueError):
        oracle.diag_dot(rng.rand(3), rng.rand(4, 2))
    N, K = 30, 6
    x = np.zeros((N, K))
    sv = np.arange(K, 0, -1.0)
    x[:K, :K] = np.diag(sv)
    U, S, V = oracle.subspace_to_SVD(x, k=3, sqrt_s=False)
    np.testing.assert_array_almost_equal(S, sv[:3])
    np.testing.assert_array_almost_equal(np.abs(U), np.eye(N)[:, :3])
    U, S, V = oracle.subspace_to_SVD(x, k=3, sqrt_s=True)
    np.testing.assert_array_almost_equal(S, np.sqrt(sv[:3]))
    a = rng.rand(N, 12)
    Ua, Sa, Va = np.linalg.svd(a, full_matrices=False)
    Vk = oracle.subspace_to_V(Ua[:, :4] * Sa[:4], oracle.as_operator(a), k=4)
    np.testing.assert_almost_equal(oracle.subspace_dist(Vk, Va[:4], Sa[:4]), 0)


def test_partitions_and_sampling():
    """tests/test_random.py:8-134."""
    for n in (10, 100, 1001):
        for f in (0.1, 0.25, 0.5):
            parts = oracle.array_constant_partition((n, 7), f=f, min_size=3)
            assert parts[0].start == 0 and parts[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
            cum = oracle.cumulative_partition(parts)
            assert all(c.start == 0 for c in cum) and [c.stop for c in cum] == [q.stop for q in parts]
    with pytest.raises(ValueError):
        oracle.array_constant_partition((10, 2), f=0.7)
    with pytest.raises(ValueError):
        oracle.array_constant_partition((10, 2), f=0)
    i1, i2 = oracle.array_split((100, 3), f=0.3, seed=42)
    assert len(i1) == 30 and len(i2) == 70 and set(i1) | set(i2) == set(range(100))
    j1, _ = oracle.array_split((100, 3), f=0.3, seed=42)
    assert np.array_equal(i1, j1) and np.array_equal(i1, np.sort(i1))
    with pytest.raises(ValueError):
        oracle.array_split((100, 3), f=1.5)


def test_moments_imputation_and_snp_array():
    """tests/test_decomp_utils.py:31-149."""
    rng = np.random.RandomState(5)
    g = rng.randint(0, 3, size=(40, 25)).astype(float)
    g[rng.rand(40, 25) < 0.1] = np.nan
    m, s = oracle.get_array_moments(g, std_method="norm")
    np.testing.assert_array_almost_equal(m, np.nanmean(g, axis=0))
    np.testing.assert_array_almost_equal(s, np.nanstd(g, axis=0))
    m, s = oracle.get_array_moments(g, std_method="binom")
    np.testing.assert_array_almost_equal(s, np.sqrt(2 * (m / 2) * (1 - m / 2)))
    filled, mask = oracle.mask_imputation(g, m)
    imputed = np.where(np.isnan(g), m, g)
    np.testing.assert_array_almost_equal(filled + mask.toarray(), imputed)
    with pytest.raises(ValueError):
        oracle.mask_imputation(np.ones((3, 3)))
    snp = oracle.make_snp_array(g, std_method="norm")
    dense = snp.toarray()
    np.testing.assert_array_almost_equal(dense.mean(axis=0), 0)
    np.testing.assert_array_almost_equal(dense, (imputed - m) / np.nanstd(g, axis=0))


def test_scaled_center_array_grid():
    """tests/test_scaledcenterarray.py:37-220, 281-361, 423-495."""
    rng = np.random.RandomState(6)
    for N in range(2, 5):
        for P in range(2, 5):
            a = rng.rand(N, P)
            for scale in (True, False):
                for center in (True, False):
                    for factor in (None, "n", "p"):
                        sa = oracle.ScaledCenterArray(scale=scale, center=center, factor=factor)
                        sa.fit(a)
                        b = a - a.mean(axis=0) if center else a
                        b = b / a.std(axis=0) if scale else b
                        f = {None: 1, "n": N, "p": P}[factor]
                        np.testing.assert_array_almost_equal(sa.array, b)
                        for K in range(1, 4):
                            x, y = rng.rand(N, K), rng.rand(P, K)
                            np.testing.assert_array_almost_equal(sa.sym_mat_mult(x), b.dot(b.T.dot(x)) / f)
                            np.testing.assert_array_almost_equal(sa.dot(y), b.dot(y))
                            np.testing.assert_array_almost_equal(sa.T.dot(x), b.T.dot(x))
                        assert sa.T.T is sa
    sa = oracle.ScaledCenterArray()
    sa.fit(rng.rand(10, 4))
    sub = sa[0:5, :]
    np.testing.assert_array_almost_equal(sub.center_vector, sa._array[0:5].mean(axis=0))   # refit on the selection
    with pytest.raises(ValueError):
        oracle.ScaledCenterArray(factor="q")
    with pytest.raises(ValueError):
        oracle.ScaledCenterArray(std_dist="poisson")
    with pytest.raises(NotImplementedError):
        sa.T.sym_mat_mult(rng.rand(4, 2))
    clone = oracle.ScaledCenterArray.fromScaledArray(rng.rand(6, 4), sa)
    np.testing.assert_array_equal(clone.center_vector, sa.center_vector)
    with pytest.raises(AttributeError):
        oracle.ScaledCenterArray.fromScaledArray(rng.rand(6, 4), oracle.ScaledCenterArray())
    low = np.ones((5, 3))
    with pytest.warns(UserWarning):
        m = oracle.ArrayMoment(low, warn=True)
        m.fit()
    np.testing.assert_array_equal(m.scale_vector, np.ones(3))       # std <= 1e-6 -> 1 (scaled_legacy.py:78-83)


def test_init_methods():
    """tests/test_init_methods.py:10-44."""
    rng = np.random.RandomState(7)
    a = rng.rand(200, 50)
    U, S, V = np.linalg.svd(a, full_matrices=False)
    x = oracle.v_init(oracle.as_operator(a), V[:5].T)
    np.testing.assert_almost_equal(oracle.subspace_dist(x, U[:, :5], S[:5]), 0)
    x = oracle.sub_svd_init(oracle.as_operator(a), k=5, warm_start_row_factor=10)
    assert x.shape == (200, 5)
    np.testing.assert_array_almost_equal(x.T @ x, np.eye(5))


def _separated_matrix():
    """deterministic 600 x 400 matrix with a separated spectrum 100 * 0.7^i over a flat noise floor (generator committed
    here)"""
    rng = np.random.RandomState(7)
    u, _ = np.linalg.qr(rng.randn(600, 12))
    v, _ = np.linalg.qr(rng.randn(400, 12))
    return (u * (100.0 * 0.7 ** np.arange(12))) @ v.T + 1e-3 * rng.randn(600, 400)


SEPARATED_S = np.array([100.0, 70.0, 49.0, 34.3, 24.01, 16.807])      # the planted values (the noise moves them by < 1e-4)


def test_golden_separated_spectrum_every_singular_value():
    """the ex1 fixture pins only S[0] tightly (its S[1:] is one cluster).  With a separated spectrum the
    oracle's PowerMethod must give EVERY singular value and vector of LAPACK's SVD to 1e-8, for all three metrics."""
    a = _separated_matrix()
    U, S, Vt = np.linalg.svd(a, full_matrices=False)
    np.testing.assert_allclose(S[:6], SEPARATED_S, rtol=1e-4)          # fixture sanity: planted values + 1e-3 noise
    for scoring in ("q-vals", "rmse", "v-subspace"):
        pm = oracle.PowerMethod(k=6, buffer=6, tol=1e-13, scoring_method=scoring, max_iter=80, factor=None, lmbd=0,
                                sub_svd_start=False)
        U_pm, S_pm, V_pm = pm.svd(a, x0=np.random.RandomState(1).randn(600, 12))
        np.testing.assert_allclose(S_pm, S[:6], rtol=1e-8)
        assert oracle.subspace_dist(U_pm, U[:, :6], S[:6]) < 1e-10
        assert oracle.subspace_dist(V_pm.T, Vt[:6].T, S[:6]) < 1e-10
