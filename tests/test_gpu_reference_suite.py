"""The reference's own driver tests (tests/test_iter_methods.py, line numbers in each docstring) run against this build.
Ground truth is dense LAPACK, the comparator is the oracle's restatement of lmdec's subspace_dist."""
import numpy as np
import pytest

from helpers import oracle

pytestmark = pytest.mark.gpu


def _np(a):
    return np.asarray(a)


def subspace_dist(a, b, s):
    return oracle.subspace_dist(_np(a), _np(b), _np(s))


def test_PowerMethod_case1(cuda):
    """:12-33 exact recovery for k = 1..9 through make_snp_array on a float matrix (rmse criterion)."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    n, p = 100, 80
    array = np.random.RandomState(0).rand(n, p)
    sc = (array - array.mean(axis=0)).dot(np.diag(1 / array.std(axis=0)))
    U, S, V = np.linalg.svd(sc, full_matrices=False)
    snp = make_snp_array(array, mean=True, std=True, std_method="norm", mask_nan=False, dtype="float64")
    for k in range(1, 10):
        PM = PowerMethod(k=k, tol=1e-9, scoring_method="rmse", max_iter=100, sub_svd_start=False,
                         init_row_sampling_factor=1, factor=None, lmbd=0)
        U_k_PM, S_k_PM, V_k_PM = PM.svd(snp, verbose=False)
        np.testing.assert_array_almost_equal(S[:k], _np(S_k_PM))
        assert V_k_PM.shape == (k, p) and U_k_PM.shape == (n, k)
        np.testing.assert_almost_equal(subspace_dist(V[:k], V_k_PM, S_k_PM), 0)
        np.testing.assert_almost_equal(subspace_dist(U[:, :k], U_k_PM, S_k_PM), 0)


def test_PowerMethod_case2_monotone_in_tol(cuda):
    """:36-66 the error shrinks with the tolerance and reaches 1e-9 / 1e-12."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    array = np.random.RandomState(1).rand(100, 100)
    sc = (array - array.mean(axis=0)).dot(np.diag(1 / array.std(axis=0)))
    snp = make_snp_array(array, mean=True, std=True, std_method="norm", mask_nan=False, dtype="float64")
    U, S, _ = np.linalg.svd(sc.dot(sc.T), full_matrices=False)
    _, _, V = np.linalg.svd(sc.T.dot(sc), full_matrices=False)
    S = np.sqrt(S)
    k = 10
    prev = [float("inf")] * 3
    for t in np.logspace(0, -12, 8):
        PM = PowerMethod(k=k, tol=t, scoring_method="q-vals", max_iter=100, factor=None, lmbd=0)
        Uk, Sk, Vk = PM.svd(snp, verbose=False)
        cur = [subspace_dist(U[:, :k], Uk, S[:k]), subspace_dist(V[:k], Vk, S[:k]), np.linalg.norm(S[:k] - _np(Sk))]
        assert all(c <= p_ + 1e-12 for c, p_ in zip(cur, prev))
        prev = cur
    assert cur[0] <= 1e-8 and cur[1] <= 1e-8 and cur[2] <= 1e-10


def test_PowerMethod_subsvd_finds_eigenvectors(cuda):
    """:69-86 geometric spectrum, default warm start with noise."""
    from lmdec_b200 import PowerMethod
    N, k = 1000, 10
    s_orig = np.array([1.01 ** i for i in range(1, N + 1)])
    PM = PowerMethod(tol=1e-16, factor=None, lmbd=.1, max_iter=100)
    U_PM, S_PM, V_PM = PM.svd(np.diag(s_orig), verbose=False, seed=0)
    np.testing.assert_array_almost_equal(s_orig[-k:][::-1], _np(S_PM), decimal=0)
    v = np.zeros((k, N))
    for j in range(k):
        v[j, N - k + j] = 1
    np.testing.assert_almost_equal(subspace_dist(V_PM, v, S_PM), 0, decimal=5)


def test_PowerMethod_all_tols_agree(cuda):
    """:103-125."""
    from lmdec_b200 import PowerMethod
    array = np.random.RandomState(2).rand(100, 80)
    res = []
    for method, tol in (("q-vals", 1e-9), ("rmse", 1e-4), ("v-subspace", 1e-9)):
        res.append(PowerMethod(k=10, tol=tol, scoring_method=method, max_iter=100, lmbd=0).svd(array, verbose=False))
    (U_q, S_q, V_q), (U_r, S_r, V_r), (U_v, S_v, V_v) = res
    np.testing.assert_array_almost_equal(_np(S_q), _np(S_r))
    np.testing.assert_array_almost_equal(_np(S_q), _np(S_v))
    for a, b in ((U_q, U_r), (U_q, U_v), (V_q, V_r), (V_q, V_v)):
        np.testing.assert_almost_equal(subspace_dist(a, b, S_q), 0)


def test_PowerMethod_factor(cuda):
    """:128-155."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    n, p, k = 100, 80, 10
    array = np.random.RandomState(3).rand(n, p)
    sym = array.dot(array.T)
    for f, factor in (("n", n), ("p", p), (None, 1)):
        U, S, _ = np.linalg.svd(sym / factor, full_matrices=False)
        S = np.sqrt(S)
        snp = make_snp_array(array, mean=False, std=False, std_method="norm", mask_nan=False, dtype="float64")
        U_PM, S_PM, _ = PowerMethod(k=k, tol=1e-9, scoring_method="q-vals", max_iter=100, factor=f,
                                    lmbd=0).svd(snp, verbose=False)
        np.testing.assert_array_almost_equal(S[:k], _np(S_PM))
        assert U_PM.shape == (n, k)
        np.testing.assert_almost_equal(subspace_dist(U[:, :k], U_PM, S_PM), 0)


def test_PowerMethod_multi_tol(cuda):
    """:188-213 convergence when ANY criterion is met."""
    from lmdec_b200 import PowerMethod
    array = np.random.RandomState(4).rand(100, 70)
    methods = ["q-vals", "rmse", "v-subspace"]
    for tols in ([1e-1, 1e-1, 1e-1], [1e-16, 1e-4, 1 - 16], [1e-6, 1e-16, 1e-16], [1e-16, 1e-16, 1e-6]):
        PM = PowerMethod(k=10, tol=tols, scoring_method=methods, factor=None, lmbd=0)
        PM.svd(array, verbose=False)
        assert any(min(PM.history.acc[m]) <= t for m, t in zip(methods, tols))
    PM = PowerMethod(k=10, tol=[1e-6, 1e-16, 1e-16], scoring_method=methods, factor=None, lmbd=0)
    PM.svd(array, verbose=False)
    assert min(PM.history.acc["q-vals"]) <= 1e-6 and min(PM.history.acc["rmse"]) >= 1e-16 \
        and min(PM.history.acc["v-subspace"]) >= 1e-16


def test_PowerMethod_parameters(cuda):
    """:216-251, 263-265, 299-315 parameter validation, max_iter, time limit, bad arrays."""
    from lmdec_b200 import PowerMethod
    rs = np.random.RandomState(5)
    for kw in (dict(k=0), dict(k=-1), dict(max_iter=0), dict(time_limit=0), dict(buffer=-1),
               dict(init_row_sampling_factor=0)):
        with pytest.raises(ValueError):
            PowerMethod(**kw)
    with pytest.raises(ValueError):
        PowerMethod(k=100).svd(rs.rand(50, 50), verbose=False)
    PM = PowerMethod(max_iter=1)
    PM.svd(rs.rand(100, 100), verbose=False)
    assert PM.num_iter == 1 and len(PM.history.iter["S"]) == 1
    PM = PowerMethod(time_limit=2, max_iter=int(1e10), tol=1e-20)
    PM.svd(rs.rand(100, 100), verbose=False)
    assert 1.5 <= PM.time <= 3.5
    PowerMethod(init_row_sampling_factor=4).svd(rs.rand(100, 100), verbose=False)
    with pytest.raises(ValueError):
        PowerMethod().svd(rs.rand(100), verbose=False)
    with pytest.raises(ValueError):
        PowerMethod().svd(rs.rand(100, 100, 100), verbose=False)
    for start in (True, False):                                      # :359-365 reuse of one instance
        PM = PowerMethod(sub_svd_start=start)
        PM.svd(rs.randn(100, 50), verbose=False)
        PM.svd(rs.randn(110, 60), verbose=False)


def test_PowerMethod_buffer_and_start(cuda):
    """:263-296 a larger buffer converges faster; warm and Gaussian starts agree."""
    from lmdec_b200 import PowerMethod
    rs = np.random.RandomState(6)
    array = rs.rand(400, 400)
    U, S, V = np.linalg.svd(array)
    S[0:15] = 1e3 + S[0:15]
    array = U.dot(np.diag(S).dot(V))
    for method in ("q-vals", "rmse"):
        PM1 = PowerMethod(buffer=2, max_iter=3, tol=1e-16, scoring_method=method)
        PM2 = PowerMethod(buffer=30, max_iter=3, tol=1e-16, scoring_method=method)
        PM1.svd(array, verbose=False, seed=1)
        PM2.svd(array, verbose=False, seed=1)
        assert PM1.history.acc[method][-1] > PM2.history.acc[method][-1]
    a = rs.rand(100, 100)
    _, S1, _ = PowerMethod(sub_svd_start=True, tol=1e-12, lmbd=0).svd(a, verbose=False)
    _, S2, _ = PowerMethod(sub_svd_start=False, tol=1e-12, lmbd=0).svd(a, verbose=False)
    np.testing.assert_array_almost_equal(_np(S1), _np(S2))


def test_PowerMethod_nan_arrays(cuda):
    """:318-328 NaN -> LinAlgError unless the array went through make_snp_array."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    array = np.random.RandomState(7).randn(100, 100)
    array[0, 0] = float("nan")
    for start in (True, False):
        PM = PowerMethod(sub_svd_start=start, max_iter=2)
        with pytest.raises(np.linalg.LinAlgError):
            PM.svd(array, verbose=False)
        clean = make_snp_array(array, mask_nan=True, std_method="norm", dtype="float64")
        PowerMethod(sub_svd_start=start, max_iter=2).svd(clean, verbose=False)


def test_PowerMethod_std_method(cuda):
    """:422-445 integer genotypes through the 2-bit store, 'norm' and 'binom' scaling."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    from lmdec_b200.store import GenotypeOperator
    N, P, k = 1000, 100, 10
    array = np.random.RandomState(8).randint(0, 3, size=(N, P))
    for method in ("norm", "binom"):
        snp = make_snp_array(array, std_method=method)
        assert isinstance(snp, GenotypeOperator)
        U_PM, S_PM, V_PM = PowerMethod(k=k, scoring_method="q-vals", tol=1e-13, factor=None).svd(snp, verbose=False)
        mean = array.mean(axis=0)
        std = array.std(axis=0) if method == "norm" else np.sqrt(2 * (mean / 2) * (1 - mean / 2))
        x = (array - mean).dot(np.diag(1 / std))
        U, S, V = np.linalg.svd(x, full_matrices=False)
        np.testing.assert_almost_equal(subspace_dist(U_PM, U[:, :k], S[:k]), 0, decimal=3)
        np.testing.assert_almost_equal(subspace_dist(V_PM, V[:k], S[:k]), 0, decimal=3)
        np.testing.assert_array_almost_equal(S[:k], _np(S_PM), decimal=2)


def test_PowerMethod_return_history(cuda):
    """:477-497."""
    from lmdec_b200 import PowerMethod
    max_iter = 5
    array = np.random.RandomState(9).random_sample((500, 500))
    PM = PowerMethod(scoring_method=["rmse", "q-vals", "v-subspace"], max_iter=max_iter, tol=[1e-16, 1e-16, 1e-16])
    history = PM.svd(array, return_history=True, verbose=False)
    for m in ("q-vals", "rmse", "v-subspace"):
        assert len(history.acc[m]) == max_iter
    assert len(history.iter["U"]) == 0 and len(history.iter["S"]) == max_iter and len(history.iter["V"]) == max_iter
    assert len(history.iter["last_value"]) == 3
    assert history.time["start"] < history.time["stop"]
    assert len(history.time["step"]) == max_iter and len(history.time["acc"]) == max_iter
    np.testing.assert_array_almost_equal(np.array(history.time["iter"]),
                                         np.array(history.time["acc"]) + np.array(history.time["step"]), decimal=3)


def test_SSPM_cases(cuda):
    """:500-566."""
    from lmdec_b200 import SuccessiveBatchedPowerMethod
    N, P, k = 100000, 100, 10
    array = np.zeros((N, P))
    s_orig = np.linspace(1, 2, P)
    array[:P, :] = np.diag(s_orig)
    array[N - 1, :] = 1
    U, S, V = np.linalg.svd(array, full_matrices=False)
    SSPM = SuccessiveBatchedPowerMethod(k=k, sub_svd_start=True, tol=1e-12, factor=None)
    U_PM, S_PM, V_PM = SSPM.svd(array, seed=0)
    np.testing.assert_almost_equal(subspace_dist(U_PM, U[:, :k], S[:k]), 0)
    np.testing.assert_almost_equal(subspace_dist(V_PM, V[:k], S[:k]), 0)
    np.testing.assert_array_almost_equal(S[:k], _np(S_PM))
    for sub_S in SSPM.history.iter["S"][:-1]:
        np.testing.assert_array_almost_equal(s_orig[::-1][:k], sub_S, decimal=5)
    # :524-538 identity
    SSPM = SuccessiveBatchedPowerMethod(k=k, sub_svd_start=True, tol=1e-12, factor=None)
    SSPM.svd(np.eye(2000), seed=0)
    for sub_S in SSPM.history.iter["S"]:
        np.testing.assert_array_almost_equal(np.ones(k), sub_S)
    # :541-566 geometric spectrum with two criteria
    N, f = 500, .1
    s_orig = np.array([1.01 ** i for i in range(1, N + 1)])
    SSPM = SuccessiveBatchedPowerMethod(k=k, sub_svd_start=True, tol=[1e-14, 1e-14], f=f,
                                        scoring_method=["q-vals", "v-subspace"], factor=None)
    U_PM, S_PM, V_PM = SSPM.svd(np.diag(s_orig), seed=0)
    np.testing.assert_array_almost_equal(s_orig[-k:][::-1], _np(S_PM))
    for i, (sub_S, sub_V) in enumerate(zip(SSPM.history.iter["S"], SSPM.history.iter["V"])):
        s = sorted(s_orig[:int(f * (i + 1) * N)], reverse=True)[:k]
        np.testing.assert_array_almost_equal(s, sub_S)
        v = np.zeros(sub_V.shape)
        for j in range(k):
            v[j, int(f * (i + 1) * N) - k + j] = 1
        np.testing.assert_almost_equal(subspace_dist(v, sub_V, sub_S), 0)


def test_SSPM_return_history(cuda):
    """:569-582."""
    from lmdec_b200 import SuccessiveBatchedPowerMethod
    max_iter, f = 5, .2
    array = np.random.RandomState(10).random_sample((1000, 1000))
    SSPM = SuccessiveBatchedPowerMethod(scoring_method=["rmse", "q-vals", "v-subspace"], f=f, max_sub_iter=max_iter,
                                        tol=[1e-16, 1e-16, 1e-16])
    history = SSPM.svd(array, return_history=True, seed=0)
    sub_iters = int(1 / f)
    assert len(history.time["iter"]) == sub_iters - 1
    assert len(history.iter["S"]) == sub_iters - 1
    assert len(history.iter["sub_svd"]) == sub_iters
    assert len(history.acc["sub_svd_acc"]) == sub_iters - 1
