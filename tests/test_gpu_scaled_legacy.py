"""The reference's ScaledCenterArray tests (tests/test_scaledcenterarray.py, line numbers in the docstrings) against this
build, on float matrices (dense path) and on genotype matrices (2-bit tile store path)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _np(a):
    return np.asarray(a)


def _cases(rng, kind, N, P):
    return rng.random((N, P)) + 1 if kind == "float" else rng.integers(0, 3, size=(N, P)).astype(np.float64)


@pytest.mark.parametrize("kind", ["float", "geno"])
def test_ScaledArray_sym_mat_mult_dot_T_grid(cuda, kind):
    """:37-88 (sym_mat_mult), :135-176 (dot), :179-220 (T.dot): tiny-shape grids over scale x center x factor x K x
    squeeze x fit_x."""
    from lmdec_b200.array import ScaledCenterArray
    rng = np.random.default_rng(0)
    dec = 6 if kind == "float" else 4            # genotype path: 22-bit fixed point
    shapes = [(2, 2), (3, 4), (4, 3)] if kind == "float" else [(40, 12), (33, 20)]
    for N, P in shapes:
        array = _cases(rng, kind, N, P)
        sd = np.std(array, axis=0)
        sd[sd <= 1e-6] = 1
        std, mu = np.diag(1 / sd), np.mean(array, axis=0)
        for factor, f in ((None, 1), ("n", N), ("p", P)):
            for K in (1, 2, 4):
                for squeeze in (True, False):
                    x, y = rng.random((N, K)), rng.random((P, K))
                    if squeeze:
                        x, y = np.squeeze(x), np.squeeze(y)
                    for fit_x in (x, None):
                        for scale in (False, True):
                            for center in (False, True):
                                b = array - mu if center else array
                                b = b.dot(std) if scale else b
                                sa = ScaledCenterArray(scale=scale, center=center, factor=factor, warn=False)
                                sa.fit(array, x=fit_x)
                                res = _np(sa.sym_mat_mult(x))
                                assert res.shape == x.shape
                                np.testing.assert_array_almost_equal(b.dot(b.T.dot(x)) / f, res, decimal=dec)
                                np.testing.assert_array_almost_equal(b.dot(y), _np(sa.dot(y)), decimal=dec)
                                np.testing.assert_array_almost_equal(b.T.dot(x), _np(sa.T.dot(x)), decimal=dec)
                                np.testing.assert_array_almost_equal(b, _np(sa.array), decimal=10)


def test_ScaledArray_state_and_getitem(cuda):
    """:223-246 (T.T is self, shared state), :318-361 (__getitem__ refits on the selection), :375-420 (shapes),
    :423-434 (std_dist), :437-495 (fromScaledArray)."""
    from lmdec_b200.array import ArrayMoment, ScaledCenterArray
    rng = np.random.default_rng(1)
    for N, P in ((3, 3), (4, 3), (4, 4)):
        array = rng.random((N, P)) + 1
        sa = ScaledCenterArray(scale=True, center=True)
        sa.fit(array)
        assert sa.T.T is sa and sa.shape == (N, P) and sa.T.shape == (P, N)
        assert sa.T._array is sa._array and sa.T._array_moment is sa._array_moment
        for sub_N in range(2, N):
            sub = array[0:sub_N, :]
            b = (sub - sub.mean(axis=0)).dot(np.diag(1 / sub.std(axis=0)))
            y = rng.random((P, 2))
            np.testing.assert_array_almost_equal(b.dot(y), _np(sa[0:sub_N, :].dot(y)))
            assert sa[0:sub_N, :].shape == (sub_N, P)
    g = rng.integers(0, 3, size=(50, 8)).astype(float)
    sb = ScaledCenterArray(std_dist="binom")
    sb.fit(g)
    p = g.mean(axis=0) / 2
    np.testing.assert_array_almost_equal(_np(sb.scale_vector), 1 / np.sqrt(2 * p * (1 - p)))
    new = rng.integers(0, 3, size=(7, 8)).astype(float)
    clone = ScaledCenterArray.fromScaledArray(new, sb)
    np.testing.assert_array_equal(_np(clone.center_vector), _np(sb.center_vector))
    y = rng.random((8, 3))
    np.testing.assert_array_almost_equal(_np(clone.dot(y)), ((new - g.mean(axis=0)) * _np(sb.scale_vector)).dot(y), decimal=4)
    with pytest.raises(ValueError):
        ScaledCenterArray.fromScaledArray(rng.random((7, 5)), sb)
    with pytest.raises(AttributeError):
        ScaledCenterArray.fromScaledArray(new, ScaledCenterArray())
    with pytest.raises(ValueError):
        ScaledCenterArray(factor="x")
    with pytest.raises(ValueError):
        ScaledCenterArray(std_dist="poisson")
    with pytest.raises(ValueError):
        ScaledCenterArray().fit(rng.random(5))
    with pytest.warns(UserWarning):                 # :498-559 ArrayMoment low-variance guard
        m = ArrayMoment(__import__("lmdec_b200.array.operators", fromlist=["DenseOperator"]).DenseOperator(np.ones((5, 3))))
        m.fit()
    np.testing.assert_array_equal(_np(m.scale_vector), np.ones(3))
    with pytest.raises(ValueError):
        ArrayMoment(None, std_dist="poisson")


def test_kernel_edge_shapes(cuda):
    """ragged / extreme shapes through the operator: one sample, one SNP, K = 1 and the widest K, every plane count,
    all-missing rows and a store whose every code is missing in one column (mean NaN -> LinAlgError downstream)."""
    import torch
    from helpers import oracle, synth_genotypes
    from lmdec_b200.decomp import make_snp_array
    rng = np.random.default_rng(2)
    for n, p, K, planes in ((1, 300, 1, 3), (300, 1, 1, 3), (130, 257, 42, 3), (65, 520, 32, 4), (257, 130, 64, 2),
                            (129, 1025, 7, 1)):
        g = synth_genotypes(n, p, 0.05 if n > 8 else 0.0, seed=n + p)
        g[:, 0] = np.where(np.isnan(g[:, 0]), 1.0, g[:, 0])
        ref = oracle.make_snp_array(g, mean=True, std=False)
        op = make_snp_array(g, mean=True, std=False, planes=planes)
        t, y = rng.standard_normal((p, K)), rng.standard_normal((n, K))
        tol = {1: 1e-1, 2: 3e-4, 3: 2e-5, 4: 1e-7}[planes]   # 8P-3 bits below the column maximum (missing data)
        x_ref, w_ref = np.asarray(ref.dot(t)), np.asarray(ref.T.dot(y))
        # the fixed point is relative to the operand's column maximum: scale the tolerance by the product's natural size
        x_scale = max(np.max(np.abs(x_ref)), np.max(np.abs(t)) * np.sqrt(p))
        w_scale = max(np.max(np.abs(w_ref)), np.max(np.abs(y)) * np.sqrt(n))
        assert np.max(np.abs(_np(op.dot(t)) - x_ref)) <= tol * x_scale
        assert np.max(np.abs(_np(op.T.dot(y)) - w_ref)) <= tol * w_scale
    # K * planes beyond the 128 tensor-core columns of one launch: column blocks (round 2; it raised in round 1)
    g43 = synth_genotypes(200, 300, 0.0, seed=1)
    t43 = rng.standard_normal((300, 43))
    x43 = np.asarray(oracle.make_snp_array(g43).dot(t43))
    assert np.max(np.abs(_np(make_snp_array(g43, planes=3).dot(t43)) - x43)) <= 2e-5 * np.max(np.abs(x43))
    g = synth_genotypes(100, 50, 0.0, seed=3)
    g[:, 7] = np.nan                                 # a SNP with no observed genotype: its mean is NaN
    from lmdec_b200 import PowerMethod
    with pytest.raises(np.linalg.LinAlgError):
        PowerMethod(k=2, buffer=2, sub_svd_start=False).svd(make_snp_array(g), verbose=False)
    with pytest.raises(ValueError):                  # values other than 0/1/2/NaN are not genotypes -> dense operator
        from lmdec_b200.store import GenotypeStore
        GenotypeStore.from_dense(np.full((4, 4), 5.0))
