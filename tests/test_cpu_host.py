"""CPU checks of the boundary and the host logic: the C-ABI library loads and exports every declared symbol, the
product path refuses to run without a CUDA device, host-side helpers agree with the oracle."""
import os
import re

import numpy as np
import pytest

from helpers import ROOT, oracle


def test_capi_exports_every_declared_symbol():
    from lmdec_b200 import _build, _capi
    header = open(os.path.join(ROOT, "include", "lmdec_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lmdec_[a-z0-9_]+)\s*\(", header))   # (typedefs / enums have no parenthesis)
    assert len(declared) >= 15
    lib = _capi.load()
    assert set(_capi.SIGNATURES) == declared, set(_capi.SIGNATURES) ^ declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lmdec_abi_version() == _capi.ABI_VERSION
    assert lib.lmdec_source_hash().decode() == _build.source_hash()      # the shipped binary is built from these sources
    # pure host helpers of the ABI (no device needed)
    assert lib.lmdec_store_bytes(1, 1) == 8192
    assert lib.lmdec_store_bytes(2504, 1_000_000) == 20 * 3907 * 8192
    assert lib.lmdec_image_bytes(1_000_000, 64) == 3907 * 8 * 64 * 32
    assert lib.lmdec_dense_store_bytes(1, 1) == 16384                       # one tile of 128 rows x 32 float32
    assert lib.lmdec_dense_store_bytes(1_000_000, 4000) == 7813 * 125 * 16384
    assert lib.lmdec_dense_image_bytes(4000, 112) == 125 * 4 * 112 * 32
    # argument validation of the round-2 entry points happens before anything touches the device
    assert lib.lmdec_gram_reduce(None, 20, 100, 20, None, None, None) != 0
    assert b"multicast" in lib.lmdec_last_error()
    assert lib.lmdec_dense_matmul(None, None, 1, 1, 1, 1, None, 0, 1, 1, None, None, None, 1, 0, None) != 0
    assert lib.lmdec_operator_apply_reduce(None, None, 1, 1, 1, 1, 0, 9, None, 1, 4, 3, None, None, None, None, None,
                                           None, None, None, None, None, None) != 0


def test_library_is_sm100a_tcgen05():
    """the shipped kernel is the tensor-core one: its SASS holds UTC*MMA / LDTM / STTM / UBLKCP."""
    import shutil
    import subprocess
    from lmdec_b200 import _capi
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _capi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    # UTCIMMA: kind::i8 (genotypes); UTCHMMA: kind::tf32 (dense float32 operands).  (multimem.red of the fused
    # reductions shows up as a plain REDG on a multicast address: nothing to grep for.)
    for mnemonic in ("UTCIMMA", "UTCHMMA", "STTM", "LDTM", "UBLKCP", "USETMAXREG"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        make_snp_array(np.zeros((8, 8)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PowerMethod(k=1, buffer=1).svd(np.random.rand(8, 8), verbose=False)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lmdec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "lmdec_oracle" not in src and "oracle/" not in src, os.path.join(dirpath, f)


def test_constructor_validation_matches_reference():
    """lmdec/decomp/iter_methods.py:59-86, 307-317."""
    from lmdec_b200 import PowerMethod, SuccessiveBatchedPowerMethod
    for cls, pm_cls in ((PowerMethod, oracle.PowerMethod),):
        for kw in (dict(max_iter=0), dict(max_iter=2.5), dict(time_limit=0), dict(k=0), dict(buffer=-1),
                   dict(init_row_sampling_factor=0), dict(tol=[1, 2]), dict(scoring_method="x"),
                   dict(scoring_method=["q-vals", "rmse"], tol=[.1])):
            with pytest.raises(ValueError):
                cls(**kw)
            with pytest.raises(ValueError):
                pm_cls(**kw)
    pm = PowerMethod(scoring_method=["q-vals", "rmse"], tol=[.1, 1e-16])
    assert pm.tol == [.1, 1e-16] and pm.k == 10 and pm.buffer == 10 and pm.max_iter == 50
    sb = SuccessiveBatchedPowerMethod()
    assert sb.f == .1 and sb.lmbd == .1 and sb.max_iter == 50


def test_partition_helpers_equal_oracle():
    from lmdec_b200.array import array_constant_partition, array_split, cumulative_partition
    for n in (10, 37, 1000, 50000):
        for f in (0.1, 0.2, 0.5):
            a = array_constant_partition((n, 5), f=f, min_size=4)
            assert a == oracle.array_constant_partition((n, 5), f=f, min_size=4)
            assert cumulative_partition(a) == oracle.cumulative_partition(a)
    for seed in (1, 42):
        x, y = array_split((500, 3), f=0.13, seed=seed)
        ox, oy = oracle.array_split((500, 3), f=0.13, seed=seed)
        assert np.array_equal(x, ox) and np.array_equal(y, oy)


def test_whitening_factor_host_logic():
    """CholeskyQR host step: r_inv^T G r_inv = I, rank-deficient fallback, NaN -> LinAlgError."""
    from lmdec_b200.array.linalg import _whiten_factor
    rng = np.random.default_rng(0)
    x = rng.standard_normal((200, 6)) * np.array([1, 10, 100, 1e3, 1e4, 1e5])
    G = x.T @ x
    r_inv, r, deficient = _whiten_factor(G)
    assert not deficient
    np.testing.assert_allclose(r_inv.T @ G @ r_inv, np.eye(6), atol=1e-8)
    np.testing.assert_allclose(r @ r_inv, np.eye(6), atol=1e-10)
    x[:, 3] = x[:, 2]                           # exactly rank deficient
    r_inv, r, deficient = _whiten_factor(x.T @ x)
    assert deficient and (np.abs(r_inv).sum(axis=0) == 0).sum() == 1
    with pytest.raises(np.linalg.LinAlgError):
        _whiten_factor(np.full((3, 3), np.nan))
