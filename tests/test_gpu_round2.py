"""Round-2 GPU tests: oracle parity at the bench workload's real shape, multi-rank NCCL parity from inside pytest,
sym_mat_mult_start, the container spelling of the genotype operator, PowerMethod.project, the PLINK loader on a
hand-written fileset, operator persistence in the reference's tree layout, wide operands (k + buffer > 42), the
``import lmdec`` shim and the per-caller context."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, oracle, structured_genotypes, synth_genotypes

pytestmark = pytest.mark.gpu


def _np(a):
    return np.asarray(a)


# --------------------------------------------------------------------------------------------------------------------
# configs[1] at full size: 2,504 x 1,000,000, 5 % missing.  The oracle cannot run the whole matrix in a test, but a
# column slab of A^T y and a product A t whose operand vanishes outside the slab ARE computable on the host exactly.
# --------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def config1(cuda):
    import torch
    from bench import make_shard
    from lmdec_b200.store import GenotypeOperator
    n, p = 2504, 1_000_000
    store = make_shard(torch, cuda, n, p, 0.05, seed=1234)
    mean, std = store.moments("binom")
    return store, GenotypeOperator(store, mean, 1.0 / std)


def test_config1_products_match_oracle_on_a_slab(cuda, config1):
    """full-size kernels (all 7,813 SNP tiles / all 3,907 contraction chunks) against oracle.make_snp_array on the
    same codes: rows [0, 20000) of A^T y, and A t for a t supported on those 20,000 SNPs.  <= 3e-6 of the largest
    entry (3 planes: 2^-22 of each operand column's maximum, plus float32 storage of the wide side)."""
    import torch
    store, op = config1
    n, p, K, slab = 2504, 1_000_000, 20, 20_000
    codes = store.to_dense(transpose=True, rows=slab).cpu().numpy()            # [slab, n] int8, -1 = missing
    g = codes.T.astype(np.float64)
    g[g < 0] = np.nan
    assert abs(np.isnan(g).mean() - 0.05) < 2e-3
    ref = oracle.make_snp_array(g, std_method="binom")
    rng = np.random.default_rng(5)
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
    # 2,048 sampled rows further into the matrix, through the public float64 API
    rows = np.sort(rng.choice(slab, size=2048, replace=False))
    w64 = _np(op.T.dot(y))[rows]
    assert np.max(np.abs(w64 - w_ref[rows])) <= 3e-6 * np.max(np.abs(w_ref))


def test_config1_slab_ten_iterations_match_oracle(cuda, config1):
    """ten power iterations on a 2,504 x 40,000 slab of the bench matrix: S within 1e-4 (north-star tolerance)."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    store, _ = config1
    codes = store.to_dense(transpose=True, rows=40_000).cpu().numpy()
    g = codes.T.astype(np.float32)
    g[g < 0] = np.nan
    x0 = np.random.default_rng(42).standard_normal((2504, 20))
    kw = dict(k=10, buffer=10, tol=1e-30, scoring_method="q-vals", sub_svd_start=False, lmbd=0, max_iter=10)
    Ur, Sr, Vr = oracle.PowerMethod(**kw).svd(oracle.make_snp_array(g, std_method="binom"), x0=x0)
    U, S, V = PowerMethod(**kw).svd(make_snp_array(g, std_method="binom"), x0=x0, verbose=False)
    np.testing.assert_allclose(_np(S), Sr, rtol=1e-4)
    assert oracle.subspace_dist(_np(U), Ur, Sr) <= 1e-4
    assert oracle.subspace_dist(_np(V).T, Vr.T, Sr) <= 1e-4


# --------------------------------------------------------------------------------------------------------------------
# multi-rank NCCL parity, collected by pytest (tests/dist_gpu_check.py was only a script)
# --------------------------------------------------------------------------------------------------------------------
def test_two_rank_nccl_parity(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2): both decompositions of tests/dist_gpu_check.py")
    # twice: with the reduction of A^T q fused into the product's epilogue (NVSwitch multicast; falls back by itself
    # where the box has none) and with the NCCL all-reduce forced
    paths = []
    for fused in ("1", "0"):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", LMDEC_FUSED_REDUCE=fused)
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533",
                              os.path.join(ROOT, "tests", "dist_gpu_check.py")],
                             capture_output=True, text=True, timeout=900, env=env)
        assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
        assert "FAIL" not in res.stdout
        paths.append([ln for ln in res.stdout.splitlines() if "reduction path" in ln])
    assert paths[0] and paths[1] and "NCCL all-reduce" in paths[1][0]


# --------------------------------------------------------------------------------------------------------------------
# sym_mat_mult_start (north_star names it; the reference has no such function -- parity is against the oracle's
# restatement of the definition, which is itself unpinned)
# --------------------------------------------------------------------------------------------------------------------
def test_sym_mat_mult_start_definition_and_range_finder(cuda):
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array, rnormal_start, sym_mat_mult_start
    g = structured_genotypes(500, 2000, miss=0.03, seed=4)
    ref = oracle.make_snp_array(g, std_method="binom")
    op = make_snp_array(g, std_method="binom")
    K = 12
    x_ref = oracle.sym_mat_mult_start(ref, K)
    x = _np(sym_mat_mult_start(op, K))
    assert x.shape == (500, K)
    assert np.max(np.abs(x - x_ref)) <= 1e-5 * np.max(np.abs(x_ref))
    # == A A^T applied to rnormal_start through the public operator protocol
    om = rnormal_start(op, K)
    np.testing.assert_allclose(x, _np(op.dot(op.T.dot(om))), rtol=0, atol=1e-6 * np.max(np.abs(x)))
    # range-finder property: one Gram application pulls the start towards the leading subspace, so the start already
    # captures most of the top-3 left singular subspace, and a PowerMethod started from it converges to the oracle's SVD
    u, s, _ = np.linalg.svd(ref.toarray().astype(np.float64), full_matrices=False)
    q, _ = np.linalg.qr(x)
    q0, _ = np.linalg.qr(_np(om))
    captured = np.linalg.norm(q.T @ u[:, :3]) ** 2 / 3
    assert captured > 0.85 and captured > 10 * np.linalg.norm(q0.T @ u[:, :3]) ** 2 / 3     # 0.89 vs 0.02 for the raw start
    pm = PowerMethod(k=3, buffer=9, tol=1e-9, sub_svd_start="sym_mat_mult", lmbd=0, max_iter=200)
    U, S, V = pm.svd(op, verbose=False)
    np.testing.assert_allclose(_np(S), s[:3] / np.sqrt(500), rtol=1e-4)
    assert oracle.subspace_dist(_np(U), u[:, :3], s[:3]) <= 1e-4


# --------------------------------------------------------------------------------------------------------------------
# the make_snp_array pattern spelled with the containers goes to the tile store
# --------------------------------------------------------------------------------------------------------------------
def test_container_spelling_takes_the_tile_store_path(cuda):
    import scipy.sparse as sp
    from lmdec_b200 import PowerMethod
    from lmdec_b200.array import ChainedArray, StackedArray
    from lmdec_b200.array.operators import as_operator
    from lmdec_b200.store import GenotypeOperator
    g = structured_genotypes(300, 900, miss=0.05, seed=11)
    mean, std = oracle.get_array_moments(g)
    filled, mask = oracle.mask_imputation(g, mean)
    arr = ChainedArray((StackedArray((StackedArray((filled.astype(np.int8), sp.coo_matrix(mask))), -mean)), 1 / std))
    op = as_operator(arr)
    assert isinstance(op, GenotypeOperator) and op.has_missing and op.recognised_from == "ChainedArray"
    assert as_operator(arr) is op                                    # packed once
    ref = oracle.make_snp_array(g)
    np.testing.assert_allclose(_np(op), ref.toarray().astype(np.float64), rtol=1e-12, atol=1e-12)
    before = op.store.lib.lmdec_launch_count()
    x0 = np.random.default_rng(2).standard_normal((300, 8))
    kw = dict(k=4, buffer=4, tol=1e-8, sub_svd_start=False, lmbd=0)
    pm = PowerMethod(**kw)
    U, S, V = pm.svd(arr, x0=x0, verbose=False)
    assert isinstance(pm.array, GenotypeOperator)                    # the drivers ran on the tensor-core operator
    assert op.store.lib.lmdec_launch_count() > before
    Ur, Sr, Vr = oracle.PowerMethod(**kw).svd(ref, x0=x0)
    np.testing.assert_allclose(_np(S), Sr, rtol=1e-4)
    assert oracle.subspace_dist(_np(V).T, Vr.T, Sr) <= 1e-4
    # variants: no mask / no scaling / arbitrary centring vector; and trees that must NOT be taken over
    c = np.random.default_rng(0).standard_normal(900)
    g0 = np.nan_to_num(g)
    for tree, dense in ((StackedArray((g0, c)), g0 + c), (ChainedArray((g0, 1 / std)), g0 / std),
                        (ChainedArray((StackedArray((g0, -mean)), 1 / std)), (g0 - mean) / std)):
        o = as_operator(tree)
        assert isinstance(o, GenotypeOperator)
        np.testing.assert_allclose(_np(o), dense, rtol=1e-12, atol=1e-12)
    not_geno = StackedArray((g0 + 0.5, -mean))
    assert as_operator(not_geno) is not_geno
    bad_mask = np.array(sp.coo_matrix(mask).toarray())
    bad_mask[np.isnan(g)] *= np.random.default_rng(1).uniform(0.5, 1.5, size=int(np.isnan(g).sum()))
    not_const = ChainedArray((StackedArray((StackedArray((filled, sp.coo_matrix(bad_mask))), -mean)), 1 / std))
    assert as_operator(not_const) is not_const                       # mask not constant per column: generic algebra


# --------------------------------------------------------------------------------------------------------------------
# PowerMethod.project (reference: documented, raises NotImplementedError; its xfail test states the contract)
# --------------------------------------------------------------------------------------------------------------------
def test_project_new_samples(cuda):
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    g = structured_genotypes(400, 1200, miss=0.03, seed=8)
    new = structured_genotypes(37, 1200, miss=0.05, seed=9)
    op = make_snp_array(g, std_method="binom")
    pm = PowerMethod(k=5, buffer=7, tol=1e-8, sub_svd_start=False, lmbd=0, max_iter=100)
    U, S, V = pm.svd(op, verbose=False)
    mean, std = oracle.get_array_moments(g, std_method="binom")
    new_imputed = np.where(np.isnan(new), mean, new)                 # missing -> the TRAINING mean
    expect = ((new_imputed - mean) / std) @ _np(V).T
    got = _np(pm.project(new, onto=_np(V).T))
    assert got.shape == (37, 5)
    assert np.max(np.abs(got - expect)) <= 2e-6 * np.max(np.abs(expect))
    np.testing.assert_allclose(_np(pm.project(new, onto=V)), got, rtol=0, atol=1e-12)      # [M, P] orientation
    raw = _np(pm.project(new, onto=_np(V).T, scale_center_x=False))
    assert np.max(np.abs(raw - new_imputed @ _np(V).T)) <= 2e-6 * np.max(np.abs(raw))
    # training samples project onto their own scores: (A V^T)_k = S_k sqrt(n) U_k
    # (up to the sign of each component: U and V come from separate SVDs, here as in the reference)
    own = _np(pm.project(g, onto=_np(V).T))
    np.testing.assert_allclose(np.abs(own), np.abs(_np(U) * _np(S) * np.sqrt(400)), rtol=0, atol=2e-4 * np.abs(own).max())
    # the legacy (scale=, center=) spelling of the reference's xfail test, dense floats
    rng = np.random.default_rng(0)
    a, b = rng.random((200, 60)), rng.random((10, 60))
    mu, sd = a.mean(0), a.std(0)
    u, s, vt = np.linalg.svd((a - mu) / sd, full_matrices=False)
    pm2 = PowerMethod(k=4, scale=True, center=True, factor=None, tol=1e-12, max_iter=300, lmbd=0)
    pm2.svd(a, verbose=False)
    np.testing.assert_array_almost_equal(_np(pm2.project(b, onto=vt[:4].T)), ((b - mu) / sd) @ vt[:4].T)
    with pytest.raises(ValueError):
        pm.project(new[:, :100], onto=_np(V).T)
    with pytest.raises(ValueError):
        PowerMethod().project(new, onto=_np(V).T)


# --------------------------------------------------------------------------------------------------------------------
# PLINK fileset written by hand: the bytes of the .bed are spelled out here, not produced by the code under test
# --------------------------------------------------------------------------------------------------------------------
def test_load_plink_array_hand_written_fileset(cuda, tmp_path):
    from lmdec_b200.array.io import load_plink_array
    from lmdec_b200.decomp import make_snp_array
    # 3 variants x 5 samples.  PLINK 1 .bed: magic 6c 1b, mode 01 (variant-major); per variant ceil(5/4) = 2 bytes; the
    # first sample sits in the two LOW bits; field 00 = hom A1, 01 = missing, 10 = het, 11 = hom A2.
    #   variant 0: samples 00 10 11 01 | 00        -> byte0 = 01_11_10_00 = 0x78, byte1 = 0x00
    #   variant 1: samples 11 11 10 00 | 01        -> byte0 = 00_10_11_11 = 0x2f, byte1 = 0x01
    #   variant 2: samples 10 01 00 11 | 10        -> byte0 = 11_00_01_10 = 0xc6, byte1 = 0x02
    bed = bytes([0x6C, 0x1B, 0x01, 0x78, 0x00, 0x2F, 0x01, 0xC6, 0x02])
    prefix = tmp_path / "toy"
    (tmp_path / "toy.bed").write_bytes(bed)
    (tmp_path / "toy.bim").write_text("".join(f"1\trs{i}\t0\t{100 + i}\tA\tG\n" for i in range(3)))
    (tmp_path / "toy.fam").write_text("".join(f"f{i} s{i} 0 0 0 -9\n" for i in range(5)))
    nan = np.nan
    a0 = np.array([[0, 1, 2, nan, 0], [2, 2, 1, 0, nan], [1, nan, 0, 2, 1]])       # variants x samples, a0 count
    st = load_plink_array(str(prefix))
    assert (st.n, st.p) == (3, 5)

    def dense(s):
        d = s.to_dense().cpu().numpy().astype(float)
        d[d < 0] = nan
        return d

    np.testing.assert_array_equal(dense(st), a0)
    np.testing.assert_array_equal(dense(load_plink_array(str(prefix), transpose=True)), a0.T)
    st1 = load_plink_array(bed=str(tmp_path / "toy.bed"), bim=str(tmp_path / "toy.bim"), fam=str(tmp_path / "toy.fam"))
    np.testing.assert_array_equal(dense(st1), (2 - a0).T)                             # samples x variants, a1 count
    for kw in (dict(), dict(bed="x.bed"), dict(path_to_plink_files=str(prefix), bed="a", bim="b", fam="c")):
        with pytest.raises(ValueError):
            load_plink_array(**kw)
    (tmp_path / "bad.bed").write_bytes(bytes([0x6C, 0x1C, 0x01]) + bed[3:])
    (tmp_path / "bad.bim").write_text((tmp_path / "toy.bim").read_text())
    (tmp_path / "bad.fam").write_text((tmp_path / "toy.fam").read_text())
    with pytest.raises(ValueError, match="magic"):
        load_plink_array(str(tmp_path / "bad"))
    (tmp_path / "short.bed").write_bytes(bed[:-1])
    (tmp_path / "short.bim").write_text((tmp_path / "toy.bim").read_text())
    (tmp_path / "short.fam").write_text((tmp_path / "toy.fam").read_text())
    with pytest.raises(ValueError, match="payload"):
        load_plink_array(str(tmp_path / "short"))
    # sample-major .bed (mode byte 00) of the same data
    smaj = bytes([0x6C, 0x1B, 0x00,
                  0b00_10_11_00, 0b00_01_11_10, 0b00_00_10_11, 0b00_11_00_01, 0b00_10_01_00])
    (tmp_path / "smaj.bed").write_bytes(smaj)
    (tmp_path / "smaj.bim").write_text((tmp_path / "toy.bim").read_text())
    (tmp_path / "smaj.fam").write_text((tmp_path / "toy.fam").read_text())
    np.testing.assert_array_equal(dense(load_plink_array(str(tmp_path / "smaj"))), a0)
    # the loaded store feeds make_snp_array; a larger random fileset against the oracle's operator on the decoded matrix
    g = synth_genotypes(301, 517, 0.06, seed=12)                                      # samples x variants
    plink = {0: 0b00, 1: 0b10, 2: 0b11}
    fields = np.vectorize(lambda v: 0b01 if np.isnan(v) else plink[int(v)])(g.T).astype(np.uint8)   # variants x samples
    pad = np.zeros((517, 304), dtype=np.uint8)
    pad[:, :301] = fields
    payload = (pad[:, 0::4] | (pad[:, 1::4] << 2) | (pad[:, 2::4] << 4) | (pad[:, 3::4] << 6)).astype(np.uint8)
    (tmp_path / "big.bed").write_bytes(bytes([0x6C, 0x1B, 0x01]) + payload.tobytes())
    (tmp_path / "big.bim").write_text("".join(f"1\trs{i}\t0\t{i}\tA\tG\n" for i in range(517)))
    (tmp_path / "big.fam").write_text("".join(f"f{i} s{i} 0 0 0 -9\n" for i in range(301)))
    big = load_plink_array(str(tmp_path / "big"), transpose=True)                     # samples x variants
    np.testing.assert_array_equal(dense(big), g)
    op = make_snp_array(big, std_method="binom")
    ref = oracle.make_snp_array(g, std_method="binom")
    np.testing.assert_allclose(_np(op), ref.toarray().astype(np.float64), rtol=1e-12, atol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):                                        # raw store with missing codes
        from lmdec_b200 import PowerMethod
        PowerMethod(k=2, buffer=2).svd(big, verbose=False)


# --------------------------------------------------------------------------------------------------------------------
# persistence: the reference's tree layout, with mean / 1/std / imputation vector / code counts
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["packed", "npy"])
def test_save_load_operator_tree(cuda, tmp_path, fmt):
    import torch
    from lmdec_b200.array.io import load_array_from_disk, save_array
    from lmdec_b200.decomp import make_snp_array
    from lmdec_b200.store import GenotypeOperator
    g = synth_genotypes(300, 517, 0.06, seed=12)
    op = make_snp_array(g, std_method="binom")
    target = str(tmp_path / "op")
    save_array(op, target, dense_file_format=fmt)
    with pytest.raises(ValueError, match="already exists"):
        save_array(op, target)
    listing = sorted(os.path.relpath(os.path.join(d, f), target) for d, _, fs in os.walk(target) for f in fs)
    if fmt == "npy":      # exactly the tree the reference's save_array writes for make_snp_array's operator
        assert listing == ["chained/0_stacked/0_stacked/0_array.npy", "chained/0_stacked/0_stacked/1_sparse/coords.txt",
                           "chained/0_stacked/0_stacked/1_sparse/data.txt", "chained/0_stacked/0_stacked/1_sparse/shape.txt",
                           "chained/0_stacked/1_array.npy", "chained/1_array.npy"]
        g_leaf = np.load(os.path.join(target, "chained/0_stacked/0_stacked/0_array.npy"))
        assert g_leaf.dtype == np.int8 and np.array_equal(g_leaf, np.nan_to_num(g).astype(np.int8))
    else:
        assert "chained/0_stacked/0_packed/sample_major.npy" in listing and "chained/1_array.npy" in listing
    back = load_array_from_disk(target)
    assert isinstance(back, GenotypeOperator)
    assert torch.equal(back.store.s_store, op.store.s_store) and torch.equal(back.store.v_store, op.store.v_store)
    assert torch.equal(back.store.code_counts(), op.store.code_counts())
    for a, b in ((back.mean, op.mean), (back.inv_std, op.inv_std)):
        assert torch.equal(a, b)                                       # the SAVED moments, bit for bit
    miss_cols = op.store.code_counts()[:, 3] > 0
    assert torch.equal(back.impute_mean[miss_cols], op.impute_mean[miss_cols])
    t = np.random.default_rng(0).standard_normal((517, 6))
    np.testing.assert_array_equal(_np(back.dot(t)), _np(op.dot(t)))
    # generic trees round-trip as containers
    from lmdec_b200.array import ChainedArray, StackedArray
    a, v = np.random.default_rng(1).standard_normal((40, 30)), np.random.default_rng(2).standard_normal(30)
    tree = ChainedArray((StackedArray((a, v)), 1 + v ** 2))
    save_array(tree, str(tmp_path / "tree"))
    tree2 = load_array_from_disk(str(tmp_path / "tree"))
    np.testing.assert_allclose(_np(tree2), _np(tree), rtol=1e-14)


# --------------------------------------------------------------------------------------------------------------------
# k + buffer beyond one tensor-core N tile (PowerMethod(k=40) raised)
# --------------------------------------------------------------------------------------------------------------------
def test_wide_operands_column_blocks(cuda):
    from lmdec_b200 import PowerMethod
    from lmdec_b200.decomp import make_snp_array
    g = structured_genotypes(400, 1500, miss=0.04, seed=13)
    ref = oracle.make_snp_array(g, std_method="binom")
    op = make_snp_array(g, std_method="binom")
    rng = np.random.default_rng(0)
    for K in (43, 50, 100):
        t, y = rng.standard_normal((1500, K)), rng.standard_normal((400, K))
        x_ref, w_ref = np.asarray(ref.dot(t)), np.asarray(ref.T.dot(y))
        assert np.max(np.abs(_np(op.dot(t)) - x_ref)) <= 3e-6 * np.max(np.abs(x_ref))
        assert np.max(np.abs(_np(op.T.dot(y)) - w_ref)) <= 3e-6 * np.max(np.abs(w_ref))
    x0 = rng.standard_normal((400, 50))
    kw = dict(k=40, buffer=10, tol=1e-7, sub_svd_start=False, lmbd=0, max_iter=30)
    Ur, Sr, Vr = oracle.PowerMethod(**kw).svd(ref, x0=x0)
    pm = PowerMethod(**kw)
    U, S, V = pm.svd(op, x0=x0, verbose=False)
    np.testing.assert_allclose(_np(S), Sr, rtol=1e-4)
    assert oracle.subspace_dist(_np(U), Ur, Sr) <= 1e-4 and oracle.subspace_dist(_np(V).T, Vr.T, Sr) <= 1e-4


# --------------------------------------------------------------------------------------------------------------------
# drop-in import and the caller-owned context
# --------------------------------------------------------------------------------------------------------------------
def test_import_lmdec_shim_and_context(cuda):
    import lmdec
    import lmdec_b200
    from lmdec import PowerMethod
    from lmdec.array.scaled_legacy import ScaledCenterArray                      # noqa: F401
    from lmdec.decomp.utils import make_snp_array
    assert PowerMethod is lmdec_b200.PowerMethod and lmdec.decomp.iter_methods is lmdec_b200.decomp.iter_methods
    from lmdec_b200 import _capi
    g = structured_genotypes(260, 800, miss=0.02, seed=3)
    a, b = make_snp_array(g), make_snp_array(g)
    assert a.store.ctx.handle.value != b.store.ctx.handle.value                   # one context per store
    a.store.ctx.set(_capi.KNOB_PAIR_MODE, 0)
    a.store.ctx.set(_capi.KNOB_SM_RESERVE, 16)
    assert a.store.ctx.get(_capi.KNOB_PAIR_MODE) == 0 and b.store.ctx.get(_capi.KNOB_PAIR_MODE) == 1
    assert b.store.ctx.get(_capi.KNOB_SM_RESERVE) == 0                            # nothing process-wide
    t = np.random.default_rng(0).standard_normal((800, 5))
    a.store.ctx.timing(True)
    xa, xb = _np(a.dot(t)), _np(b.dot(t))
    ms, launches = a.store.ctx.timing_read()
    assert launches == 1 and ms > 0 and b.store.ctx.timing_read()[1] == 0
    np.testing.assert_array_equal(xa, xb)                                         # knobs change schedules, not results
    with pytest.raises(_capi.LmdecError):
        a.store.ctx.set(99, 1)
