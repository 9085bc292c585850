#!/usr/bin/env python
"""Runs the BASELINE.json configurations that are not the bench workload on ONE B200 and prints one JSON line each:
time-to-converged SVD, iterations, A^T.A.Q TFLOP/s of the steady-state step and the roofline fraction of the product
kernel.  Results are copied into BASELINE.md section 5.

    python scripts/run_configs.py [c1] [c3] [c4] [c5] [target]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_shard, peaks  # noqa: E402
from lmdec_b200 import PowerMethod, SuccessiveBatchedPowerMethod, _capi  # noqa: E402
from lmdec_b200.store import GenotypeOperator  # noqa: E402

dev = torch.device("cuda", 0)
lib = _capi.load()


def step_rate(op, K, n, p, steps=10):
    pm = PowerMethod(k=K - 10, buffer=10, tol=1e-30, sub_svd_start=False, lmbd=0, max_iter=10 ** 6)
    pm._reset_history()
    x = pm._initialization(op, x0=np.random.default_rng(1).standard_normal((n, K)))
    for _ in range(3):
        pm._solution_accuracy(x)
        x = pm._solution_step(x)
    torch.cuda.synchronize()
    ctx = op.store.ctx if hasattr(op, 'store') else None
    if ctx is not None:
        ctx.timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pm._solution_accuracy(x)
        x = pm._solution_step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tot, cnt = ctypes.c_double(0), ctypes.c_int64(0)
    if ctx is not None:
        tot.value, cnt.value = ctx.timing_read()
        ctx.timing(False)
    out = {"ms_per_step": ms, "tflops_step": 4.0 * n * p * K / (ms * 1e-3) / 1e12}
    if cnt.value:
        k_ms = tot.value / cnt.value
        ach = 2.0 * n * p * K / (k_ms * 1e-3) / 1e12
        out.update({"geno_mma_ms": k_ms, "geno_mma_tflops": ach, "roofline_frac": ach / peaks()["tensor_tflops"]})
    return out


def geno_config(name, n, p, k, miss, tol=1e-4, sbpm=False):
    t0 = time.time()
    store = make_shard(torch, dev, n, p, miss, seed=5)
    mean, std = store.moments("binom")
    op = GenotypeOperator(store, mean, 1.0 / std)
    torch.cuda.synchronize()
    build = time.time() - t0
    K = k + 10
    res = {"config": name, "shape": [n, p], "K": K, "missing": miss, "store_build_s": build}
    x0 = np.random.default_rng(42).standard_normal((n, K))
    t0 = time.time()
    pm = PowerMethod(k=k, buffer=10, tol=tol, scoring_method="q-vals", sub_svd_start=False, lmbd=0, max_iter=100)
    U, S, V = pm.svd(op, x0=x0, verbose=False)
    torch.cuda.synchronize()
    res.update({"svd_s": time.time() - t0, "iterations": pm.num_iter, "tol": tol, "S_top3": np.asarray(S)[:3].tolist()})
    res.update(step_rate(op, K, n, p))
    if sbpm:
        t0 = time.time()
        sb = SuccessiveBatchedPowerMethod(k=k, buffer=10, f=.1, lmbd=0, tol=tol, scoring_method="q-vals", max_sub_iter=100)
        Ub, Sb, Vb = sb.svd(op, x0=x0[: n // 10])
        torch.cuda.synchronize()
        res.update({"sbpm_s": time.time() - t0, "sbpm_batches": len(sb.history.iter["sub_svd"]),
                    "sbpm_S_rel_diff_vs_pm": float(np.max(np.abs(np.asarray(Sb) - np.asarray(S)) / np.asarray(S)))})
    print(json.dumps(res), flush=True)
    del op, store
    torch.cuda.empty_cache()


def dense_config():
    """configs[4]: dense float 1M x 4k ChainedArray, k=100 (K=110), sym_mat_mult_start; float32 tile stores, 3xTF32 tcgen05
    product (csrc/dense_mma.cuh)."""
    from lmdec_b200.array import ChainedArray
    n, p, k = 1_000_000, 4_000, 100
    g = torch.Generator(device=dev).manual_seed(3)
    u = torch.randn((n, 128), generator=g, device=dev, dtype=torch.float32)
    v = torch.randn((128, p), generator=g, device=dev, dtype=torch.float32)
    s = (100.0 * 0.97 ** torch.arange(128, device=dev, dtype=torch.float32))
    a = (u * s) @ v / np.sqrt(n) + torch.randn((n, p), generator=g, device=dev, dtype=torch.float32)
    del u
    arr = ChainedArray([a])
    t0 = time.time()
    pm = PowerMethod(k=k, buffer=10, tol=1e-4, scoring_method="q-vals", sub_svd_start="sym_mat_mult", lmbd=0,
                     max_iter=100, factor="n")
    U, S, V = pm.svd(arr, verbose=False)
    torch.cuda.synchronize()
    res = {"config": "c5 dense float 1Mx4k ChainedArray k=100 (K=110), sym_mat_mult_start",
           "product_path": arr.arrays[0].path,
           "svd_s": time.time() - t0, "iterations": pm.num_iter, "S_top3": np.asarray(S)[:3].tolist()}
    res.update(step_rate(arr, 110, n, p, steps=5))
    res["hbm_frac_of_step"] = (2 * 4.0 * n * p) / (res["ms_per_step"] * 1e-3) / (peaks()["hbm_gbs"] * 1e9)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c4", "c3", "c5"]
    if "c1" in which:
        geno_config("c1 10k x 50k, k=10, no missing", 10_000, 50_000, 10, 0.0)
    if "c4" in which:
        geno_config("c4 UKB-shaped per-GPU shard at 8 GPUs: 61,000 x 800,000, k=20 (K=30), no missing", 61_000, 800_000,
                    20, 0.0)
    if "target" in which:
        geno_config("target per-GPU shard at 8 GPUs: 12,500 x 500,000, k=10", 12_500, 500_000, 10, 0.0)
    if "c3" in which:
        geno_config("c3 50k x 2M, k=20 (K=30), SBPM 10 batches", 50_000, 2_000_000, 20, 0.0, sbpm=True)
    if "c5" in which:
        dense_config()
