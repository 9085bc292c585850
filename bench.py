#!/usr/bin/env python
"""Benchmark of the decomposition hot path (BASELINE.json): one power-iteration step  x <- A (A^T orth(x)) / n.

  python bench.py --gpus N --steps K --warmup W            this build (CUDA, one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (NumPy oracle port, host cores)

Workload at every N: BASELINE.json configs[1] per GPU -- a 1000G-shaped 2,504 x 1,000,000 genotype shard with 5 % iid
missing (mean-imputed), centre + binomial scale, k = 10 (+ buffer 10 => K = 20); weak scaling.  The multi-GPU
decomposition follows the shape (`--shard auto`, stated in config.parallelism): with n << p (this workload) the SNP axis
is sharded -- every GPU holds all samples of its SNP block, A^T q stays local, the only collective per step is the
all-reduce of A t (n x K float64, 400 KB); `--shard rows` forces the sample-row decomposition (all-reduce of A^T q,
p x K float32, and of the K x K Gram matrices), which is the right cut when the sample axis is the long one.

A "step" is one full iteration of PowerMethod.svd's loop body (accuracy: CholeskyQR2 + K x K SVD + q-vals; step:
A^T q, A t) on data resident in HBM.  `value` = 4 n p K flops per step / time.  `e2e` is the SAME full step through the
public API with the iterate in HOST memory: x (pinned host) -> H2D -> orthonormalise -> A^T q -> A t -> / n -> D2H.

Key "target" (every N; `--no-target` skips it): the north-star problem, top-10 PCA of a 100,000 x 500,000 genotype matrix,
samples row-sharded over the N GPUs (strong scaling; the p x K reduction of A^T q fused into the product kernel's epilogue
through the NVSwitch multicast address where the box has one), run to convergence, with the product checked against the
oracle's float64 arithmetic on a column slab.  Key "configs4_dense" (N = 1; `--no-dense` skips it): the two products of
BASELINE configs[4] (dense float32 1M x 4k, K = 110) on the 3xTF32 tcgen05 kernel, as GB/s against the HBM roofline, with
the library float32 GEMM timed beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, N_COLS, K_TOP, BUFFER, MISS = 2504, 1_000_000, 10, 10, 0.05
METRIC = "A^T.A.Q throughput of one power-iteration step (TFLOP/s; 4npK flop per step)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        # the timed region is tens of milliseconds at full clocks (see "clocks"): the BURST cuBLAS figure is the
        # denominator; the sustained one (seconds-long loop under the power cap) is reported beside it
        return {"hbm_gbs": d["hbm_gbs"], "tensor_tflops": d["bf16_tflops"],
                "tensor_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json, burst bf16: kernel timed in a short region at full clocks)"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1590.0, "tensor_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md, burst)"}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the NumPy oracle (port of the reference's dask path) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_step_throughput(steps: int, warmup: int, budget_s: float = 25.0):
    """Times oracle PowerMethod._solution_step on a column sample of the workload (same n, K, missing rate)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lmdec_oracle as oracle

    def synth_genotypes(n, p, miss, seed, dtype):
        rng = np.random.default_rng(seed)
        maf = rng.uniform(0.05, 0.5, size=p)
        g = np.random.default_rng(seed + 1).binomial(2, maf, size=(n, p)).astype(dtype)
        g[np.random.default_rng(seed + 2).random((n, p)) < miss] = np.nan
        return g

    cores = os.cpu_count() or 1
    try:   # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core it can use
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    p_s = 40_000
    g = synth_genotypes(N_ROWS, p_s, MISS, seed=0, dtype=np.float32)
    op = oracle.make_snp_array(g, std_method="binom")
    pm = oracle.PowerMethod(k=K_TOP, buffer=BUFFER, sub_svd_start=False, lmbd=0)
    x = pm._initialization(op, x0=np.random.default_rng(42).standard_normal((N_ROWS, K_TOP + BUFFER)))
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x = pm._solution_step(x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_begin > budget_s and len(times) >= 1:
            break
    t = float(np.mean(times))
    flops = 4.0 * N_ROWS * p_s * (K_TOP + BUFFER)
    return {"value": flops / t / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
            "sample": f"oracle PowerMethod._solution_step on a {N_ROWS}x{p_s} column sample of the workload "
                      f"({len(times)} steps, {t * 1e3:.1f} ms/step, threaded BLAS, float64)",
            "ms_per_step": t * 1e3, "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_step_throughput(args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def resolve_shard(shard):
    """`auto`: shard the LONG axis.  configs[1] has n = 2,504 samples against p = 1,000,000 SNPs per GPU, so the SNP axis
    is sharded (sample-row shards would all-reduce the p x K side, 80 MB per step, for 0.3 ms of compute)."""
    if shard == "auto":
        return "rows" if N_ROWS >= N_COLS else "snps"
    return shard


def workload_config(n_gpus, shard="snps"):
    shard = resolve_shard(shard)
    rows = shard == "rows"
    return {"workload": f"configs[1]: 1000G-shaped {N_ROWS}x{N_COLS} 2-bit genotypes per GPU, 5% missing "
                        f"(mean-impute), center+binom scale, PowerMethod k={K_TOP} buffer={BUFFER} (K=20), "
                        f"3 int8 planes (22-bit fixed point)",
            "global_shape": [N_ROWS * n_gpus, N_COLS] if rows else [N_ROWS, N_COLS * n_gpus], "K": K_TOP + BUFFER,
            "parallelism": "single GPU" if n_gpus == 1 else
                           (f"samples row-sharded x{n_gpus}: all-reduce of A^T q (p x K float32) and of the Gram matrices"
                            if rows else
                            f"SNPs sharded x{n_gpus} (chosen by shape: n = {N_ROWS} << p; every GPU holds all "
                            f"{N_ROWS} samples of its {N_COLS} SNPs): one all-reduce of A t (n x K float64, 400 KB) "
                            f"per step.  The north-star's sample-row decomposition is measured on the north-star "
                            f"problem (key `target`)"),
            "l2": "inputs larger than L2 (2 x 626 MB tile stores per step vs 126 MB L2)"}


# ---------------------------------------------------------------------------------------------------------------
# this build
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU, sampled through NVML from a background thread about
    every millisecond with host timestamps; `summary(t0, t1)` keeps the samples that fall inside the timed region (and
    reports how many did).  Falls back to `nvidia-smi -lms 10` when the NVML bindings are missing."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.samples = []   # (host time, sm MHz, max MHz, [active reasons])
        self.proc = None
        self._stop = False
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._nvml, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._bits = [(pynvml.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                          (pynvml.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (pynvml.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "10"], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi -lms 10"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self._nvml, self._h
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append((time.time(), mhz, self._max, [nm for bit, nm in self._bits if mask & bit]))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                reasons = [nm for nm, v in zip(self.NAMES, f[2:]) if v.lower().startswith("active")]
                self.samples.append((time.time(), float(f[0]), float(f[1]), reasons))
            except Exception:
                pass

    def stop(self):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        used, where = (inside, "timed region") if len(inside) >= 2 else \
            ([s for s in self.samples if s[0] >= t0 - 0.5], "timed region and the loaded phases around it")
        reasons = sorted({r for s in used for r in s[3]})
        return {"sm_mhz": float(np.median([s[1] for s in used])) if used else None,
                "sm_max_mhz": used[0][2] if used else None, "reasons": reasons, "samples": len(used),
                "samples_in_timed_region": len(inside), "window": where, "source": self.source}


def make_shard(torch, device, n, p, miss, seed, n_pop=0):
    """synthetic genotypes generated on the device in row blocks (never a dense float matrix in HBM).
    n_pop > 0: population structure -- every sample belongs to one of n_pop populations whose allele frequencies are the
    common ones shifted by N(0, 0.08^2) per SNP (the same on every rank), so that n_pop - 1 principal components stand
    clear of the bulk as they do in real cohorts."""
    from lmdec_b200.store import GenotypeStore
    gen = torch.Generator(device=device).manual_seed(seed)
    g0 = torch.Generator(device=device).manual_seed(0)
    maf = torch.rand(p, generator=g0, device=device) * 0.45 + 0.05
    if n_pop:
        pops = (maf + 0.08 * torch.randn((n_pop, p), generator=g0, device=device)).clamp_(0.02, 0.98)
    store = GenotypeStore(n, p, device=device)
    for r0 in range(0, n, 256):
        nb = min(256, n - r0)
        if n_pop:
            lab = torch.randint(0, n_pop, (nb,), generator=gen, device=device)
            maf = pops[lab]
        g = (torch.rand((nb, p), generator=gen, device=device) < maf).to(torch.int8)
        g += (torch.rand((nb, p), generator=gen, device=device) < maf).to(torch.int8)
        if miss:
            g[torch.rand((nb, p), generator=gen, device=device) < miss] = -1
        store.pack_rows(g, r0)
    store.finish()
    return store


def run_target(torch, dist, device, world, rank, group, planes):
    """North-star problem (BASELINE.json): top-10 PCA of a 100,000 x 500,000 genotype matrix, samples ROW-sharded over
    the N GPUs of the box (strong scaling), PowerMethod run to convergence (q-vals, tol 1e-4).  Collectives per
    iteration: all-reduce of A^T q (p x K float32, 40 MB) and of the K x K Gram matrices (CholeskyQR2).  The converged
    factors are checked against the ORACLE's arithmetic on a column slab: (A^T U)[J] computed on the host in float64 from
    the decoded genotypes must equal V[:, J]^T S sqrt(n)."""
    from lmdec_b200 import PowerMethod
    from lmdec_b200.store import GenotypeOperator
    n, p, k, buffer, slab = 100_000, 500_000, 10, 10, 512
    K = k + buffer
    bounds = [n * r // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    t0 = time.time()
    store = make_shard(torch, device, hi - lo, p, 0.0, seed=7000 + rank, n_pop=6)
    if world > 1:
        store.group, store.n_global, store.row_offset = group, n, lo
    mean, std = store.moments("binom")
    op = GenotypeOperator(store, mean, 1.0 / std, planes=planes)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    x0 = np.random.default_rng(42).standard_normal((n, K))[lo:hi]
    kw = dict(k=k, buffer=buffer, tol=1e-4, scoring_method="q-vals", sub_svd_start=False, lmbd=0, max_iter=200)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    PowerMethod(**dict(kw, max_iter=3)).svd(op, x0=x0, verbose=False)      # warm-up (allocations, cuBLAS handles)
    barrier()
    pm = PowerMethod(**kw)
    ts = time.time()
    U, S, V = pm.svd(op, x0=x0, verbose=False)
    barrier()
    seconds = time.time() - ts
    # steady-state iteration rate (device timed, max over ranks)
    pm2 = PowerMethod(**dict(kw, tol=1e-30, max_iter=10 ** 6))
    pm2._reset_history()
    x = pm2._initialization(op, x0=x0)
    for _ in range(3):
        pm2._solution_accuracy(x)
        x = pm2._solution_step(x)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        pm2._solution_accuracy(x)
        x = pm2._solution_step(x)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_iter = float(ms.item())
    # ---- oracle check on the first `slab` SNPs: gather U (n x k) and the decoded slab (n x slab) on rank 0
    u_loc = U.t.contiguous()
    w_prod = op.tdot_t(u_loc, out_dtype=torch.float64)[:slab].cpu().numpy()     # (A^T U)[J] through the product path
    codes = store.to_dense(transpose=True, rows=slab).T.contiguous()            # [n_loc, slab] int8
    if world > 1:
        sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
        u_all = [torch.empty((sz, k), dtype=u_loc.dtype, device=device) for sz in sizes]
        c_all = [torch.empty((sz, slab), dtype=torch.int8, device=device) for sz in sizes]
        dist.all_gather(u_all, u_loc)
        dist.all_gather(c_all, codes)
        u_full, c_full = torch.cat(u_all), torch.cat(c_all)
    else:
        u_full, c_full = u_loc, codes
    out = None
    fused_reduce = any(isinstance(key, tuple) and key[0] == "symm" and val for key, val in op._buf.items())
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lmdec_oracle as oracle
        g = c_full.cpu().numpy().astype(np.float64)
        ref = oracle.make_snp_array(g, std_method="binom")
        w_ref = np.asarray(ref.T.dot(u_full.cpu().numpy()))                     # (A^T U)[J], float64 on the host
        S_np, V_np = np.asarray(S), np.asarray(V)
        w_gpu = V_np[:, :slab].T * (S_np * np.sqrt(n))
        sign = np.sign(np.sum(w_ref * w_gpu, axis=0))
        err = float(np.linalg.norm(w_ref - w_gpu * sign) / np.linalg.norm(w_ref))
        s_slab = np.linalg.norm(w_ref, axis=0) / np.linalg.norm(V_np[:, :slab], axis=1) / np.sqrt(n)
        flops_iter = 4.0 * n * p * K
        pk = peaks()
        t_roof = max(flops_iter / world / (pk["tensor_tflops"] * 1e12), (n / world) * p / 4 / (pk["hbm_gbs"] * 1e9))
        col_err = np.linalg.norm(w_ref - w_gpu * sign, axis=0) / np.linalg.norm(w_ref, axis=0)
        prod_err = float(np.max(np.abs(w_prod - w_ref)) / np.max(np.abs(w_ref)))
        out = {"problem": "top-10 PCA of a 100,000 x 500,000 2-bit genotype matrix (6 synthetic populations, no "
                          "missing codes), centre + binom scale, K = 20, q-vals tol 1e-4, x0 = default_rng(42)",
               "parallelism": "single GPU" if world == 1 else
                              f"samples row-sharded x{world} (strong scaling): reduction of A^T q (p x K float32, 40 MB) "
                              + ("fused into the product's epilogue (multimem.red through the NVSwitch multicast address)"
                                 if fused_reduce else "by NCCL all-reduce")
                              + (", K x K Gram matrices reduced the same way" if fused_reduce
                                 else ", NCCL all-reduce of the K x K Gram matrices"),
               "n_gpus": world, "iterations": int(pm.num_iter), "seconds_to_converged": seconds,
               "tflops_to_converged": flops_iter * pm.num_iter / seconds / 1e12,
               "ms_per_iteration": ms_iter, "tflops": flops_iter / (ms_iter * 1e-3) / 1e12,
               "roofline_frac_per_gpu": t_roof / (ms_iter * 1e-3), "store_build_seconds": t_build,
               "S": [float(v) for v in S_np],
               "oracle_slab_check": {"slab_snps": slab,
                                     # parity of the product itself at full size (oracle float64 vs the tensor-core path)
                                     "max_err_AtU_product_vs_oracle": prod_err,
                                     # SVD identity A^T U = V S sqrt(n): holds per component once that component has
                                     # converged (the separated population components; the bulk ones are still rotating)
                                     "rel_err_AtU_vs_V_S_sqrt_n": err,
                                     "rel_err_per_component": [float(v) for v in col_err],
                                     "S_from_oracle_slab_rel_dev": float(np.max(np.abs(s_slab - S_np) / S_np))}}
    del op, store
    torch.cuda.empty_cache()
    return out


def run_dense(torch, device, reps=5):
    """BASELINE configs[4] (not the headline workload; reported beside it at N = 1): the two products of a dense float32
    1M x 4k operator with K = 110 on the tcgen05 3xTF32 kernel (csrc/dense_mma.cuh) -- HBM-bound by its algorithmic bytes
    (4 n p per product), so the figure of merit is GB/s against the measured HBM roofline; the library float32 GEMM
    (cuBLAS through torch) is timed beside it and both are compared with each other."""
    from lmdec_b200.array.operators import DenseOperator
    n, p, K = 1_000_000, 4000, 110
    g = torch.Generator(device=device).manual_seed(3)
    a = torch.empty((n, p), dtype=torch.float32, device=device)
    for r0 in range(0, n, 65536):
        a[r0:r0 + 65536].normal_(generator=g)
    t = torch.randn((p, K), generator=g, dtype=torch.float64, device=device)
    y = torch.randn((n, K), generator=g, dtype=torch.float64, device=device)

    def timeit(fn, reps):
        for _ in range(2):
            r = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    op = DenseOperator(a, keep_dtype=True)
    if op.path != "tcgen05":
        return {"error": "dense operator did not take the tcgen05 path"}
    tiles = op._native_tiles()
    tiles.store(False), tiles.store(True)
    out = {"workload": "configs[4]: dense float32 1,000,000 x 4,000, K = 110 (k = 100 + buffer 10)",
           "kernel": "dense_mma_kernel (tcgen05 kind::tf32, 3xTF32, float32 accumulation)", "bound": "hbm"}
    res = {}
    for name, fn in (("dot", lambda: op.dot_t(t)), ("tdot", lambda: op.tdot_t(y))):
        tiles.ctx.timing(True)
        ms, r = timeit(fn, reps)
        k_ms, k_n = tiles.ctx.timing_read()
        tiles.ctx.timing(False)
        res[name] = r
        out[name + "_ms"] = ms
        out[name + "_kernel_ms"] = k_ms / max(k_n, 1)
    pk = peaks()
    k_avg = 0.5 * (out["dot_kernel_ms"] + out["tdot_kernel_ms"])
    out["achieved_GBs"] = 4.0 * n * p / (k_avg * 1e-3) / 1e9
    out["peak_GBs"] = pk["hbm_gbs"]
    out["frac"] = out["achieved_GBs"] / pk["hbm_gbs"]
    DenseOperator.native = False
    try:
        op2 = DenseOperator(a, keep_dtype=True)
        ms_d, x_c = timeit(lambda: op2.dot_t(t), 2)
        ms_t, w_c = timeit(lambda: op2.tdot_t(y), 2)
        out["cublas_fp32_dot_ms"], out["cublas_fp32_tdot_ms"] = ms_d, ms_t
        out["max_rel_diff_vs_cublas_fp32"] = max(float((res["dot"] - x_c).abs().max() / x_c.abs().max()),
                                                 float((res["tdot"] - w_c).abs().max() / w_c.abs().max()))
    finally:
        DenseOperator.native = True
    del a, op, tiles, res
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from lmdec_b200 import PowerMethod, _capi
    from lmdec_b200 import store as store_mod
    from lmdec_b200.store import GenotypeOperator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path is the only path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    lib = _capi.load()
    if args.target_only:
        target = run_target(torch, dist, device, world, rank, group, args.planes)
        if rank == 0:
            print(json.dumps({"target": target, "env": {k: v for k, v in os.environ.items() if k.startswith("LMDEC_")}}))
        if world > 1:
            dist.destroy_process_group()
        return
    K = K_TOP + BUFFER
    by_rows = resolve_shard(args.shard) == "rows" and world > 1
    n_global = N_ROWS * world if by_rows else N_ROWS
    p_global = N_COLS if by_rows else N_COLS * world

    t0 = time.time()
    store = make_shard(torch, device, N_ROWS, N_COLS, MISS, seed=1000 + rank)
    if by_rows:
        store.group, store.n_global, store.row_offset = group, n_global, rank * N_ROWS
    elif world > 1:
        store.snp_group, store.p_global, store.col_offset = group, p_global, rank * N_COLS
    mean, std = store.moments("binom")
    op = GenotypeOperator(store, mean, 1.0 / std, planes=args.planes)
    torch.cuda.synchronize()
    t_build = time.time() - t0

    pm = PowerMethod(k=K_TOP, buffer=BUFFER, tol=1e-30, scoring_method="q-vals", sub_svd_start=False, lmbd=0,
                     max_iter=10 ** 6)
    pm._reset_history()
    x0_full = np.random.default_rng(42).standard_normal((n_global, K))
    my_rows = slice(rank * N_ROWS, (rank + 1) * N_ROWS) if by_rows else slice(0, N_ROWS)
    x = pm._initialization(op, x0=x0_full[my_rows])

    def one_step(x):
        pm._solution_accuracy(x)
        return pm._solution_step(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        x = one_step(x)
    barrier()
    sampler = ClockSampler(local)
    time.sleep(0.05)
    store_mod.KERNEL_EVENTS = []
    store.ctx.timing(True)
    launches0 = lib.lmdec_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile_range:
        torch.cuda.profiler.start()      # ncu --profile-from-start off captures only the timed steps
    t_region0 = time.time()
    e0.record()
    for _ in range(args.steps):
        x = one_step(x)
    e1.record()
    barrier()
    t_region1 = time.time()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = lib.lmdec_launch_count() - launches0
    kernel_events = store_mod.KERNEL_EVENTS
    store_mod.KERNEL_EVENTS = None
    k_ms = {"dot": [], "tdot": [], "apass": []}   # whole operator application (stats + planes + product + finalize)
    for tag, a, b in kernel_events:
        k_ms[tag].append(a.elapsed_time(b))
    mma_total_ms, mma_launches = store.ctx.timing_read()
    store.ctx.timing(False)

    # ---- e2e: the SAME full step (accuracy + step of PowerMethod.svd's loop) with the iterate in pinned HOST memory:
    # H2D of x, CholeskyQR2 + K x K SVD + q-vals, A^T q, A t, / n, D2H of the next iterate -- all inside the timed region
    x_host = x.detach().cpu().pin_memory()
    out_host = torch.empty((N_ROWS, K), dtype=torch.float64).pin_memory()
    pm_e = PowerMethod(k=K_TOP, buffer=BUFFER, tol=1e-30, scoring_method="q-vals", sub_svd_start=False, lmbd=0,
                       max_iter=10 ** 6)
    pm_e._reset_history()
    pm_e._initialization(op, x0=x_host)

    def e2e_step():
        xd = x_host.to(device, non_blocking=True)
        pm_e._solution_accuracy(xd)
        out_host.copy_(pm_e._solution_step(xd), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the user holds the result on the host
        x_host.copy_(out_host)

    for _ in range(3):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    sampler.stop()
    e2e_ms_total = f0.elapsed_time(f1)

    times = torch.tensor([ms_total, e2e_ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_step = float(times[0].item()) / args.steps
    e2e_ms_step = float(times[1].item()) / args.steps
    flops_step = 4.0 * n_global * p_global * K

    # ---- time-to-converged SVD through PowerMethod.svd (informative; BASELINE metric's first half)
    svd_info = None
    if rank == 0 or world > 1:
        kw2 = dict(k=K_TOP, buffer=BUFFER, tol=1e-4, scoring_method="q-vals", sub_svd_start=False, lmbd=0)
        # first call: pays the lazy loading of every kernel svd() touches for the first time (finalization: cuBLAS
        # float64 GEMMs, the large-iterate QR) -- reported, then the same call again is the time-to-converged
        barrier()
        ts = time.time()
        PowerMethod(max_iter=3, **kw2).svd(op, x0=x0_full[my_rows], verbose=False)
        torch.cuda.synchronize()
        t_first = time.time() - ts
        pm2 = PowerMethod(max_iter=100, **kw2)
        barrier()
        ts = time.time()
        pm2.svd(op, x0=x0_full[my_rows], verbose=False)
        torch.cuda.synchronize()
        svd_info = {"seconds": time.time() - ts, "iterations": pm2.num_iter, "tol": 1e-4,
                    "ms_per_iteration": (time.time() - ts) * 1e3 / max(pm2.num_iter, 1),
                    "first_call_3_iterations_seconds": t_first, "store_build_seconds": t_build}

    # ---- the north-star problem (strong scaling over the N GPUs): see run_target
    target = None
    if args.target:
        del op, store, pm, pm_e
        torch.cuda.empty_cache()
        try:
            target = run_target(torch, dist, device, world, rank, group, args.planes)
        except Exception as e:      # never lose the headline line to the extra
            target = {"error": f"{type(e).__name__}: {e}"}

    dense = None
    if world == 1 and args.dense:
        try:
            dense = run_dense(torch, device)
        except Exception as e:      # never lose the headline line to the extra
            dense = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        pk = peaks()
        avg_kernel_ms = mma_total_ms / max(mma_launches, 1)
        flops_launch = 2.0 * N_ROWS * N_COLS * K          # per GPU, per geno_mma launch (one of the two products)
        achieved = flops_launch / (avg_kernel_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02", "geno_mma_traffic.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r01", "geno_mma_traffic.json")
        if os.path.exists(tpath) and args.planes == 3:
            with open(tpath) as f:
                traffic = json.load(f)["dram_bytes_per_launch_mean"]   # ncu --set full, dram read + write, per launch
        t_tensor = flops_launch / (pk["tensor_tflops"] * 1e12)
        t_hbm = (N_ROWS * N_COLS / 4) / (pk["hbm_gbs"] * 1e9)
        line = {
            "metric": METRIC, "value": flops_step / (ms_step * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8 x int8 -> int32 tensor cores (exact), f64 small algebra",
            "data": "synthetic", "config": workload_config(world, args.shard),
            "clocks": sampler.summary(t_region0, t_region1),
            "e2e": {"value": flops_step / (e2e_ms_step * 1e-3) / 1e12, "unit": "TFLOP/s",
                    "h2d_bytes_per_step": N_ROWS * K * 8, "d2h_bytes_per_step": N_ROWS * K * 8,
                    "ms_per_step": e2e_ms_step,
                    "api": "one full PowerMethod iteration (accuracy: CholeskyQR2 + K x K SVD + q-vals; step: "
                           "x <- A (A^T q) / n through GenotypeOperator) with the iterate x in pinned host memory: "
                           "H2D of x and D2H of the next x every step"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tensor_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tensor_tflops"], "traffic": traffic,
                         "frac_of_sustained_peak": achieved / pk["tensor_tflops_sustained"],
                         "step_frac": flops_step / world / (ms_step * 1e-3) / 1e12 / pk["tensor_tflops"],
                         "kernel": "geno_mma_kernel (2 launches per step: A^T q and A t)",
                         "avg_launch_ms": avg_kernel_ms, "launches_timed": mma_launches,
                         "kernel_share_of_step": mma_total_ms / ms_total,
                         "operator_dot_ms": float(np.mean(k_ms["dot"])) if k_ms["dot"] else None,
                         "operator_tdot_ms": float(np.mean(k_ms["tdot"])) if k_ms["tdot"] else None,
                         "operator_apass_ms": float(np.mean(k_ms["apass"])) if k_ms["apass"] else None,
                         "algorithmic_flops_per_launch": flops_launch,
                         "algorithmic_bytes_per_launch": N_ROWS * N_COLS / 4,
                         "t_tensor_ms": t_tensor * 1e3, "t_hbm_ms": t_hbm * 1e3, "peak_source": pk["source"]},
            "svd": svd_info,
        }
        if target is not None:
            line["target"] = target
        if dense is not None:
            line["configs4_dense"] = dense
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu_step_throughput(3, 1).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--planes", type=int, default=3)
    ap.add_argument("--shard", default="auto", choices=["auto", "snps", "rows"],
                    help="multi-GPU decomposition: by shape (auto: SNP shards for this n << p workload), or forced")
    ap.add_argument("--no-target", dest="target", action="store_false",
                    help="skip the north-star problem (100k x 500k, row-sharded, run to convergence)")
    ap.add_argument("--target-only", action="store_true", help="run only the north-star problem (key `target`)")
    ap.add_argument("--no-dense", dest="dense", action="store_false",
                    help="skip the configs[4] dense float products (N = 1 only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
