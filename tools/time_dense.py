"""Times the two products of a dense float32 operator (BASELINE configs[4] shape by default) on the tcgen05 path and on
cuBLAS.  Usage: python tools/time_dense.py [n p K reps]      prints one JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200.array.operators import DenseOperator

a_ = sys.argv[1:]
n, p, K = (int(a_[0]), int(a_[1]), int(a_[2])) if len(a_) >= 3 else (1_000_000, 4000, 110)
reps = int(a_[3]) if len(a_) > 3 else 5
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
a = torch.empty((n, p), dtype=torch.float32, device=dev)
for r0 in range(0, n, 65536):
    a[r0:r0 + 65536].normal_(generator=g)
t = torch.randn((p, K), generator=g, dtype=torch.float64, device=dev)
y = torch.randn((n, K), generator=g, dtype=torch.float64, device=dev)
out = {"shape": [n, p, K], "bytes_A": 4 * n * p}


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


op = DenseOperator(a, keep_dtype=True)
assert op.path == "tcgen05"
ts = torch.cuda.Event(enable_timing=True); te = torch.cuda.Event(enable_timing=True)
ts.record(); op._native_tiles().store(False); op._native_tiles().store(True); te.record(); torch.cuda.synchronize()
out["pack_both_stores_ms"] = round(ts.elapsed_time(te), 2)
op._native_tiles().ctx.timing(True)
ms_dot, x_nat = timeit(lambda: op.dot_t(t))
k_ms, k_n = op._native_tiles().ctx.timing_read()
op._native_tiles().ctx.timing(False); op._native_tiles().ctx.timing(True)
ms_tdot, w_nat = timeit(lambda: op.tdot_t(y))
k2_ms, k2_n = op._native_tiles().ctx.timing_read()
op._native_tiles().ctx.timing(False)
out.update(dot_ms=round(ms_dot, 3), tdot_ms=round(ms_tdot, 3), dot_kernel_ms=round(k_ms / max(k_n, 1), 3),
           tdot_kernel_ms=round(k2_ms / max(k2_n, 1), 3))
out["dot_GBs"] = round(4 * n * p / (k_ms / max(k_n, 1) * 1e-3) / 1e9, 1)
out["tdot_GBs"] = round(4 * n * p / (k2_ms / max(k2_n, 1) * 1e-3) / 1e9, 1)
if os.environ.get("CUBLAS", "1") != "0":
    DenseOperator.native = False
    op2 = DenseOperator(a, keep_dtype=True)
    assert op2.path == "cublas"
    ms_c_dot, x_c = timeit(lambda: op2.dot_t(t))
    ms_c_tdot, w_c = timeit(lambda: op2.tdot_t(y))
    out.update(cublas_fp32_dot_ms=round(ms_c_dot, 3), cublas_fp32_tdot_ms=round(ms_c_tdot, 3))
    out["max_rel_diff_vs_cublas"] = [float((x_nat - x_c).abs().max() / x_c.abs().max()),
                                     float((w_nat - w_c).abs().max() / w_c.abs().max())]
print(json.dumps(out))
