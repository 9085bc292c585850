"""Error of the dense tcgen05 product against float64 as a function of the contraction length (prints one line each)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200.array.operators import DenseOperator
DenseOperator.native_min_elems = 0
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
for p in (32, 64, 256, 1024, 4096, 16384, 65536):
    n, K = 512, 16
    for kind in ("randn", "positive"):
        a = torch.randn((n, p), generator=g, device=dev, dtype=torch.float32)
        t = torch.randn((p, K), generator=g, device=dev, dtype=torch.float64)
        if kind == "positive":
            a, t = a.abs(), t.abs()
        op = DenseOperator(a, keep_dtype=True)
        ref = a.double() @ t
        out = op.dot_t(t)
        cub = (a @ t.float()).double()
        rel = ((out - ref) / ref.abs().max())
        print(f"p={p:6d} {kind:8s} max|err|/max|ref| {float(rel.abs().max()):.2e}  mean signed err/|ref| "
              f"{float(((out - ref) / ref.abs().clamp_min(1e-30)).mean()):+.2e}   cuBLAS fp32: "
              f"{float(((cub - ref) / ref.abs().max()).abs().max()):.2e}")
