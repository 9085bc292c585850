"""Where each role of geno_mma_kernel waits (library built with -DLMDEC_TRACE): cycles CTA 0's role warps spend inside
their barrier waits during ONE launch, against the launch's total cycles.  Usage:
LMDEC_LIB=ab/lib_trace.so python tools/wait_profile.py [dot|tdot] [miss] [n p K]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard
from lmdec_b200 import _capi
from lmdec_b200.store import GenotypeOperator

which = sys.argv[1] if len(sys.argv) > 1 else "tdot"
miss = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
n, p, K = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (2504, 1_000_000, 20)
dev = torch.device("cuda", 0)
lib = _capi.load()
store = make_shard(torch, dev, n, p, miss, seed=1)
mean, std = store.moments("binom")
op = GenotypeOperator(store, mean, 1.0 / std)
t = torch.randn((p, K), dtype=torch.float64, device=dev)
y = torch.randn((n, K), dtype=torch.float64, device=dev)
fn = (lambda: op.dot_t(t)) if which == "dot" else (lambda: op.tdot_t(y))
for _ in range(3):
    fn()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
lib.lmdec_wait_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.lmdec_wait_read(buf, 1)
fn()
torch.cuda.synchronize()
lib.lmdec_wait_read(buf, 1)
w = np.frombuffer(buf, dtype=np.int64).copy()
total = w[15]
names = ["epilogue warp 0: accumulators full", "issuer 0: accumulators free", "issuer 0: B stage full",
         "issuer 0: A stage full", "issuer 2: first MMA of the tile done", "producer warp 0: tile arrived",
         "producer warp 0: A stage free", "tile loader: tile slot free", "B loader: B stage free"]
print(f"{which} missing={miss} shape=({n},{p}) K={K}: kernel {total} cycles on CTA 0; share of it spent waiting:")
for i, nm in enumerate(names):
    print(f"  {nm:42s} {w[i]:10d}  {100.0 * w[i] / total:5.1f} %")
