"""Debug timeline of cholqr2_small_kernel (library built with -DLMDEC_TRACE): cycles between the phases of CTA 0.
Usage: LMDEC_LIB=scratch/lib_trace.so python tools/trace_cqr.py [n] [K]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200 import _capi
from lmdec_b200.device import cholqr2_small

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2504
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
lib = _capi.load()
x = torch.randn((n, K), dtype=torch.float64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    ev0.record()
    cholqr2_small(x)
    ev1.record()
torch.cuda.synchronize()
print(f"n={n} K={K}: {ev0.elapsed_time(ev1) * 1000:.1f} us (CUDA events, warm, includes 2 empty allocs)")
buf = (ctypes.c_longlong * (12 * 64))()
lib.lmdec_trace_read.argtypes = [ctypes.c_void_p]
lib.lmdec_trace_read(buf)
tr = np.frombuffer(buf, dtype=np.int64).reshape(12, 64)[7]
names = {0: "start", 1: "slice loaded", 2: "gram1 local", 3: "cluster sync", 4: "DSMEM sum", 5: "chol1", 6: "apply1",
         10: "gram2 local", 11: "cluster sync", 12: "DSMEM sum", 13: "chol2", 14: "apply2 (global)", 16: "final sync"}
prev = tr[0]
for i in sorted(names):
    print(f"{names[i]:>18s}  +{tr[i] - prev:7d} cycles   (t = {tr[i] - tr[0]})")
    prev = tr[i]
