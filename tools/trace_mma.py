"""Debug timeline of geno_mma_kernel (library built with -DLMDEC_TRACE): per chunk, cycles between the pipeline events
of CTA 0 (producer warp 0 and issuer 0).  Usage: LMDEC_LIB=ab/lib_trace.so python tools/trace_mma.py [dot|tdot] [miss] [n p K]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard
from lmdec_b200 import _capi
from lmdec_b200.store import GenotypeOperator

which = sys.argv[1] if len(sys.argv) > 1 else "dot"
miss = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
dev = torch.device("cuda", 0)
lib = _capi.load()
n, p, K = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (2504, 400_000, 20)
store = make_shard(torch, dev, n, p, miss, seed=1)
mean, std = store.moments("binom")
op = GenotypeOperator(store, mean, 1.0 / std)
t = torch.randn((p, K), dtype=torch.float64, device=dev)
y = torch.randn((n, K), dtype=torch.float64, device=dev)
for _ in range(3):
    (op.dot_t(t) if which == "dot" else op.tdot_t(y))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (12 * 64))()
lib.lmdec_trace_read.argtypes = [ctypes.c_void_p]
lib.lmdec_trace_read(buf)
tr = np.frombuffer(buf, dtype=np.int64).reshape(12, 64).copy()
t0 = tr[0, 0]
names = ["tile arrived", "A stage free", "decode issued", "st complete", "issuer sees full", "mma committed", "B full seen"]
print(f"{which} missing={miss}: cycles relative to the first event; one row per (tile-)chunk")
print("chunk " + " ".join(f"{nm:>17s}" for nm in names) + "   period")
for c in range(8, 40):
    row = [tr[r, c] - t0 for r in range(7)]
    print(f"{c:5d} " + " ".join(f"{v:17d}" for v in row) + f"   {tr[3, c] - tr[3, c - 1]:6d}")
d = np.diff(tr[3, 8:56])
print("median period (cycles per tile-chunk):", np.median(d))
print("median decode (A free -> decode issued):", np.median(tr[2, 8:56] - tr[1, 8:56]), " st wait:", np.median(tr[3, 8:56] - tr[2, 8:56]))
print("median producer wait for A stage (tile arrived -> A free):", np.median(tr[1, 8:56] - tr[0, 8:56]))
print("median issuer: full seen -> committed:", np.median(tr[5, 8:56] - tr[4, 8:56]), "; st complete -> issuer sees full:", np.median(tr[4, 8:56] - tr[3, 8:56]))
print("median issuer: previous commit -> B full seen:", np.median(tr[6, 9:56] - tr[5, 8:55]), "; B full seen -> A full seen:", np.median(tr[4, 8:56] - tr[6, 8:56]))
print("median issuer: commit -> B released:", np.median(tr[8, 8:55] - tr[5, 8:55]), "; B released -> next B full seen:", np.median(tr[6, 9:56] - tr[8, 8:55]))

# per work item (tile): issuer 0 done issuing (row 7), epilogue sees the accumulators full (9), tile finalized (10)
print("tile  issued(7)  acc_full_seen(9)  finalized(10)   epilogue cycles   tile period")
for i in range(1, 6):
    print(f"{i:4d} {tr[7, i] - t0:10d} {tr[9, i] - t0:12d} {tr[10, i] - t0:14d} {tr[10, i] - tr[9, i]:14d} {tr[10, i] - tr[10, i - 1]:14d}")
print("median epilogue cycles per tile:", np.median(tr[10, 1:6] - tr[9, 1:6]), " median tile period:", np.median(np.diff(tr[10, 0:6])))
print("tile loader: median cycles between copies issued:", np.median(np.diff(tr[11, 8:56])), " copy issued -> tile seen by producer 0:", np.median(tr[0, 8:56] - tr[11, 8:56]))
