"""Where the time of PowerMethod.svd on a configs[4]-shaped dense ChainedArray goes (two calls; cProfile of the second)."""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200 import PowerMethod
from lmdec_b200.array import ChainedArray
n, p, k = (int(sys.argv[1]), 4000, 100) if len(sys.argv) > 1 else (1_000_000, 4000, 100)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
u = torch.randn((n, 128), generator=g, device=dev, dtype=torch.float32)
v = torch.randn((128, p), generator=g, device=dev, dtype=torch.float32)
s = (100.0 * 0.97 ** torch.arange(128, device=dev, dtype=torch.float32))
a = (u * s) @ v / np.sqrt(n) + torch.randn((n, p), generator=g, device=dev, dtype=torch.float32)
del u
torch.cuda.synchronize()
kw = dict(k=k, buffer=10, tol=1e-4, scoring_method="q-vals", sub_svd_start="sym_mat_mult", lmbd=0, max_iter=100, factor="n")
for call in range(2):
    t0 = time.time()
    arr = ChainedArray([a])
    t1 = time.time()
    pm = PowerMethod(**kw)
    pr = cProfile.Profile()
    if call == 1:
        pr.enable()
    U, S, V = pm.svd(arr, verbose=False)
    torch.cuda.synchronize()
    if call == 1:
        pr.disable()
    print(f"call {call}: ChainedArray() {t1 - t0:.3f} s, svd {time.time() - t1:.3f} s, iterations {pm.num_iter}, "
          f"history times: iter sum {sum(pm.history.time['iter']):.3f} acc sum {sum(pm.history.time['acc']):.3f}")
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22)
print(st.getvalue()[:6000])
