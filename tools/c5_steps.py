"""A few PowerMethod iterations on a configs[4]-shaped dense float32 ChainedArray (for ncu launch lists / timing).
Usage: python tools/c5_steps.py [n p k steps]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200 import PowerMethod
from lmdec_b200.array import ChainedArray
a_ = sys.argv[1:]
n, p, k = (int(a_[0]), int(a_[1]), int(a_[2])) if len(a_) >= 3 else (200_000, 4000, 100)
steps = int(a_[3]) if len(a_) > 3 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
a = torch.randn((n, p), generator=g, device=dev, dtype=torch.float32)
arr = ChainedArray([a])
K = k + 10
pm = PowerMethod(k=k, buffer=10, tol=1e-30, sub_svd_start=False, lmbd=0, max_iter=10 ** 6)
pm._reset_history()
x = pm._initialization(arr, x0=np.random.default_rng(1).standard_normal((n, K)))
for _ in range(2):
    pm._solution_accuracy(x); x = pm._solution_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
t0 = time.time()
for _ in range(steps):
    pm._solution_accuracy(x); x = pm._solution_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ms per step", (time.time() - t0) * 1e3 / steps, "path", arr.arrays[0].path)
