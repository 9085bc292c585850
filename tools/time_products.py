"""Times the two products of one iteration on the bench workload (or another shape) for the loaded library.
Usage: [LMDEC_LIB=ab/lib_x.so] python tools/time_products.py [n p K miss reps]      prints one JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_shard
from lmdec_b200.store import GenotypeOperator

a = sys.argv[1:]
n, p, K = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (2504, 1_000_000, 20)
miss = float(a[3]) if len(a) > 3 else 0.05
reps = int(a[4]) if len(a) > 4 else 20
dev = torch.device("cuda", 0)
store = make_shard(torch, dev, n, p, miss, seed=1)
mean, std = store.moments("binom")
op = GenotypeOperator(store, mean, 1.0 / std)
g = torch.Generator(device=dev).manual_seed(3)
y = torch.linalg.qr(torch.randn((n, K), generator=g, dtype=torch.float64, device=dev))[0]
t = op.tdot_t(y)
out = {"lib": os.environ.get("LMDEC_LIB", "default"), "shape": [n, p, K, miss]}
for name, fn in (("tdot_ms", lambda: op.tdot_t(y)), ("dot_ms", lambda: op.dot_t(t)),
                 ("step_ms", lambda: op.dot_t(op.tdot_t(y)))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / reps, 4)
print(json.dumps(out))
