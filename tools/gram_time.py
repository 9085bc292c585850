import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lmdec_b200.device import gram, gram_chol
dev = torch.device("cuda", 0)
for n, K in ((1_000_000, 110), (1_000_000, 64), (1_000_000, 30), (200_000, 110)):
    x = torch.randn((n, K), dtype=torch.float64, device=dev)
    def t(fn, reps=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): r = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r
    a, g1 = t(lambda: gram(x))
    b, g2 = t(lambda: x.T @ x)
    c, _ = t(lambda: gram_chol(x))
    print(n, K, "lmdec_gram %.3f ms  torch x.T@x %.3f ms  gram_chol %.3f ms  rel diff %.2e" % (a, b, c, float((g1 - g2).abs().max() / g2.abs().max())))
