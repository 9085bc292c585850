import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lmdec_oracle as oracle  # noqa: E402  (test infrastructure only)


def synth_genotypes(n, p, miss=0.0, seed=0, dtype=np.float64):
    """maf_j ~ U(0.05, 0.5), g ~ Binomial(2, maf_j), iid missing mask."""
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.05, 0.5, size=p)
    g = np.random.default_rng(seed + 1).binomial(2, maf, size=(n, p)).astype(dtype)
    if miss:
        g[np.random.default_rng(seed + 2).random((n, p)) < miss] = np.nan
    return g


def structured_genotypes(n, p, n_pop=4, miss=0.0, seed=0):
    """genotypes with population structure so the leading singular values are separated."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.1, 0.5, size=p)
    pops = np.clip(base + 0.15 * rng.standard_normal((n_pop, p)), 0.02, 0.98)
    lab = rng.integers(0, n_pop, size=n)
    g = rng.binomial(2, pops[lab]).astype(np.float64)
    if miss:
        g[rng.random((n, p)) < miss] = np.nan
    return g
