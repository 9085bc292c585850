"""Generates tests/golden/ex1_golden.npz from the reference's stored known-answer SVD.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py

The reference keeps the exact top-10 SVD of a deterministic 20000 x 8000 matrix (np.random.seed(10);
randn - 0.05*rand; lmdec/examples/array.py:50-62) in lmdec/examples/ex1_{Uk,Sk,Vk}.npy.  Singular values 2..10 form a
tight cluster (229.4 .. 230.7), so only the leading pair of vectors is a well-conditioned fixture; we keep all ten
singular values, the first left / right singular vectors, and the Gram matrices Uk^T Uk / Vk Vk^T as checksums.
"""
import os

import numpy as np

REF = "/root/reference/lmdec/examples"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ex1_golden.npz")

if __name__ == "__main__":
    Uk, Sk, Vk = (np.load(os.path.join(REF, f"ex1_{n}.npy")) for n in ("Uk", "Sk", "Vk"))
    assert Uk.shape == (20000, 10) and Vk.shape == (10, 8000) and Sk.shape == (10,)
    np.savez_compressed(OUT, Sk=Sk, u1=Uk[:, 0], v1=Vk[0, :], uu=Uk.T @ Uk, vv=Vk @ Vk.T,
                        u_colsum=Uk.sum(axis=0), v_rowsum=Vk.sum(axis=1))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
