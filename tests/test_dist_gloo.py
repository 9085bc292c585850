"""world_size = 2 on CPU (gloo), both decompositions.  The row-sharded iteration -- local Gram + K x K all-reduce (CholeskyQR2), p x K
all-reduce of A^T q -- reproduces the single-process oracle step.  The orthonormalisation is the package's own host
logic (lmdec_b200/array/linalg.py) with the Gram kernel swapped for a CPU matmul; the big products are the oracle's.
The SNP-sharded iteration -- local A^T q, one n x K all-reduce of A t, replicated orthonormalisation, V from the
row-sharded p x K matrix t r -- must give the same iterate and the same right singular subspace."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle, structured_genotypes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lmdec_b200.array.linalg import orthonormalize
        g = structured_genotypes(301, 400, miss=0.04, seed=3)
        n, K = g.shape[0], 8
        mean, std = oracle.get_array_moments(g)                      # global moments (all-reduced counts on the GPU)
        bounds = [0, 150, n]
        rows = slice(bounds[rank], bounds[rank + 1])
        a_loc = (np.where(np.isnan(g[rows]), mean, g[rows]) - mean) / std
        x = np.random.default_rng(42).standard_normal((n, K))[rows]
        cpu_gram = lambda t, y=None, want_colsum=False: t.T @ (t if y is None else y)
        for _ in range(3):
            q, r = orthonormalize(torch.from_numpy(x), group=dist.group.WORLD, gram_fn=cpu_gram)
            t = torch.from_numpy(a_loc.T @ q.numpy())
            dist.all_reduce(t)                                       # the p x K all-reduce
            x = a_loc @ t.numpy() / n
        gathered = [None] * world
        dist.all_gather_object(gathered, x)
        if rank == 0:
            out.put(np.concatenate(gathered, axis=0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_sharded_iteration_matches_oracle_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    x_dist = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = structured_genotypes(301, 400, miss=0.04, seed=3)
    pm = oracle.PowerMethod(k=4, buffer=4, sub_svd_start=False, lmbd=0)
    x = pm._initialization(oracle.make_snp_array(g), x0=np.random.default_rng(42).standard_normal((301, 8)))
    for _ in range(3):
        x = pm._solution_step(x)
    # same subspace and same singular values (column signs of q are a QR convention)
    s_d, s_o = np.linalg.svd(x_dist, compute_uv=False), np.linalg.svd(x, compute_uv=False)
    np.testing.assert_allclose(s_d, s_o, rtol=1e-10)
    qd, qo = np.linalg.qr(x_dist)[0], np.linalg.qr(x)[0]
    assert oracle.subspace_dist(qd, qo, np.ones(8), epsilon=0) < 1e-10


def _snp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lmdec_b200.array.linalg import orthonormalize
        g = structured_genotypes(301, 400, miss=0.04, seed=3)
        n, p, K = g.shape[0], g.shape[1], 8
        cols = slice(0, 170) if rank == 0 else slice(170, p)
        mean, std = oracle.get_array_moments(g[:, cols])               # per-SNP moments are local to the SNP shard
        a_loc = (np.where(np.isnan(g[:, cols]), mean, g[:, cols]) - mean) / std
        x = np.random.default_rng(42).standard_normal((n, K))         # replicated iterate
        cpu_gram = lambda t, y=None, want_colsum=False: t.T @ (t if y is None else y)
        for _ in range(3):
            q, r = orthonormalize(torch.from_numpy(x), group=None, gram_fn=cpu_gram)      # no collective
            t_loc = a_loc.T @ q.numpy()                                # local slice of A^T q
            xt = torch.from_numpy(a_loc @ t_loc / n)
            dist.all_reduce(xt)                                        # the n x K all-reduce
            x = xt.numpy()
        # right singular vectors: orthonormalise the row-sharded p x K matrix t r over the group
        w = torch.from_numpy(t_loc @ r)
        v_loc, _ = orthonormalize(w, group=dist.group.WORLD, gram_fn=cpu_gram)
        gathered = [None] * world
        dist.all_gather_object(gathered, v_loc.numpy())
        if rank == 0:
            out.put((x, np.concatenate(gathered, axis=0)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_snp_sharded_iteration_matches_oracle_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_snp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    x_dist, v_dist = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = structured_genotypes(301, 400, miss=0.04, seed=3)
    pm = oracle.PowerMethod(k=4, buffer=4, sub_svd_start=False, lmbd=0)
    a = oracle.make_snp_array(g)
    x = pm._initialization(a, x0=np.random.default_rng(42).standard_normal((301, 8)))
    for _ in range(3):
        x_prev, x = x, pm._solution_step(x)
    s_d, s_o = np.linalg.svd(x_dist, compute_uv=False), np.linalg.svd(x, compute_uv=False)
    np.testing.assert_allclose(s_d, s_o, rtol=1e-10)
    qd, qo = np.linalg.qr(x_dist)[0], np.linalg.qr(x)[0]
    assert oracle.subspace_dist(qd, qo, np.ones(8), epsilon=0) < 1e-10
    # V spans the column space of A^T x_prev
    a_dense = a.toarray() if hasattr(a, "toarray") else np.asarray(a)
    v_o = np.linalg.qr(a_dense.T @ x_prev)[0]
    np.testing.assert_allclose(v_dist.T @ v_dist, np.eye(8), atol=1e-10)
    assert oracle.subspace_dist(v_dist, v_o, np.ones(8), epsilon=0) < 1e-10
