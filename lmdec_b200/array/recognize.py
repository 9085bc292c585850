"""Recognise the container spelling of the genotype operator and route it to the 2-bit tile store.

``make_snp_array`` of the reference returns (lmdec/decomp/utils.py:278-289)

    ChainedArray(( StackedArray(( StackedArray((G, M)), -mean )), 1 / std ))

with G the int8 dosages (NaN -> 0), M a sparse matrix holding the imputed value at every NaN, ``-mean`` a (p,) vector
and ``1/std`` a (p,) diagonal.  A user who spells this out with the containers -- or loads it from a ``save_array`` tree
(lmdec/array/io.py:129-273) -- gets the same tensor-core path as ``make_snp_array``: ``as_operator`` calls
``genotype_operator_from_containers`` and, when every piece checks out, replaces the container by a
``GenotypeOperator`` over a freshly packed store.  Anything that does not match exactly (a dense member that is not
{0,1,2}, a mask that is not constant per column, a mask entry on a non-zero dosage, more members, other shapes) is left
alone and runs through the generic container algebra.

Accepted, innermost to outermost, every level optional except the matrix itself:
    G                     dense {0,1,2}
    Stacked[G, M]         M sparse, values constant per column, G == 0 wherever M is set   (mean imputation)
    Stacked[., c]         c of shape (p,) or (1, p)                                         (centring, mean = -c)
    Chained[., d]         d of shape (p,)                                                   (scaling, 1/std = d)
"""
from __future__ import annotations

from typing import Optional

import torch

from ..device import F64


def _split_stacked(node, p):
    """Stacked[X, v] with v a (p,) / (1, p) vector -> (X, v) ; Stacked[X] -> (X, None); else None"""
    from .stacked import StackedArray, _Vector
    if not isinstance(node, StackedArray) or node.ndim != 2:
        return None
    mats = [a for a in node.arrays if not isinstance(a, _Vector)]
    vecs = [a for a in node.arrays if isinstance(a, _Vector)]
    if len(vecs) > 1 or any(tuple(v.shape) not in [(p,), (1, p)] for v in vecs):
        return None
    return mats, (vecs[0].v if vecs else None)


def genotype_operator_from_containers(arr, planes: int = 3) -> Optional[object]:
    """GenotypeOperator equivalent to the container tree `arr`, or None when the tree is not the make_snp_array
    pattern.  Only called on unsharded, single-GPU inputs."""
    from ..store import GenotypeOperator, GenotypeStore, NotGenotypeError
    from .chained import ChainedArray, _Diag
    from .operators import DenseOperator, SparseOperator
    from .stacked import StackedArray

    if len(getattr(arr, "shape", ())) != 2:
        return None
    n, p = arr.shape
    node, inv_std = arr, None
    if isinstance(node, ChainedArray):
        members = node.arrays
        if len(members) == 2 and isinstance(members[1], _Diag) and members[1].shape == (p,):
            node, inv_std = members[0], members[1].v
        elif len(members) == 1:
            node = members[0]
        else:
            return None
    mean = None
    mats = [node]
    if isinstance(node, StackedArray):
        got = _split_stacked(node, p)
        if got is None:
            return None
        mats, c = got
        if c is not None:
            mean = -c
        if len(mats) == 1 and isinstance(mats[0], StackedArray):       # Stacked[ Stacked[G, M], -mean ]
            inner = _split_stacked(mats[0], p)
            if inner is None or inner[1] is not None:
                return None
            mats = inner[0]
    dense = [m for m in mats if isinstance(m, DenseOperator)]
    sparse = [m for m in mats if isinstance(m, SparseOperator)]
    if len(dense) != 1 or len(sparse) > 1 or len(dense) + len(sparse) != len(mats):
        return None
    g, mask = dense[0], (sparse[0] if sparse else None)
    if tuple(g.shape) != (n, p) or (mask is not None and tuple(mask.shape) != (n, p)):
        return None
    dev = g.a.device
    impute = None
    crow = col = None
    if mask is not None:
        csr = mask.a
        crow, col, val = csr.crow_indices(), csr.col_indices(), csr.values().to(F64)
        impute = torch.zeros(p, dtype=F64, device=dev)
        impute[col] = val
        if not bool((impute[col] == val).all().item()):        # imputed value must be one number per column
            return None
    # cheap rejection before anything is packed: a float matrix whose first rows already hold something other than
    # {0, 1, 2, NaN} is not a genotype matrix (a 16 GB dense member would otherwise be scanned and packed for nothing)
    head = g.a[:min(n, 64)]
    if not bool(((head == 0) | (head == 1) | (head == 2) | torch.isnan(head)).all().item()):
        return None
    store = GenotypeStore(n, p, device=dev)
    rows_per_block = 2048
    for r0 in range(0, n, rows_per_block):
        r1 = min(n, r0 + rows_per_block)
        blk = g.a[r0:r1].to(F64, copy=True)
        if mask is not None:
            lo, hi = int(crow[r0].item()), int(crow[r1].item())
            if hi > lo:
                rows = torch.repeat_interleave(torch.arange(r0, r1, device=dev), crow[r0 + 1:r1 + 1] - crow[r0:r1]) - r0
                cols = col[lo:hi]
                if bool((blk[rows, cols] != 0).any().item()):    # the mask must sit on filled (zero) entries
                    return None
                blk[rows, cols] = float("nan")
        store.pack_rows(blk, r0)
    try:
        store.finish()
    except NotGenotypeError:
        return None
    op = GenotypeOperator(store, mean, inv_std, planes=planes)
    if impute is not None and op.has_missing:
        op.impute_mean = impute.contiguous()
    op.recognised_from = type(arr).__name__
    return op
