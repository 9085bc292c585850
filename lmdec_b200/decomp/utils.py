"""Moments, mean-imputation and construction of the genotype operator (reference lmdec/decomp/utils.py:12-289)."""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import torch

from ..device import DeviceArray, as_tensor
from ..store import GenotypeOperator, GenotypeStore, NotGenotypeError

PowerMethodSummary = namedtuple("PowerMethodSummary", ["time", "acc", "iter"])


def _as_store(array, device=None, **kw) -> GenotypeStore:
    if isinstance(array, GenotypeStore):
        return array
    if isinstance(array, GenotypeOperator):
        return array.store
    return GenotypeStore.from_dense(array, device=device, **kw)


def get_array_moments(array, mean: bool = True, std: bool = True, std_method: str = "binom", axis: int = 0):
    """nan-aware column mean and 'binom' / 'norm' standard deviation (utils.py:64-82), from exact integer counts.

    `array` is a genotype matrix ({0,1,2,NaN}) or a GenotypeStore.  Returns DeviceArrays (or None)."""
    if axis != 0:
        raise NotImplementedError("moments are taken over samples (axis=0)")
    store = _as_store(array)
    if std and std_method not in ("binom", "norm"):
        raise NotImplementedError(f"std_method, {std_method}, is not implemented ")
    m, s = store.moments(std_method if std else "binom")
    return (DeviceArray(m) if mean else None), (DeviceArray(s) if std else None)


def mask_imputation(array, mask_values=None, fill_value=0, mask_method: str = "mean", mask_axis: int = 0):
    """Split a matrix with NaNs into (filled matrix, mask holding the imputed value at every NaN) so that
    filled + mask == imputed matrix (utils.py:164-192).  Dense device tensors here; the decomposition path itself never
    materialises them (the tile store keeps the NaN positions as code 3)."""
    if mask_axis != 0:
        raise NotImplementedError("column-wise imputation only")
    a = as_tensor(array)
    nan = torch.isnan(a)
    if mask_values is None:
        if mask_method != "mean":
            raise NotImplementedError(f"mask_method, {mask_method}, is not implemented ")
        mask_values = torch.nanmean(a, dim=0)
    else:
        mask_values = as_tensor(mask_values)
    if not bool(nan.any().item()):
        raise ValueError("expected array to have maskable values, but got none.")
    filled = torch.where(nan, torch.full_like(a, float(fill_value)), a)
    mask = torch.where(nan, mask_values.expand_as(a), torch.zeros_like(a))
    return DeviceArray(filled), DeviceArray(mask)


def _make_dense_snp_array(array, mean, std, std_method, mask_nan):
    from ..array.chained import ChainedArray
    from ..array.stacked import StackedArray
    a = as_tensor(array)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    nan = torch.isnan(a)
    col_mean = torch.nanmean(a, dim=0)
    if std:
        if std_method == "binom":
            u = col_mean / 2
            col_std = torch.sqrt(2 * u * (1 - u))
        elif std_method == "norm":
            cnt = (~nan).sum(dim=0)
            dev = torch.where(nan, torch.zeros_like(a), a - col_mean)
            col_std = torch.sqrt((dev * dev).sum(dim=0) / cnt)
        else:
            raise NotImplementedError(f"std_method, {std_method}, is not implemented ")
    if mask_nan and bool(nan.any().item()):
        a = torch.where(nan, col_mean.expand_as(a), a)
    out = a
    if mean:
        out = StackedArray((out, -col_mean))
    if std:
        out = ChainedArray((out, 1.0 / col_std))
    if not mean and not std:
        from ..array.operators import DenseOperator
        out = DenseOperator(out)
    return out


def make_snp_array(array, mean: bool = True, std: bool = True, std_method: str = "binom", dtype="int8",
                   rechunk="auto", mask_method: str = "mean", mask_nan: bool = True, planes: int = 3,
                   device=None, group=None, n_global: Optional[int] = None, row_offset: int = 0,
                   snp_group=None, p_global: Optional[int] = None, col_offset: int = 0) -> GenotypeOperator:
    """SNP = ((array + mask) - mean) diag(1/std), lazily, over a device-resident 2-bit store (utils.py:259-289).

    `dtype` and `rechunk` are accepted for signature compatibility (the store is always 2-bit, there are no chunks).
    `planes` is the fixed-point width of the tensor-core operand (8*planes - 2 bits below each column's maximum).
    Multi-GPU (one process per GPU): `array` is this rank's block of samples (`group`, `n_global`, `row_offset`) or of
    SNPs (`snp_group`, `p_global`, `col_offset`) -- see GenotypeStore.
    """
    if mask_method != "mean":
        raise NotImplementedError(f"mask_method, {mask_method}, is not implemented ")
    if isinstance(array, GenotypeStore):
        store = array
    else:
        try:
            store = GenotypeStore.from_dense(array, device=device, group=group, n_global=n_global,
                                             row_offset=row_offset, snp_group=snp_group, p_global=p_global,
                                             col_offset=col_offset)
        except NotGenotypeError:
            # not a genotype matrix (the reference's own tests feed random floats, tests/test_iter_methods.py:12-33):
            # same operator over a dense device matrix.  Only this error is caught: a bad shape or a bad sharding
            # specification must surface, and the dense operator knows nothing about shards.
            if group is not None or snp_group is not None:
                raise
            return _make_dense_snp_array(array, mean, std, std_method, mask_nan)
    if store.has_missing and not mask_nan:
        raise ValueError("array holds NaN; the 2-bit store always mean-imputes them (mask_nan=True)")
    m, s = store.moments(std_method if std else "binom")
    return GenotypeOperator(store, m if mean else None, (1.0 / s) if std else None, planes=planes)
