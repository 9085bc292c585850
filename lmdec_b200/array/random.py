"""Row sampling and partitions: pure host index arithmetic (reference lmdec/array/random.py:6-92, 188-278)."""
from __future__ import annotations

import warnings
from typing import List, Sequence, Tuple

import numpy as np


def _get_array_shape(array_shape) -> Tuple[int, int]:
    if len(array_shape) == 2:
        return int(array_shape[0]), int(array_shape[1])
    if len(array_shape) == 1:
        return int(array_shape[0]), 1
    raise ValueError("Can only split 2-D arrays. Shape = {}".format(array_shape))


def array_split(array_shape, f: float, axis: int = 0, seed: int = 42):
    """Seeded split of the row (or column) indices into a sorted sample of size int(f*n) and its complement."""
    n, p = _get_array_shape(array_shape)
    if axis == 1:
        return array_split((p, n), f=f, axis=0, seed=seed)
    if axis != 0:
        raise ValueError("Can only split on axis = 0 or 1. axis = {}".format(axis))
    sample_size = int(f * n)
    if f > 1:
        raise ValueError("Cannot split more than unity. p = {}".format(f))
    if f < 0:
        raise ValueError("Cannot split non-positive fraction. p = {}".format(f))
    if sample_size == n:
        warnings.warn("Degenerate Split. Decrease p. p = {} ~= 1".format(f))
    elif sample_size == 0:
        warnings.warn("Degenerate Split. Decrease p. p = {} ~= 0".format(f))
    index = np.arange(0, n)
    np.random.seed(seed=seed)
    np.random.shuffle(index)
    return np.sort(index[:sample_size]), np.sort(index[sample_size:])


def cumulative_partition(partitions: Sequence[slice]) -> List[slice]:
    out = []
    first = partitions[0]
    start, step, prev_end = first.start, first.step, first.start
    for part in partitions:
        if part.step != step:
            raise ValueError("Cannot find cumulative partition of non uniform step slices.")
        if prev_end == part.start:
            out.append(slice(start, part.stop, step))
            prev_end = part.stop
    return out


def array_constant_partition(array_shape, f: float, min_size: int = 5, axis: int = 0) -> List[slice]:
    n, p = _get_array_shape(array_shape)
    if axis == 1:
        return array_constant_partition((p, n), f=f, min_size=min_size, axis=0)
    if axis != 0:
        raise ValueError("Can only partition on axis = 0 or 1. axis = {}".format(axis))
    if f <= 0:
        raise ValueError("Cannot partition with non-positive fraction.")
    if f > 0.5:
        raise ValueError("Cannot partition with fraction more than 0.5.")
    if min_size < 1:
        raise ValueError("Min size must be a positive integer.")
    num_partitions = int(1 / f)
    if num_partitions == 1:
        warnings.warn("Only 1 Partition was created. int(1/p) == 1")
    size = max(n // num_partitions, min_size)
    out, start, left = [], 0, n
    while left >= size:
        out.append(slice(start, start + size))
        start += size
        left -= size
    if left >= min_size and left > 0:
        out.append(slice(start, n))
    elif 0 < left < min_size:
        out[-1] = slice(out[-1].start, n)
    return out
