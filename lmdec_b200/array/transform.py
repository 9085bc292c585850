"""u / s shaping for rmse (reference lmdec/array/transform.py:8-78)."""
from __future__ import annotations

import torch

from ..device import as_tensor


def acc_format_svd(u, s, array_shape, square: bool = True):
    """Returns (u [m, k] device tensor, s [k] device tensor of the diagonal)."""
    m, n = array_shape
    ut = as_tensor(u)
    if ut.ndim == 1:
        ut = ut[:, None]
    if ut.ndim != 2:
        raise Exception("u is of the wrong shape, {}".format(tuple(ut.shape)))
    m_vector, k_vector = ut.shape
    st = as_tensor(s)
    if square:
        st = st ** 2
    if st.ndim == 1:
        k_values = st.shape[0]
    elif st.ndim == 2:
        if st.shape[0] != st.shape[1]:
            raise Exception("s is the wrong shape, {}".format(tuple(st.shape)))
        k_values = st.shape[0]
        st = torch.diagonal(st)
    else:
        raise Exception("s is the wrong shape, {}".format(tuple(st.shape)))
    if m == m_vector:
        pass
    elif m == k_vector:
        ut = ut.T
        m_vector, k_vector = k_vector, m_vector
    else:
        raise Exception("u or s is of the wrong shape")
    if k_values != k_vector:
        raise Exception("u or s is of the wrong shape")
    return ut, st
