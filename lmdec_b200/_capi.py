"""ctypes binding of the C ABI declared in ``include/lmdec_b200.h``.

Loading fails loudly when the library is missing: the CUDA path is the only path.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_LIB = None

_i64, _i32, _p = C.c_int64, C.c_int, C.c_void_p
SIGNATURES = {
    "lmdec_last_error": (C.c_char_p, []),
    "lmdec_abi_version": (_i32, []),
    "lmdec_launch_count": (C.c_longlong, []),
    "lmdec_store_bytes": (_i64, [_i64, _i64]),
    "lmdec_image_bytes": (_i64, [_i64, _i32]),
    "lmdec_pack_2bit": (_i32, [_p, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _p, _i64, _i64, _p, _p]),
    "lmdec_unpack_2bit": (_i32, [_p, _i64, _i64, _p, _i64, _p]),
    "lmdec_recode_bed": (_i32, [_p, _i64, _i64, _p, _p]),
    "lmdec_store_transpose": (_i32, [_p, _i64, _i64, _p, _p]),
    "lmdec_code_counts": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "lmdec_quantize_planes": (_i32, [_p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "lmdec_colabsmax": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
    "lmdec_geno_matmul": (_i32, [_p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _p, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "lmdec_finalize": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "lmdec_gram_work_doubles": (_i64, [_i64, _i32]),
    "lmdec_gram": (_i32, [_p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p]),
    "lmdec_qvals": (_i32, [_p, _p, _i32, _i32, _p, _p]),
    "lmdec_apply_work_doubles": (_i64, [_i32]),
    "lmdec_operator_apply": (_i32, [_p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p,
                                    _p, _p, _p, _i64, _p]),
    "lmdec_chol_whiten": (_i32, [_p, _i32, C.c_double, _p, _p, _p, _p]),
    "lmdec_gram_chol": (_i32, [_p, _i64, _i64, _i32, C.c_double, _p, _p, _p, _p, _p, _p]),
    "lmdec_cholqr2_small_max_elems": (_i64, []),
    "lmdec_cholqr2_grid_max_elems": (_i64, []),
    "lmdec_cholqr2_grid_work_doubles": (_i64, [_i32]),
    "lmdec_cholqr2_grid": (_i32, [_p, _i64, _i64, _i32, C.c_double, _p, _i64, _p, _p, _p, C.c_double, _p]),
    "lmdec_cholqr2_small": (_i32, [_p, _i64, _i64, _i32, C.c_double, _p, _i64, _p, _p, C.c_double, _p]),
    "lmdec_qr_pack": (_i32, [_p, _p, _p, _p, _p, _i32, _p, _p]),
    "lmdec_kernel_timing": (None, [_i32]),
    "lmdec_set_sm_reserve": (None, [_i32]),
    "lmdec_set_pair_mode": (None, [_i32]),
    "lmdec_set_issuer_groups": (None, [_i32]),
    "lmdec_kernel_timing_read": (_i32, [_p, _p]),
    "lmdec_sqdiff": (_i32, [_p, _i64, _p, _i64, _p, _i64, _i32, _p, _p, _p]),
}


class LmdecError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Return the loaded library (building it first when the sources are newer and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("LMDEC_LIB") or _build.LIB      # LMDEC_LIB: A/B a differently built library
    if not os.path.exists(path):
        path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("LMDEC_PAIR_MODE") is not None:      # A/B switch for the two-tiles-per-operand-chunk schedule
        lib.lmdec_set_pair_mode(int(os.environ["LMDEC_PAIR_MODE"]))
    if os.environ.get("LMDEC_ISSUER_GROUPS"):
        lib.lmdec_set_issuer_groups(int(os.environ["LMDEC_ISSUER_GROUPS"]))
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().lmdec_last_error().decode()
        raise LmdecError(f"{what or 'lmdec call'} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
