"""ctypes binding of the C ABI declared in ``include/lmdec_b200.h``.

Loading fails loudly when the library is missing: the CUDA path is the only path.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_LIB = None


class LmdecError(RuntimeError):
    pass

_i64, _i32, _p = C.c_int64, C.c_int, C.c_void_p
SIGNATURES = {
    "lmdec_last_error": (C.c_char_p, []),
    "lmdec_abi_version": (_i32, []),
    "lmdec_source_hash": (C.c_char_p, []),
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
    "lmdec_geno_matmul": (_i32, [_p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _p, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "lmdec_finalize": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "lmdec_gram_work_doubles": (_i64, [_i64, _i32]),
    "lmdec_gram": (_i32, [_p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p]),
    "lmdec_gram_reduce": (_i32, [_p, _i64, _i64, _i32, _p, _p, _p]),
    "lmdec_qvals": (_i32, [_p, _p, _i32, _i32, _p, _p]),
    "lmdec_apply_work_doubles": (_i64, [_i32]),
    "lmdec_operator_apply": (_i32, [_p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p,
                                    _p, _p, _p, _p, _i64, _p]),
    "lmdec_operator_apply_reduce": (_i32, [_p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i64, _i32, _i32, _p, _p, _p, _p,
                                           _p, _p, _p, _p, _p, _p, _p]),
    "lmdec_apass": (_i32, [_p, _p, _p, _i64, _i64, _i64, _i32, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                           _p, _p, _p, _i64, C.c_double, _p]),
    "lmdec_chol_whiten": (_i32, [_p, _i32, C.c_double, _p, _p, _p, _p]),
    "lmdec_gram_chol": (_i32, [_p, _i64, _i64, _i32, C.c_double, _p, _p, _p, _p, _p, _p]),
    "lmdec_cholqr2_small_max_elems": (_i64, []),
    "lmdec_cholqr2_grid_max_elems": (_i64, []),
    "lmdec_cholqr2_grid_work_doubles": (_i64, [_i32]),
    "lmdec_cholqr2_grid": (_i32, [_p, _i64, _i64, _i32, C.c_double, _p, _i64, _p, _p, _p, C.c_double, _p]),
    "lmdec_cholqr2_small": (_i32, [_p, _i64, _i64, _i32, C.c_double, _p, _i64, _p, _p, C.c_double, _p]),
    "lmdec_qr_pack": (_i32, [_p, _p, _p, _p, _p, _i32, _p, _p]),
    "lmdec_ctx_create": (_p, []),
    "lmdec_ctx_destroy": (None, [_p]),
    "lmdec_ctx_set": (_i32, [_p, _i32, _i32]),
    "lmdec_ctx_get": (_i32, [_p, _i32]),
    "lmdec_ctx_timing": (None, [_p, _i32]),
    "lmdec_ctx_timing_read": (_i32, [_p, _p, _p]),
    "lmdec_sqdiff": (_i32, [_p, _i64, _p, _i64, _p, _i64, _i32, _p, _p, _p]),
    "lmdec_dense_store_bytes": (_i64, [_i64, _i64]),
    "lmdec_dense_image_bytes": (_i64, [_i64, _i32]),
    "lmdec_dense_pack": (_i32, [_p, _i32, _i64, _i64, _i64, _i32, _p, _p]),
    "lmdec_dense_unpack": (_i32, [_p, _i64, _i64, _p, _i64, _p]),
    "lmdec_dense_matmul": (_i32, [_p, _p, _i64, _i64, _i64, _i64, _p, _i32, _i64, _i32, _p, _p, _p, _i64, _i32, _p]),
}


def lib_path() -> str:
    return _build.LIB


ABI_VERSION = 3


def load():
    """Return the loaded library.  The in-tree library is rebuilt first when it is missing or was built from other
    sources than the ones on disk (content hash compiled into it, `_build.is_stale`); without nvcc a stale library is
    an error, never a silent run of old kernels behind new signatures.  `LMDEC_LIB` loads an explicitly chosen
    (A/B) library as it is."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("LMDEC_LIB")      # LMDEC_LIB: A/B a differently built library
    if not path:
        path = _build.LIB
        if _build.is_stale():
            if not _build.have_nvcc():
                raise LmdecError(f"{path} is missing or was built from different sources "
                                 f"({_build.built_hash()} != {_build.source_hash()}) and nvcc is not available")
            path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.lmdec_abi_version() != ABI_VERSION:
        raise LmdecError(f"{path}: ABI version {lib.lmdec_abi_version()}, this package binds version {ABI_VERSION}")
    _LIB = lib
    return lib


# knobs of lmdec_ctx_set (include/lmdec_b200.h, enum lmdec_knob)
(KNOB_SM_RESERVE, KNOB_PAIR_MODE, KNOB_PAIR_MIN_CHUNKS, KNOB_ISSUER_GROUPS, KNOB_PRODUCER_GROUPS,
 KNOB_PRODUCER_GROUPS_MISSING, KNOB_ACC_SETS, KNOB_B_RESIDENT, KNOB_ROUND_ROBIN) = range(9)


class Context:
    """An lmdec_ctx: the tunables of the product kernel and the CUDA-event timing of its launches.  One per caller
    (GenotypeStore holds the one its operators launch through); nothing in the library is process-wide."""

    def __init__(self):
        self.lib = load()
        self.handle = C.c_void_p(self.lib.lmdec_ctx_create())
        if os.environ.get("LMDEC_PAIR_MODE") is not None:      # A/B switches
            v = int(os.environ["LMDEC_PAIR_MODE"])
            self.set(KNOB_PAIR_MODE, v)
            if v > 1:
                self.set(KNOB_PAIR_MIN_CHUNKS, v)
        for env, knob in (("LMDEC_ISSUER_GROUPS", KNOB_ISSUER_GROUPS), ("LMDEC_ACC_SETS", KNOB_ACC_SETS),
                          ("LMDEC_PRODUCER_GROUPS_MISSING", KNOB_PRODUCER_GROUPS_MISSING),
                          ("LMDEC_PRODUCER_GROUPS", KNOB_PRODUCER_GROUPS), ("LMDEC_B_RESIDENT", KNOB_B_RESIDENT),
                          ("LMDEC_ROUND_ROBIN", KNOB_ROUND_ROBIN)):
            if os.environ.get(env):
                self.set(knob, int(os.environ[env]))

    def set(self, knob: int, value: int):
        check(self.lib.lmdec_ctx_set(self.handle, knob, int(value)), "lmdec_ctx_set")

    def get(self, knob: int) -> int:
        return self.lib.lmdec_ctx_get(self.handle, knob)

    def timing(self, on: bool):
        self.lib.lmdec_ctx_timing(self.handle, int(on))

    def timing_read(self):
        """(total ms, launches) of the product kernel since timing(True)"""
        ms, n = C.c_double(0), C.c_int64(0)
        check(self.lib.lmdec_ctx_timing_read(self.handle, C.byref(ms), C.byref(n)), "lmdec_ctx_timing_read")
        return ms.value, n.value

    def __del__(self):
        try:
            if self.handle:
                self.lib.lmdec_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


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
