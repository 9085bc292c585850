"""Drop-in import shim: ``import lmdec`` resolves to the B200-native implementation (``lmdec_b200``).

User code written against TNonet/lmdec keeps its imports -- ``from lmdec import PowerMethod``,
``from lmdec.decomp import SuccessiveBatchedPowerMethod``, ``from lmdec.array.scaled_legacy import ScaledCenterArray``,
``from lmdec.decomp.utils import make_snp_array``, ``from lmdec.array.io import load_plink_array`` -- and runs on the
GPU path (reference package root: lmdec/__init__.py:1).  Every ``lmdec.<module>`` the reference has and this build
implements is the SAME module object as ``lmdec_b200.<module>``.  Not provided (out of the hot path):
``lmdec.array.set_config``, ``lmdec.array.reduce``, ``lmdec.array.wrappers``, ``lmdec.examples``.
"""
import importlib
import sys

import lmdec_b200 as _impl
from lmdec_b200 import PowerMethod, SuccessiveBatchedPowerMethod  # noqa: F401

_MODULES = [
    "decomp", "decomp.iter_methods", "decomp.init_methods", "decomp.utils",
    "array", "array.chained", "array.stacked", "array.scaled_legacy", "array.matrix_ops", "array.metrics",
    "array.random", "array.transform", "array.io",
]
for _name in _MODULES:
    _mod = importlib.import_module("lmdec_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod

__all__ = ["PowerMethod", "SuccessiveBatchedPowerMethod", "decomp", "array"]
