"""lmdec_b200: B200-native implementation of TNonet/lmdec's decomposition hot path behind lmdec's own API."""
from .decomp.iter_methods import PowerMethod, SuccessiveBatchedPowerMethod  # noqa: F401  (lmdec/__init__.py:1)

__all__ = ["PowerMethod", "SuccessiveBatchedPowerMethod"]
