"""`from bilateral_filter import sparse_bilateral_filtering` (preprocess.py:16, utils.py:10) resolves here."""
from opticalflowfromdepth_b200.bilateral_filter import sparse_bilateral_filtering  # noqa: F401

__all__ = ["sparse_bilateral_filtering"]
