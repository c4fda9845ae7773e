"""`from alt_cuda.fw import FW` (preprocess.py:17) resolves here: the fused-prologue FW."""
from opticalflowfromdepth_b200.fw import FW  # noqa: F401
