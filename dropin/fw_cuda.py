"""`import fw_cuda` (alt_cuda/fw.py:7) resolves here: the reference's own fw.py then runs on the B200 kernels."""
from opticalflowfromdepth_b200.fw_cuda import forward_warping  # noqa: F401
