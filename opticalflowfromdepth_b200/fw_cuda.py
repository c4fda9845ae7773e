"""Drop-in for the reference's torch extension module `fw_cuda` (alt_cuda/fw_cuda.cpp:15-30).

`forward_warping(obj, safe_y, safe_x, depth) -> [output, valid, collision]` with the reference's input checks and
error messages (fw_cuda.cpp:11-13,20-23); float32 or float64 tensors, as AT_DISPATCH_FLOATING_TYPES allows
(fw_cuda_kernel.cu:70).  The work is done by ofd_splat_targets in libofd_b200.so.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["forward_warping"]


def forward_warping(obj, safe_y, safe_x, depth):
    # CHECK_INPUT order and wording of fw_cuda.cpp:20-23 (ops._check raises the same RuntimeError texts)
    for name, t in (("obj", obj), ("safe_y", safe_y), ("safe_x", safe_x), ("depth", depth)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    with torch.cuda.device(obj.device):  # OptionalCUDAGuard(device_of(obj)), fw_cuda.cpp:24
        out, valid, collision = ops.splat_targets(obj, safe_y, safe_x, depth)
    return [out, valid, collision]
