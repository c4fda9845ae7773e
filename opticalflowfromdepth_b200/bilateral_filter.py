"""`sparse_bilateral_filtering` with the signature of the reference's bilateral_filter.py:13-60, on the GPU.

On the branch the reference takes (discontinuity map given, mask None) the filter is a depth-edge-gated median;
each iteration is one ofd_bilateral_iter launch.  numpy in -> numpy out (copies inside); CUDA tensor in -> CUDA tensor
out.  `image`, sigma_r, sigma_s, HR, gsHR, edge_id and num_gs_iter have no effect on the reference's return value and
are accepted and ignored here as well; `mask` is not supported (the reference never passes one).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

__all__ = ["sparse_bilateral_filtering"]


def sparse_bilateral_filtering(depth, image, filter_size, sigma_r=0.5, sigma_s=4.0, depth_threshold=0.04, HR=False,
                               mask=None, gsHR=True, edge_id=None, num_iter=None, num_gs_iter=None, device=None):
    if mask is not None:
        raise NotImplementedError("sparse_bilateral_filtering: the mask path is not implemented on the B200 path")
    if num_iter is None:
        raise TypeError("'NoneType' object cannot be interpreted as an integer")  # range(None) in the reference (:33)
    is_numpy = isinstance(depth, np.ndarray)
    if is_numpy:
        if depth.dtype not in (np.float32, np.float64):
            depth = depth.astype(np.float64)
        dev = torch.device(device if device is not None else "cuda")
        d0 = torch.from_numpy(np.ascontiguousarray(depth)).to(dev)
    else:
        d0 = depth.contiguous()
        if d0.dtype not in (torch.float32, torch.float64):
            d0 = d0.double()
        dev = d0.device
    cur = d0
    with torch.cuda.device(dev):
        for i in range(num_iter):
            cur = ops.bilateral_iter(cur, d0, int(filter_size[i]), float(depth_threshold))
    if num_iter == 0:
        cur = d0.clone()
    return cur.cpu().numpy() if is_numpy else cur
