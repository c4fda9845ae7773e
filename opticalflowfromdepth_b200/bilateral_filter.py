"""`sparse_bilateral_filtering` with the signature of the reference's bilateral_filter.py:13-60, on the GPU.

On the branch the reference takes (discontinuity map given, mask None) the filter is a depth-edge-gated median;
each iteration is one ofd_bilateral_iter launch.  numpy in -> numpy out (copies inside); CUDA tensor in -> CUDA tensor
out.  `image`, sigma_r, sigma_s, HR, gsHR, edge_id and num_gs_iter have no effect on the reference's return value and
are accepted and ignored here as well; a binary `mask` takes the reference's mask path (windows 3 / 5 / 7; the reference itself never
passes one).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

__all__ = ["sparse_bilateral_filtering", "sparse_bilateral_filtering_batch"]


def sparse_bilateral_filtering(depth, image, filter_size, sigma_r=0.5, sigma_s=4.0, depth_threshold=0.04, HR=False,
                               mask=None, gsHR=True, edge_id=None, num_iter=None, num_gs_iter=None, device=None):
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
    m_u8, coef_f64 = None, False
    if mask is not None:
        # the reference's mask path with a BINARY mask; its median coefficients are float32 * mask.dtype (bilateral_filter.py:180-182)
        mk = torch.as_tensor(np.asarray(mask) if isinstance(mask, np.ndarray) else mask)
        if tuple(mk.shape) != tuple(d0.shape):
            raise ValueError("mask must have the shape of depth")
        if not bool(((mk == 0) | (mk == 1)).all()):
            raise NotImplementedError("sparse_bilateral_filtering: only binary masks (0 / 1) are implemented; a fractional mask "
                                      "turns the reference's median into a weighted one")
        coef_f64 = torch.promote_types(torch.float32, mk.dtype) == torch.float64
        m_u8 = (mk != 0).to(device=dev, dtype=torch.uint8).contiguous()
        if any(int(f) not in (3, 5, 7) for f in filter_size[:num_iter]):
            raise NotImplementedError("sparse_bilateral_filtering: the mask path supports filter sizes 3, 5 and 7")
    with torch.cuda.device(dev):
        for i in range(num_iter):
            if m_u8 is None:
                cur = ops.bilateral_iter(cur, d0, int(filter_size[i]), float(depth_threshold))
            else:
                cur = ops.bilateral_iter_masked(cur, d0, m_u8, coef_f64, int(filter_size[i]), float(depth_threshold))
    if num_iter == 0:
        cur = d0.clone()
    return cur.cpu().numpy() if is_numpy else cur


def sparse_bilateral_filtering_batch(depths, filter_size, depth_threshold=0.04, num_iter=None, normalize=False, return_packed=False):
    """sparse_bilateral_filtering for a RAGGED batch of CUDA depth maps (BASELINE config 2: mixed-resolution frames): every
    depths[i] is an [H_i, W_i] tensor of one dtype and device, filtered independently exactly as the single-image call
    would, but each iteration is ONE launch over all images (ofd_bilateral_iter_batch).  Returns a list of tensors that are
    views into one packed buffer.  normalize=True first applies utils.normalize_depth to every image (one ragged launch
    triple, ofd_normalize_depth_ragged) - the cfg2 front end `normalize_depth -> bilateral` without per-image launches.
    return_packed=True returns (packed buffer, shapes, pixel offsets) instead - the layout ops.disparity_pair_ragged takes."""
    if num_iter is None:
        raise TypeError("'NoneType' object cannot be interpreted as an integer")
    if not depths:
        return (torch.empty(0), [], []) if return_packed else []
    dev, dt = depths[0].device, depths[0].dtype
    shapes = [tuple(d.shape) for d in depths]
    if any(len(s_) != 2 for s_ in shapes) or any(d.device != dev or d.dtype != dt for d in depths):
        raise ValueError("depths must be [H,W] tensors of one dtype on one device")
    sizes = [h * w for h, w in shapes]
    offsets = [0]
    for n in sizes[:-1]:
        offsets.append(offsets[-1] + n)
    packed0 = torch.cat([d.reshape(-1) for d in depths])
    with torch.cuda.device(dev):
        if normalize:
            packed0 = ops.normalize_depth_ragged(packed0, sizes, offsets)
        cur = packed0
        for i in range(num_iter):
            cur = ops.bilateral_iter_batch(cur, packed0, shapes, offsets, int(filter_size[i]), float(depth_threshold))
    if num_iter == 0:
        cur = packed0.clone()
    if return_packed:
        return cur, shapes, offsets
    return [cur[o:o + n].view(h, w) for (h, w), o, n in zip(shapes, offsets, sizes)]
