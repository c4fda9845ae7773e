"""torch-CPU restatement of the reference's per-pixel flow math — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

float32 reference for the floating-point kernels: each function follows the reference op for op (same torch ops, same
order, same dtypes), so on the CPU it is bit-identical to the reference's own Python (checked by
tests/golden/make_golden.py against /root/reference and pinned in tests/golden/*.npz).
"""
from __future__ import annotations

import torch


def get_random(random_range, random_begin, random_sign=True):
    """utils.py:96-100"""
    sign = torch.randint(0, 2, (1,))[0] * 2 - 1 if random_sign else torch.tensor(1)
    value = torch.rand(1)[0] * random_range + torch.tensor(random_begin)
    return sign * value


def normalize_depth(depth):
    """utils.py:102-116 on a copy (the reference half-mutates its argument)."""
    depth = depth.clone()
    depth[depth == 0] = 100
    depth[depth > 100] = 100
    dmin = depth.min()
    depth[depth == 100] = 0
    dmax = depth.max()
    out = (depth - dmin) * 98 / (dmax - dmin) + 1
    out[out == ((0 - dmin) * 98 / (dmax - dmin) + 1)] = 100
    return out


def fix_warped_depth(depth):
    """utils.py:123-126 on a copy."""
    depth = depth.clone()
    depth[depth == 0] = 100
    depth[depth > 99.5] = 100
    return depth


def disparity_flow(depth, sBf):
    """preprocess.py:239-254 with random_sign=False: depth[1,H,W], sBf 0-dim float32 tensor -> flow[2,H,W]."""
    disparity = sBf / depth
    return torch.cat((disparity, torch.zeros_like(disparity)), axis=0) * -1.0


def intrinsics(h, w):
    """preprocess.py:194-209"""
    K = torch.tensor([[[0.58, 0, 0.5, 0], [0, 0.58, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]]], dtype=torch.float32)
    K[:, 0, :] *= w
    K[:, 1, :] *= h
    return K, torch.linalg.inv(K)


def rot_from_axisangle(vec):
    """geometry.py:108-153"""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    Cc = 1 - ca
    x, y, z = axis[..., 0].unsqueeze(1), axis[..., 1].unsqueeze(1), axis[..., 2].unsqueeze(1)
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * Cc, y * Cc, z * Cc
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), dtype=torch.float32)
    rot[:, 0, 0] = torch.squeeze(x * xC + ca)
    rot[:, 0, 1] = torch.squeeze(xyC - zs)
    rot[:, 0, 2] = torch.squeeze(zxC + ys)
    rot[:, 1, 0] = torch.squeeze(xyC + zs)
    rot[:, 1, 1] = torch.squeeze(y * yC + ca)
    rot[:, 1, 2] = torch.squeeze(yzC - xs)
    rot[:, 2, 0] = torch.squeeze(zxC - ys)
    rot[:, 2, 1] = torch.squeeze(yzC + xs)
    rot[:, 2, 2] = torch.squeeze(z * zC + ca)
    rot[:, 3, 3] = 1
    return rot


def transformation_from_parameters(axisangle, translation):
    """geometry.py:70-105, invert=False"""
    R = rot_from_axisangle(axisangle)
    T = torch.zeros(translation.shape[0], 4, 4, dtype=torch.float32)
    for k in range(4):
        T[:, k, k] = 1
    T[:, :3, 3, None] = translation.clone().contiguous().view(-1, 3, 1)
    return torch.matmul(T, R)


def random_motion(arange=1. / 36., abase=1. / 36., trange=0.1, tbase=0.1):
    """preprocess.py:212-235 (draw order ax, ay, az, cx, cy, cz)."""
    import math

    ang = [get_random(math.pi * arange, math.pi * abase) for _ in range(3)]
    mot = [get_random(trange, tbase) for _ in range(3)]
    axisangle = torch.tensor([[ang]], dtype=torch.float32)
    translation = torch.tensor([[mot]], dtype=torch.float32)
    return transformation_from_parameters(axisangle, translation), axisangle, translation


def reproject_flow(depth, T1, eps=1e-7):
    """preprocess.py:265-298 + geometry.py:20-42,56-67: depth[1,h,w] (f32|f64), T1[1,4,4] -> flow[2,h,w] float32."""
    _, h, w = depth.shape
    depth = depth.unsqueeze(0)
    K, inv_K = intrinsics(h, w)
    grid = torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy")
    id_coords = torch.stack(grid, axis=0).type(torch.float32)
    ones = torch.ones(1, 1, h * w, dtype=torch.float32)
    pix = torch.unsqueeze(torch.stack([id_coords[0].view(-1), id_coords[1].view(-1)], 0), 0)
    pix = torch.cat([pix, ones], 1)
    cam = torch.matmul(inv_K[:, :3, :3], pix)                       # geometry.py:38
    cam = depth.view(1, 1, -1) * cam                                # :39
    cam = torch.cat([cam, ones], 1).type(torch.float32)             # :40
    P = torch.matmul(K, T1)[:, :3, :]                               # :57
    cp = torch.matmul(P, cam)                                       # :59
    pc = cp[:, :2, :] / (cp[:, 2, :].unsqueeze(1) + eps)            # :61
    pc = pc.view(1, 2, h, w).permute(0, 2, 3, 1)                    # :62-63
    pc[..., 0] /= w - 1                                             # :64
    pc[..., 1] /= h - 1                                             # :65
    pc = (pc - 0.5) * 2                                             # :66
    p1 = (pc + 1) / 2                                               # preprocess.py:284
    p1[:, :, :, 0] *= w - 1                                         # :285
    p1[:, :, :, 1] *= h - 1                                         # :286
    p0 = torch.stack(grid, axis=-1).type(torch.float32)             # :288-289
    flow = (p1 - p0).permute(0, 3, 1, 2)                            # :290-291
    return flow.squeeze(0)


def special_flow(h, w, kind):
    """preprocess.py:24-105 for a fresh SpecialFlow instance: kind 5 flip, 6 rotate, 7 shear -> (flow, back_flow)
    [2,h,w] float32, plus the 10 host parameters the product kernel takes (cx, cy, M, Mrev)."""
    grid = torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy")
    p0 = torch.stack(grid, axis=-1).type(torch.float32)
    params = None
    if kind == 7:
        s = get_random(0.15, 0.2)
        shear = torch.tensor([[1, s], [0, 1]]).type(torch.float32)
        rev = torch.tensor([[1, -s], [0, 1]]).type(torch.float32)
        p1, pp = p0 @ shear, p0 @ rev
        params = [0.0, 0.0] + [float(v) for v in shear.reshape(-1)] + [float(v) for v in rev.reshape(-1)]
    elif kind == 6:
        c0 = (get_random(w / 4, w / 2) + w / 2, get_random(h / 4, h / 2) + h / 2)
        c0 = torch.tensor(c0)
        theta = torch.deg2rad(get_random(2, 8))
        rot = torch.tensor([[torch.cos(theta), -torch.sin(theta)], [torch.sin(theta), torch.cos(theta)]]).type(torch.float32)
        rev = torch.tensor([[torch.cos(-theta), -torch.sin(-theta)], [torch.sin(-theta), torch.cos(-theta)]]).type(torch.float32)
        p1 = (p0 - c0) @ rot + c0
        pp = (p0 - c0) @ rev + c0
        params = [float(c0[0]), float(c0[1])] + [float(v) for v in rot.reshape(-1)] + [float(v) for v in rev.reshape(-1)]
    elif kind == 5:
        g = torch.meshgrid(torch.arange(w), torch.arange(h - 1, -1, -1), indexing="xy")
        p1 = torch.stack(g, axis=-1).type(torch.float32)
        pp = p1
    else:
        raise ValueError(kind)
    return (p1 - p0).permute(2, 0, 1), (pp - p0).permute(2, 0, 1), params


def inpaint_mask(valid, collision):
    """utils.py:137-149 up to the cv2.inpaint call, numpy only: valid, collision [H,W] float -> uint8 mask [H,W].
    (3x3 dilation with out-of-image taps ignored, as cv2.dilate's default border does.)"""
    import numpy as np

    H = np.asarray(valid)
    M = (1 - (H == np.asarray(collision))).astype(np.uint8)
    pad = np.pad(M, 1, "constant")
    Mp = np.zeros_like(M)
    for dj in range(3):
        for di in range(3):
            Mp = np.maximum(Mp, pad[dj:dj + M.shape[0], di:di + M.shape[1]])
    P = (Mp == M).astype(np.uint8)
    Hp = (H * P).astype(np.uint8)
    return (1 - Hp).astype(np.uint8)
