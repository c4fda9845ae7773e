"""Entry points of the reference's geometry.py (`BackprojectDepth`, `Project3D`, `transformation_from_parameters`)
backed by libofd_b200 kernels, plus the fused depth -> 6-DoF flow fast path.

Reference behaviour: geometry.py:17-42 (BackprojectDepth), :45-67 (Project3D), :70-153 (pose from axis-angle and
translation).  The per-pixel work runs in ofd_backproject / ofd_project / ofd_reproject_flow; the 4x4 pose algebra
is a handful of scalars and stays on the host side in torch, as the reference does it.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib, ops

__all__ = ["BackprojectDepth", "Project3D", "transformation_from_parameters", "rot_from_axisangle", "get_translation_matrix"]


class BackprojectDepth(nn.Module):
    """depth[b,1,h,w], inv_K[b,4,4] -> homogeneous camera points [b,4,h*w] (float32)."""

    def __init__(self, b, h, w, device):
        super().__init__()
        self.b, self.h, self.w = b, h, w
        self.device = torch.device(device)

    @torch.no_grad()
    def forward(self, depth, inv_K):
        depth = depth.to(self.device).contiguous()
        if depth.dtype not in (torch.float32, torch.float64):
            depth = depth.float()
        invk3 = inv_K[:, :3, :3].to(device=self.device, dtype=torch.float32).reshape(self.b, 9).contiguous()
        pts = torch.empty((self.b, 4, self.h * self.w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("ofd_backproject", ops._ptr(depth), ops._DT[depth.dtype], ops._ptr(invk3), self.b, self.h, self.w,
                      ops._ptr(pts), ops._stream(self.device))
        return pts


class Project3D(nn.Module):
    """points[b,4,h*w], K[b,4,4], T[b,4,4] -> (pix_coords[b,h,w,2] in [-1,1], z[b,1,h*w])."""

    def __init__(self, b, h, w, eps=1e-7):
        super().__init__()
        self.b, self.h, self.w, self.eps = b, h, w, eps

    @torch.no_grad()
    def forward(self, points, K, T):
        dev = points.device
        P = torch.matmul(K, T)[:, :3, :].to(device=dev, dtype=torch.float32).reshape(self.b, 12).contiguous()
        points = points.contiguous()
        pix = torch.empty((self.b, self.h, self.w, 2), dtype=torch.float32, device=dev)
        z = torch.empty((self.b, 1, self.h * self.w), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("ofd_project", ops._ptr(points), ops._ptr(P), C.c_float(self.eps), self.b, self.h, self.w,
                      ops._ptr(pix), ops._ptr(z), ops._stream(dev))
        return pix, z


def _skew_free_rotation(axisangle):
    """Rodrigues rotation as a [b,4,4] matrix from an axis-angle vector [b,1,3] (geometry.py:108-153)."""
    theta = torch.norm(axisangle, 2, 2, True)
    n = axisangle / (theta + 1e-7)
    c, s = torch.cos(theta), torch.sin(theta)
    one_c = 1 - c
    nx, ny, nz = (n[..., k].unsqueeze(1) for k in range(3))
    sx, sy, sz = nx * s, ny * s, nz * s
    cx, cy, cz = nx * one_c, ny * one_c, nz * one_c
    xy, yz, zx = nx * cy, ny * cz, nz * cx
    R = torch.zeros((axisangle.shape[0], 4, 4), dtype=torch.float32, device=axisangle.device)
    entries = {(0, 0): nx * cx + c, (0, 1): xy - sz, (0, 2): zx + sy,
               (1, 0): xy + sz, (1, 1): ny * cy + c, (1, 2): yz - sx,
               (2, 0): zx - sy, (2, 1): yz + sx, (2, 2): nz * cz + c}
    for (r, q), v in entries.items():
        R[:, r, q] = torch.squeeze(v)
    R[:, 3, 3] = 1
    return R


def _translation_matrix(t):
    """[b,1,3] -> [b,4,4] homogeneous translation (geometry.py:91-105)."""
    M = torch.eye(4, dtype=torch.float32, device=t.device).repeat(t.shape[0], 1, 1)
    M[:, :3, 3] = t.contiguous().view(-1, 3)
    return M


def rot_from_axisangle(vec):
    """geometry.rot_from_axisangle (geometry.py:108-153): axis-angle [b,1,3] -> rotation as a [b,4,4] matrix."""
    return _skew_free_rotation(vec)


def get_translation_matrix(translation_vector):
    """geometry.get_translation_matrix (geometry.py:91-105): [b,1,3] -> homogeneous translation [b,4,4]."""
    return _translation_matrix(translation_vector)


def transformation_from_parameters(axisangle, translation, invert=False):
    """Pose matrix M = T(t) R(axisangle), or R^T T(-t) when invert (geometry.py:70-88)."""
    R = _skew_free_rotation(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    Tm = _translation_matrix(t)
    return torch.matmul(R, Tm) if invert else torch.matmul(Tm, R)


def camera_constants(K, inv_K, T):
    """Pack the 21 per-frame scalars the fused kernel needs: inv_K[:3,:3] (9) | (K @ T)[:3,:] (12), float32 [b,21]."""
    b = K.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    return torch.cat((inv_K[:, :3, :3].reshape(b, 9), P.reshape(b, 12)), dim=1).to(torch.float32).contiguous()


@torch.no_grad()
def depth_to_flow(depth, K, inv_K, T, eps=1e-7):
    """Fused BackprojectDepth + Project3D + de-normalisation + (p1 - p0) (preprocess.py:265-298 minus the RNG):
    depth[b,1,h,w] (f32|f64, CUDA), K/inv_K/T[b,4,4] -> flow[b,2,h,w] float32, one kernel."""
    cam = camera_constants(K.cpu(), inv_K.cpu(), T.cpu()).to(depth.device)
    with torch.cuda.device(depth.device):
        return ops.reproject_flow(depth.contiguous(), cam, eps)
