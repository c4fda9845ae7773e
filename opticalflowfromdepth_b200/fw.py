"""`FW` — the forward-warp module with the call signature of the reference's alt_cuda/fw.py:11-59.

FW(device).forward(obj[C,H,W], flow[2,H,W], depth[1,H,W]) -> (output[C,H,W], valid[1,H,W], collision[1,H,W]),
all float32.  The reference builds a pixel grid on the CPU, copies it to the device and runs ~12 torch kernels
before calling its extension; here target computation (p0 + flow, clamp, truncate — in the flow's dtype) is fused
into the z-test kernel (ofd_splat_flow), so one call is two kernel launches and no host->device copy.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops

__all__ = ["FW"]


class FW(nn.Module):
    def __init__(self, device=None):
        super().__init__()
        self.device = device

    def set_shape(self, obj_shape):
        print(f"{obj_shape = }")

    @torch.no_grad()
    def forward(self, obj, flow, depth, epilogue=ops.EPI_NONE, aux=None):
        dev = torch.device(self.device) if self.device is not None else obj.device
        # fw.py:31: p1 = p0(float32) + flow  -> computed in promote(float32, flow.dtype)
        fdt = torch.promote_types(torch.float32, flow.dtype)
        flow_b = flow.to(device=dev, dtype=fdt).contiguous().unsqueeze(0)
        obj_b = obj.to(device=dev, dtype=torch.float32).contiguous().unsqueeze(0)      # fw.py:40
        depth_b = depth.to(device=dev, dtype=torch.float32).contiguous().unsqueeze(0)  # fw.py:43
        aux_b = None if aux is None else aux.to(device=dev, dtype=torch.float32).contiguous().unsqueeze(0)
        with torch.cuda.device(dev):
            out, valid, collision = ops.splat_flow(obj_b, flow_b, depth_b, epilogue=epilogue, aux=aux_b)
        return out.squeeze(0), valid.squeeze(0), collision.squeeze(0)
