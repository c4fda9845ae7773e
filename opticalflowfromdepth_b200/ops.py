"""Tensor-level operators over the C ABI (include/ofd_b200.h).  torch is used only for device memory and streams.

Every function takes contiguous CUDA tensors, launches on torch's current stream and returns new tensors.
Batched layouts are [B,C,H,W]; see the header for the exact semantics and the reference lines each op replaces.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import EPI_BACK, EPI_CONCAT, EPI_NONE, F32, F64  # noqa: F401  (re-exported)

_DT = {torch.float32: F32, torch.float64: F64}


def _on_device(fn):
    """Run an operator with its tensors' GPU current: the native launchers query and launch on the CURRENT device
    (occupancy, SM count, kernel launch), so a call on tensors of another GPU must switch to it first.  Also
    rejects tensor arguments that live on different devices."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            items = a if isinstance(a, (tuple, list)) else (a,)
            for t in items:
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    if dev is None:
                        dev = t.device
                    elif t.device != dev:
                        raise ValueError(f"{fn.__name__}: tensors on different devices ({dev} and {t.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check(name: str, t: torch.Tensor, dtype=None, shape=None):
    # same wording as the reference's CHECK_INPUT (alt_cuda/fw_cuda.cpp:11-13)
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")


class _KeyWorkspace:
    """Per (device, stream) cache of the packed-key plane.  Armed once; splats leave it armed."""

    def __init__(self):
        self._bufs = {}

    def get(self, device: torch.device, B: int, H: int, W: int):
        need = _lib.load().ofd_workspace_bytes(B, H, W)
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
            _lib.call("ofd_workspace_reset", _ptr(buf), C.c_size_t(buf.numel()), _stream(device))
            self._bufs[key] = buf
        return buf

    def poison(self, device: torch.device):
        """Forget cached buffers of a device (after a failed call the keys may be half-consumed)."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        for k in [k for k in self._bufs if k[0] == idx]:
            del self._bufs[k]

    def clear(self):
        self._bufs.clear()


workspace = _KeyWorkspace()


def new_counters(device) -> torch.Tensor:
    """A zeroed counter block (uint64 slots, see OFD_CNT_* in the header), stored as int64."""
    return torch.zeros(_lib.CNT_SLOTS, dtype=torch.int64, device=device)


def _run_splat(fn: str, device, *args):
    try:
        _lib.call(fn, *args)
    except Exception:
        workspace.poison(device)
        raise


@_on_device
def splat_targets(obj, safe_y, safe_x, depth, want_winner=False, counters=None):
    """fw_cuda.forward_warping (alt_cuda/fw_cuda.cpp:15-26) on [B,C,H,W] / [B,1,H,W] float32 tensors."""
    for n, t in (("obj", obj), ("safe_y", safe_y), ("safe_x", safe_x), ("depth", depth)):
        _check(n, t)
    if obj.dim() != 4:
        raise ValueError("obj must be [B,C,H,W]")
    B, Cc, H, W = obj.shape
    for n, t in (("safe_y", safe_y), ("safe_x", safe_x), ("depth", depth)):
        _check(n, t, dtype=obj.dtype, shape=(B, 1, H, W))
    if obj.dtype not in _DT:
        raise TypeError(f"obj must be float32 or float64, got {obj.dtype}")
    planes = 2 if obj.dtype == torch.float64 else 1  # float64 keeps depth and source id in separate key planes
    out = torch.empty_like(obj)
    valid = torch.empty_like(depth)
    collision = torch.empty_like(depth)
    winner = torch.empty((B, 1, H, W), dtype=torch.int32, device=obj.device) if want_winner else None
    ws = workspace.get(obj.device, planes * B, H, W)
    _run_splat("ofd_splat_targets", obj.device, _ptr(obj), _ptr(safe_y), _ptr(safe_x), _ptr(depth), _DT[obj.dtype], B, Cc, H, W,
               _ptr(out), _ptr(valid), _ptr(collision), _ptr(winner), _ptr(counters), _ptr(ws),
               C.c_size_t(ws.numel()), _stream(obj.device))
    return (out, valid, collision, winner) if want_winner else (out, valid, collision)


@_on_device
def splat_flow(obj, flow, depth, epilogue=EPI_NONE, aux=None, want_winner=False, counters=None, want_collision=True, horizontal=False,
               valid_mul=None):
    """FW.forward (alt_cuda/fw.py:19-59) batched: obj[B,C,H,W] f32, flow[B,2,H,W] f32|f64, depth[B,1,H,W] f32.
    want_collision=False skips the collision plane (returned as None): 4 B/px less for callers that only use valid.
    horizontal=True is the caller's promise that flow[:,1] is +-0 everywhere (the pipeline's disparity flows): C == 2 float32 splats
    then take the row-local shared-memory kernel (ofd_splat_flow_rows: one launch, no key plane); anything else ignores the hint.
    valid_mul[B,1,H,W]: the returned valid plane is valid * valid_mul (fused into the row-local kernel; the epilogue's masking uses the raw valid)."""
    _check("obj", obj, dtype=torch.float32)
    if obj.dim() != 4:
        raise ValueError("obj must be [B,C,H,W]")
    B, Cc, H, W = obj.shape
    _check("flow", flow, dtype=(torch.float32, torch.float64), shape=(B, 2, H, W))
    _check("depth", depth, dtype=torch.float32, shape=(B, 1, H, W))
    if aux is not None:
        _check("aux", aux, dtype=torch.float32, shape=(B, Cc, H, W))
    if valid_mul is not None:
        _check("valid_mul", valid_mul, dtype=torch.float32, shape=(B, 1, H, W))
    out = torch.empty_like(obj)
    valid = torch.empty_like(depth)
    collision = torch.empty_like(depth) if want_collision else None
    if horizontal and Cc == 2 and flow.dtype == torch.float32 and not want_winner and counters is None and W <= 2048 and H <= 65535 and B <= 65535:
        _lib.call("ofd_splat_flow_rows", _ptr(obj), _ptr(flow), _ptr(depth), B, Cc, H, W, _ptr(out), _ptr(valid), _ptr(collision),
                  int(epilogue), _ptr(aux), _ptr(valid_mul), _stream(obj.device))
        return out, valid, collision
    winner = torch.empty((B, 1, H, W), dtype=torch.int32, device=obj.device) if want_winner else None
    ws = workspace.get(obj.device, B, H, W)
    _run_splat("ofd_splat_flow", obj.device, _ptr(obj), _ptr(flow), _DT[flow.dtype], _ptr(depth), B, Cc, H, W,
               _ptr(out), _ptr(valid), _ptr(collision), _ptr(winner), int(epilogue), _ptr(aux), _ptr(counters),
               _ptr(ws), C.c_size_t(ws.numel()), _stream(obj.device))
    if valid_mul is not None:
        valid = valid * valid_mul
    return (out, valid, collision, winner) if want_winner else (out, valid, collision)


@_on_device
def disparity_flow(depth, sBf):
    """Convert.depth_to_disparity + disparity_to_flow (preprocess.py:239-254): depth[B,1,H,W] -> flow[B,2,H,W]."""
    _check("depth", depth, dtype=(torch.float32, torch.float64))
    B, _, H, W = depth.shape
    _check("sBf", sBf, dtype=torch.float32, shape=(B,))
    flow = torch.empty((B, 2, H, W), dtype=depth.dtype, device=depth.device)
    _lib.call("ofd_disparity_flow", _ptr(depth), _DT[depth.dtype], _ptr(sBf), B, H, W, _ptr(flow), _stream(depth.device))
    return flow


@_on_device
def disparity_pair(img0, depth0, sBf, want_flow=True, want_collision=True, counters=None, out=None):
    """One fused virtual-stereo flow pair (preprocess.py:355-366 minus inpaint).

    Returns (img1[B,3,H,W], depth1[B,1,H,W], back_flow[B,2,H,W], flow[B,2,H,W]|None, valid, collision|None).
    `out` may hold those six tensors preallocated (None entries are skipped outputs) to keep the allocator out of a
    timed loop."""
    _check("img0", img0, dtype=torch.float32)
    B, c3, H, W = img0.shape
    if c3 != 3:
        raise ValueError("img0 must be [B,3,H,W]")
    _check("depth0", depth0, dtype=(torch.float32, torch.float64), shape=(B, 1, H, W))
    _check("sBf", sBf, dtype=torch.float32, shape=(B,))
    dev = img0.device
    if out is None:
        f32 = dict(dtype=torch.float32, device=dev)
        out = (torch.empty((B, 3, H, W), **f32), torch.empty((B, 1, H, W), **f32), torch.empty((B, 2, H, W), **f32),
               torch.empty((B, 2, H, W), **f32) if want_flow else None, torch.empty((B, 1, H, W), **f32),
               torch.empty((B, 1, H, W), **f32) if want_collision else None)
    else:
        for n, t, c in zip(("img1", "depth1", "back_flow", "flow", "valid", "collision"), out, (3, 1, 2, 2, 1, 1)):
            if t is not None:
                _check(n, t, dtype=torch.float32, shape=(B, c, H, W))
    img1, depth1, back, flow, valid, coll = out
    _lib.call("ofd_disparity_pair", _ptr(img0), _ptr(depth0), _DT[depth0.dtype], _ptr(sBf), B, H, W, _ptr(img1),
              _ptr(depth1), _ptr(back), _ptr(flow), _ptr(valid), _ptr(coll), _ptr(counters), _stream(dev))
    return img1, depth1, back, flow, valid, coll


@_on_device
def disparity_pair_ragged(img0, depth0, sBf, shapes, offsets, want_flow=True, want_collision=True, counters=None, out=None):
    """disparity_pair over a ragged batch in one launch (ofd_disparity_pair_ragged): img0 / depth0 are 1-D packed CUDA buffers,
    image i is shapes[i] = (H_i, W_i) and starts at PIXEL offset offsets[i] - a C-channel tensor holds it densely as
    [C,H_i,W_i] at element C * offsets[i] (img0: 3 * total pixels, depth0: total pixels, float32 or float64); sBf[n] float32.

    Returns packed (img1[3P], depth1[P], back_flow[2P], flow[2P]|None, valid[P], collision[P]|None), P = depth0.numel();
    `ragged_views(packed, C, shapes, offsets)` turns one of them into per-image [C,H_i,W_i] views."""
    _check("depth0", depth0, dtype=(torch.float32, torch.float64))
    P = depth0.numel()
    _check("img0", img0, dtype=torch.float32, shape=(3 * P,))
    n = len(shapes)
    _check("sBf", sBf, dtype=torch.float32, shape=(n,))
    if len(offsets) != n or any(off < 0 or off + h * w > P for (h, w), off in zip(shapes, offsets)):
        raise ValueError("an image extends past the end of the packed buffer")
    dev = depth0.device
    if out is None:
        f32 = dict(dtype=torch.float32, device=dev)
        out = (torch.empty(3 * P, **f32), torch.empty(P, **f32), torch.empty(2 * P, **f32),
               torch.empty(2 * P, **f32) if want_flow else None, torch.empty(P, **f32), torch.empty(P, **f32) if want_collision else None)
    else:
        for nm, t, c in zip(("img1", "depth1", "back_flow", "flow", "valid", "collision"), out, (3, 1, 2, 2, 1, 1)):
            if t is not None:
                _check(nm, t, dtype=torch.float32, shape=(c * P,))
    img1, depth1, back, flow, valid, coll = out
    Hs = (C.c_int * n)(*[int(h) for h, _ in shapes])
    Ws = (C.c_int * n)(*[int(w) for _, w in shapes])
    offs = (C.c_size_t * n)(*[int(o) for o in offsets])
    _lib.call("ofd_disparity_pair_ragged", _ptr(img0), _ptr(depth0), _DT[depth0.dtype], _ptr(sBf), n, Hs, Ws, offs, _ptr(img1),
              _ptr(depth1), _ptr(back), _ptr(flow), _ptr(valid), _ptr(coll), _ptr(counters), _stream(dev))
    return img1, depth1, back, flow, valid, coll


def ragged_views(packed, channels: int, shapes, offsets):
    """Per-image [C,H_i,W_i] views into a packed ragged buffer (layout of disparity_pair_ragged)."""
    return [packed[channels * o:channels * (o + h * w)].view(channels, h, w) for (h, w), o in zip(shapes, offsets)]


@_on_device
def reproject_flow(depth, cam, eps=1e-7):
    """Convert.depth_to_random_flow (preprocess.py:265-298) fused: depth[B,1,H,W] f32|f64, cam[B,21] f32 -> flow[B,2,H,W]."""
    _check("depth", depth, dtype=(torch.float32, torch.float64))
    B, _, H, W = depth.shape
    _check("cam", cam, dtype=torch.float32, shape=(B, 21))
    flow = torch.empty((B, 2, H, W), dtype=torch.float32, device=depth.device)
    _lib.call("ofd_reproject_flow", _ptr(depth), _DT[depth.dtype], _ptr(cam), C.c_float(eps), B, H, W, _ptr(flow),
              _stream(depth.device))
    return flow


@_on_device
def frame_splat(img, depth, flow, valid_in=None, want_collision=True, want_raw_valid=False, counters=None):
    """Image+flow splat of the frame pipeline (preprocess.py:372-382): returns
    (img_out, depth_out, back_flow, valid', collision|None, raw_valid|None)."""
    _check("img", img, dtype=torch.float32)
    B, c3, H, W = img.shape
    if c3 != 3:
        raise ValueError("img must be [B,3,H,W]")
    _check("depth", depth, dtype=torch.float32, shape=(B, 1, H, W))
    _check("flow", flow, dtype=(torch.float32, torch.float64), shape=(B, 2, H, W))
    if valid_in is not None:
        _check("valid_in", valid_in, dtype=torch.float32, shape=(B, 1, H, W))
    dev = img.device
    f32 = dict(dtype=torch.float32, device=dev)
    img_o = torch.empty((B, 3, H, W), **f32)
    dep_o = torch.empty((B, 1, H, W), **f32)
    back = torch.empty((B, 2, H, W), **f32)
    valid = torch.empty((B, 1, H, W), **f32)
    coll = torch.empty((B, 1, H, W), **f32) if want_collision else None
    raw = torch.empty((B, 1, H, W), **f32) if want_raw_valid else None
    ws = workspace.get(dev, B, H, W)
    if flow.dtype == torch.float64:
        # float64 warp flow (dataset path): targets in float64, payload = the float32 rounding (fw.py:31 vs :45)
        payload = flow.float()
        _run_splat("ofd_frame_splat_f64", dev, _ptr(img), _ptr(depth), _ptr(flow), _ptr(payload), _ptr(valid_in), B, H, W,
                   _ptr(img_o), _ptr(dep_o), _ptr(back), _ptr(valid), _ptr(coll), _ptr(raw), _ptr(counters), _ptr(ws),
                   C.c_size_t(ws.numel()), _stream(dev))
        return img_o, dep_o, back, valid, coll, raw
    _run_splat("ofd_frame_splat", dev, _ptr(img), _ptr(depth), _ptr(flow), _ptr(valid_in), B, H, W, _ptr(img_o),
               _ptr(dep_o), _ptr(back), _ptr(valid), _ptr(coll), _ptr(raw), _ptr(counters), _ptr(ws),
               C.c_size_t(ws.numel()), _stream(dev))
    return img_o, dep_o, back, valid, coll, raw


def concat_frame_splat_applies(flowBC, W: int, H: int, B: int) -> bool:
    """Whether ofd_concat_frame_splat takes these shapes (float32, W % 4 == 0, W <= 2048)."""
    return flowBC.dtype == torch.float32 and W % 4 == 0 and W <= 2048 and H <= 65535 and B <= 65535


@_on_device
def concat_frame_splat(flowBC, warp_flow, depthB, flowAB, img, depth_src, valid_mul=None, want_collision=True, counters=None):
    """ConcatFlow along a horizontal warp flow fused with the frame splat along its result (preprocess.py:400-411, 414-424; pairs 0->2'
    and 1->3' of a group): returns (flowAC, flowAC_valid, img_out, depth_out, back_flow, valid', collision|None) - what
    splat_flow(flowBC, warp_flow, depthB, EPI_CONCAT, aux=flowAB, horizontal=True, valid_mul=valid_mul) followed by
    frame_splat(img, depth_src, flowAC, flowAC_valid) returns, bit for bit, in two launches instead of three."""
    _check("flowBC", flowBC, dtype=torch.float32)
    B, c2, H, W = flowBC.shape
    if c2 != 2:
        raise ValueError("flowBC must be [B,2,H,W]")
    for n, t, c in (("warp_flow", warp_flow, 2), ("depthB", depthB, 1), ("flowAB", flowAB, 2), ("img", img, 3), ("depth_src", depth_src, 1)):
        _check(n, t, dtype=torch.float32, shape=(B, c, H, W))
    if valid_mul is not None:
        _check("valid_mul", valid_mul, dtype=torch.float32, shape=(B, 1, H, W))
    dev = flowBC.device
    f32 = dict(dtype=torch.float32, device=dev)
    flowAC = torch.empty((B, 2, H, W), **f32)
    flowAC_valid = torch.empty((B, 1, H, W), **f32)
    img_o = torch.empty((B, 3, H, W), **f32)
    dep_o = torch.empty((B, 1, H, W), **f32)
    back = torch.empty((B, 2, H, W), **f32)
    valid = torch.empty((B, 1, H, W), **f32)
    coll = torch.empty((B, 1, H, W), **f32) if want_collision else None
    ws = workspace.get(dev, B, H, W)
    _run_splat("ofd_concat_frame_splat", dev, _ptr(flowBC), _ptr(warp_flow), _ptr(depthB), _ptr(flowAB), _ptr(valid_mul), _ptr(img),
               _ptr(depth_src), B, H, W, _ptr(flowAC), _ptr(flowAC_valid), _ptr(img_o), _ptr(dep_o), _ptr(back), _ptr(valid), _ptr(coll),
               _ptr(counters), _ptr(ws), C.c_size_t(ws.numel()), _stream(dev))
    return flowAC, flowAC_valid, img_o, dep_o, back, valid, coll


@_on_device
def reproject_pair(img, depth, cam, valid_in=None, eps=1e-7, want_collision=True, want_raw_valid=False, counters=None):
    """Fused 6-DoF flow pair (preprocess.py:372-382): the flow is computed inside the z-test and written once.
    Returns (img_out, depth_out, back_flow, flow, valid', collision|None, raw_valid|None)."""
    _check("img", img, dtype=torch.float32)
    B, c3, H, W = img.shape
    if c3 != 3:
        raise ValueError("img must be [B,3,H,W]")
    _check("depth", depth, dtype=torch.float32, shape=(B, 1, H, W))
    _check("cam", cam, dtype=torch.float32, shape=(B, 21))
    if valid_in is not None:
        _check("valid_in", valid_in, dtype=torch.float32, shape=(B, 1, H, W))
    dev = img.device
    f32 = dict(dtype=torch.float32, device=dev)
    img_o = torch.empty((B, 3, H, W), **f32)
    dep_o = torch.empty((B, 1, H, W), **f32)
    back = torch.empty((B, 2, H, W), **f32)
    flow = torch.empty((B, 2, H, W), **f32)
    valid = torch.empty((B, 1, H, W), **f32)
    coll = torch.empty((B, 1, H, W), **f32) if want_collision else None
    raw = torch.empty((B, 1, H, W), **f32) if want_raw_valid else None
    ws = workspace.get(dev, B, H, W)
    _run_splat("ofd_reproject_pair", dev, _ptr(img), _ptr(depth), _ptr(cam), C.c_float(eps), _ptr(valid_in), B, H, W,
               _ptr(img_o), _ptr(dep_o), _ptr(back), _ptr(flow), _ptr(valid), _ptr(coll), _ptr(raw), _ptr(counters),
               _ptr(ws), C.c_size_t(ws.numel()), _stream(dev))
    return img_o, dep_o, back, flow, valid, coll, raw


@_on_device
def normalize_depth(depth):
    """utils.normalize_depth (utils.py:102-116), out of place, per frame of depth[B,1,H,W] (f32|f64)."""
    _check("depth", depth, dtype=(torch.float32, torch.float64))
    B, _, H, W = depth.shape
    out = torch.empty_like(depth)
    scratch = torch.empty(2 * max(B, 1), dtype=torch.int64, device=depth.device)
    _lib.call("ofd_normalize_depth", _ptr(depth), _DT[depth.dtype], B, H, W, _ptr(out), _ptr(scratch), _stream(depth.device))
    return out


@_on_device
def normalize_depth_ragged(packed, counts, offsets):
    """utils.normalize_depth per image of a ragged batch: `packed` is a 1-D CUDA buffer (f32|f64) holding image i as
    counts[i] elements at offsets[i]; returns the normalised packed buffer (same layout)."""
    _check("packed", packed, dtype=(torch.float32, torch.float64))
    n = len(counts)
    if any(o + c > packed.numel() for c, o in zip(counts, offsets)):
        raise ValueError("an image extends past the end of the packed buffer")
    cnt = (C.c_size_t * n)(*[int(c) for c in counts])
    off = (C.c_size_t * n)(*[int(o) for o in offsets])
    out = torch.empty_like(packed)
    scratch = torch.empty(2 * max(n, 1), dtype=torch.int64, device=packed.device)
    _lib.call("ofd_normalize_depth_ragged", _ptr(packed), _DT[packed.dtype], n, cnt, off, _ptr(out), _ptr(scratch), _stream(packed.device))
    return out


@_on_device
def depth_from_png(raw, kind: str, dtype=torch.float64):
    """The depth loaders' arithmetic on the device: `raw` is the decoded PNG payload as a uint8 / uint16 (stored as
    int16 bit pattern is not accepted: pass torch.uint16) CUDA tensor of any shape; kind "reldepth" = utils.get_depth with
    smooth_closer (utils.py:47-59,118-121), "disparity" = utils.get_disparity + Convert.disparity_to_depth (utils.py:61-72,
    preprocess.py:257-262).  Returns a tensor of `raw`'s shape in float64 (the reference's dtype) or float32."""
    if not raw.is_cuda or not raw.is_contiguous():
        raise RuntimeError("raw must be a contiguous CUDA tensor")
    bits = {torch.uint8: 8, torch.uint16: 16}.get(raw.dtype)
    if bits is None:
        raise TypeError(f"raw must be uint8 or uint16, got {raw.dtype}")
    code = {"reldepth": _lib.SRC_RELDEPTH, "disparity": _lib.SRC_DISPARITY}[kind]
    out = torch.empty(raw.shape, dtype=dtype, device=raw.device)
    _lib.call("ofd_depth_from_png", _ptr(raw), bits, code, C.c_size_t(raw.numel()), _ptr(out), _DT[dtype], _stream(raw.device))
    return out


@_on_device
def fix_warped_depth_(depth):
    """utils.fix_warped_depth (utils.py:123-126), in place."""
    _check("depth", depth, dtype=torch.float32)
    _lib.call("ofd_fix_warped_depth", _ptr(depth), C.c_size_t(depth.numel()), _stream(depth.device))
    return depth


@_on_device
def inpaint_mask(valid, collision):
    """The mask utils.inpaint hands to cv2.inpaint (utils.py:137-149): uint8 [B,1,H,W], 1 = pixel to fill."""
    _check("valid", valid, dtype=torch.float32)
    _check("collision", collision, dtype=torch.float32, shape=valid.shape)
    B, _, H, W = valid.shape
    mask = torch.empty((B, 1, H, W), dtype=torch.uint8, device=valid.device)
    _lib.call("ofd_inpaint_mask", _ptr(valid), _ptr(collision), B, H, W, _ptr(mask), _stream(valid.device))
    return mask


@_on_device
def resize_bilinear_aa(t, size):
    """torchvision's T.Resize(size) of a float tensor (antialiased bilinear, dataloader.py:31-32,57-58) on the device:
    t[..., H, W] float32|float64 CUDA -> [..., H_out, W_out] (ofd_resize_bilinear_aa)."""
    _check("t", t, dtype=(torch.float32, torch.float64))
    if t.dim() < 2:
        raise ValueError("t must be [..., H, W]")
    H, W = t.shape[-2:]
    Ho, Wo = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    B = t.numel() // (H * W) if H * W else 0
    out = torch.empty(t.shape[:-2] + (Ho, Wo), dtype=t.dtype, device=t.device)
    tmp = torch.empty((B, H, Wo), dtype=t.dtype, device=t.device) if (Ho != H and Wo != W) else None
    _lib.call("ofd_resize_bilinear_aa", _ptr(t), _DT[t.dtype], B, H, W, Ho, Wo, _ptr(out), _ptr(tmp), _stream(t.device))
    return out


class JpegDecoder:
    """nvJPEG on the device (ofd_jpeg_*): `decode(jpeg_bytes)` -> float32 [3,H,W] CUDA tensor in B, G, R channel order - what
    utils.get_img (utils.py:17-25: cv2.imread(path, -1) -> float32 CHW) returns, decoded on the GPU from the compressed bytes."""

    def __init__(self, device=0):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._h = C.c_void_p(0)
        _lib.call("ofd_jpeg_decoder_create", int(self.device.index or 0), C.byref(self._h))

    def info(self, data: bytes):
        h, w, c = C.c_int(0), C.c_int(0), C.c_int(0)
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        _lib.call("ofd_jpeg_info", self._h, buf, C.c_size_t(len(data)), C.byref(h), C.byref(w), C.byref(c))
        return h.value, w.value, c.value

    def decode(self, data: bytes) -> torch.Tensor:
        buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
        h, w, c = C.c_int(0), C.c_int(0), C.c_int(0)
        _lib.call("ofd_jpeg_info", self._h, buf, C.c_size_t(len(data)), C.byref(h), C.byref(w), C.byref(c))
        out = torch.empty((3, h.value, w.value), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("ofd_jpeg_decode", self._h, buf, C.c_size_t(len(data)), _ptr(out), h.value, w.value, _stream(self.device))
        return out

    def close(self):
        if self._h:
            _lib.load().ofd_jpeg_decoder_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_inpaint_ws = {}


@_on_device
def inpaint_telea(img, mask, radius: int = 3, want_stats: bool = False):
    """cv2.inpaint(img_u8, mask, radius, cv2.INPAINT_TELEA) of utils.inpaint (utils.py:149) on the device, batched:
    img[B,3,H,W] float32 CUDA (uint8-valued), mask[B,1,H,W] uint8 (1 = fill) -> float32 [B,3,H,W] (uint8-valued).
    Layer-ordered fast marching (ofd_inpaint_telea); want_stats also returns (layers, filled pixels) and synchronises."""
    _check("img", img, dtype=torch.float32)
    if img.dim() != 4 or img.shape[1] != 3:
        raise ValueError("img must be [B,3,H,W]")
    B, _, H, W = img.shape
    _check("mask", mask, dtype=torch.uint8, shape=(B, 1, H, W))
    out = torch.empty_like(img)
    need = _lib.load().ofd_inpaint_workspace_bytes(B, H, W)
    key = (img.device.index, torch.cuda.current_stream(img.device).cuda_stream)
    ws = _inpaint_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=img.device)
        _inpaint_ws[key] = ws
    stats = (C.c_uint32 * 2)() if want_stats else None
    _lib.call("ofd_inpaint_telea", _ptr(img), _ptr(mask), B, H, W, int(radius), _ptr(out), _ptr(ws), C.c_size_t(ws.numel()),
              stats, _stream(img.device))
    return (out, (int(stats[0]), int(stats[1]))) if want_stats else out


def special_flow(kind: int, params, H: int, W: int, device):
    """SpecialFlow (preprocess.py:24-105): returns (flow[2,H,W], back_flow[2,H,W]); params = 10 host floats or None."""
    flow = torch.empty((2, H, W), dtype=torch.float32, device=device)
    back = torch.empty((2, H, W), dtype=torch.float32, device=device)
    arr = None
    if params is not None:
        arr = (C.c_float * 10)(*[float(v) for v in params])
    with torch.cuda.device(flow.device):
        _lib.call("ofd_special_flow", int(kind), arr, H, W, _ptr(flow), _ptr(back), _stream(flow.device))
    return flow, back


def _host_params(kinds, params, B):
    kinds_arr = (C.c_int * B)(*[int(k) for k in kinds])
    flat = (C.c_float * (10 * B))()
    for b in range(B):
        if params[b] is not None:
            for k in range(10):
                flat[10 * b + k] = float(params[b][k])
    return kinds_arr, flat


def special_flow_batch(kinds, params, H: int, W: int, device):
    """One SpecialFlow per sample in one launch: kinds[B] in {5,6,7}, params[B] = 10 host floats or None (kind 5)."""
    B = len(kinds)
    flow = torch.empty((B, 2, H, W), dtype=torch.float32, device=device)
    back = torch.empty((B, 2, H, W), dtype=torch.float32, device=device)
    kinds_arr, flat = _host_params(kinds, params, B)
    with torch.cuda.device(flow.device):
        _lib.call("ofd_special_flow_batch", kinds_arr, flat, B, H, W, _ptr(flow), _ptr(back), _stream(flow.device))
    return flow, back


@_on_device
def augment_pairs(img0, depth0, img1, depth1, flow01, back_flow01, kinds, params, want_collision=True, counters=None):
    """ofd_augment_pairs: the geometric branch of augment_flow (preprocess.py:116-147 minus inpaint) for B pairs, ONE call.
    Returns a dict of the ten result tensors plus the masks of the two image warps."""
    _check("img0", img0, dtype=torch.float32)
    B, c3, H, W = img0.shape
    if c3 != 3:
        raise ValueError("img0 must be [B,3,H,W]")
    _check("img1", img1, dtype=torch.float32, shape=(B, 3, H, W))
    _check("depth0", depth0, dtype=torch.float32, shape=(B, 1, H, W))
    _check("depth1", depth1, dtype=torch.float32, shape=(B, 1, H, W))
    _check("flow01", flow01, dtype=torch.float32, shape=(B, 2, H, W))
    _check("back_flow01", back_flow01, dtype=torch.float32, shape=(B, 2, H, W))
    if len(kinds) != B or len(params) != B:
        raise ValueError("kinds and params must have one entry per sample")
    dev = img0.device
    f32 = dict(dtype=torch.float32, device=dev)
    names = (("special_flow", 2), ("back_special_flow", 2), ("aug_img0", 3), ("aug_depth0", 1), ("aug0_flow", 2),
             ("back_aug0_flow", 2), ("aug_img1", 3), ("aug_depth1", 1), ("aug1_flow", 2), ("back_aug1_flow", 2),
             ("valid_img0", 1), ("collision_img0", 1), ("valid_img1", 1), ("collision_img1", 1), ("scratch_valid", 1))
    r = {n: torch.empty((B, c, H, W), **f32) for n, c in names if want_collision or not n.startswith("collision")}
    kinds_arr, flat = _host_params(kinds, params, B)
    ws = workspace.get(dev, B, H, W)
    _run_splat("ofd_augment_pairs", dev, _ptr(img0), _ptr(depth0), _ptr(img1), _ptr(depth1), _ptr(flow01), _ptr(back_flow01),
               kinds_arr, flat, B, H, W, *[_ptr(r.get(n)) for n, _ in names], _ptr(counters), _ptr(ws),
               C.c_size_t(ws.numel()), _stream(dev))
    r.pop("scratch_valid")
    return r


@_on_device
def bilateral_iter(depth_in, depth_orig, window: int, threshold: float):
    """One iteration of sparse_bilateral_filtering (bilateral_filter.py:33-58) on [H,W] f32|f64 CUDA tensors."""
    _check("depth_in", depth_in, dtype=(torch.float32, torch.float64))
    _check("depth_orig", depth_orig, dtype=depth_in.dtype, shape=depth_in.shape)
    if depth_in.dim() != 2:
        raise ValueError("depth must be [H,W]")
    H, W = depth_in.shape
    out = torch.empty_like(depth_in)
    _lib.call("ofd_bilateral_iter", _ptr(depth_in), _ptr(depth_orig), _DT[depth_in.dtype], H, W, int(window),
              C.c_double(threshold), _ptr(out), _stream(depth_in.device))
    return out


@_on_device
def bilateral_iter_masked(depth_in, depth_orig, mask_u8, coef_f64: bool, window: int, threshold: float):
    """One iteration with the reference's binary mask (ofd_bilateral_iter_masked): mask_u8 [H,W] uint8 CUDA, 0 = masked."""
    _check("depth_in", depth_in, dtype=(torch.float32, torch.float64))
    _check("depth_orig", depth_orig, dtype=depth_in.dtype, shape=depth_in.shape)
    _check("mask", mask_u8, dtype=torch.uint8, shape=depth_in.shape)
    if depth_in.dim() != 2:
        raise ValueError("depth must be [H,W]")
    H, W = depth_in.shape
    out = torch.empty_like(depth_in)
    _lib.call("ofd_bilateral_iter_masked", _ptr(depth_in), _ptr(depth_orig), _ptr(mask_u8), int(bool(coef_f64)), _DT[depth_in.dtype],
              H, W, int(window), C.c_double(threshold), _ptr(out), _stream(depth_in.device))
    return out


@_on_device
def bilateral_iter_batch(packed_in, packed_orig, shapes, offsets, window: int, threshold: float):
    """One iteration over a ragged batch: packed_in / packed_orig are 1-D CUDA buffers holding image i ([H_i,W_i] dense) at
    element offset offsets[i]; returns the filtered packed buffer (same layout)."""
    _check("packed_in", packed_in, dtype=(torch.float32, torch.float64))
    _check("packed_orig", packed_orig, dtype=packed_in.dtype, shape=packed_in.shape)
    n = len(shapes)
    if any(off + h * w > packed_in.numel() for (h, w), off in zip(shapes, offsets)):
        raise ValueError("an image extends past the end of the packed buffer")
    Hs = (C.c_int * n)(*[int(h) for h, _ in shapes])
    Ws = (C.c_int * n)(*[int(w) for _, w in shapes])
    offs = (C.c_size_t * n)(*[int(o) for o in offsets])
    out = torch.empty_like(packed_in)
    _lib.call("ofd_bilateral_iter_batch", _ptr(packed_in), _ptr(packed_orig), _DT[packed_in.dtype], n, Hs, Ws, offs, int(window),
              C.c_double(threshold), _ptr(out), _stream(packed_in.device))
    return out


PLANE_OP_DTYPE = [("src", "<u8"), ("dst", "<u8"), ("op", "<i4"), ("p", "<f4")]  # struct ofd_plane_op (include/ofd_b200.h)


def plane_ops(table, hw: int, device, gray_stride: Optional[int] = None):
    """ofd_plane_ops: `table` = numpy structured array of PLANE_OP_DTYPE (device addresses of float32 planes of `hw` elements, an
    OFD_PLANE_* code and its parameter); one upload and one launch for the whole table.  The caller keeps the tensors behind the
    addresses alive until the stream has run the launch (ordinary torch stream semantics: they are used on the current stream)."""
    import numpy as np

    table = np.ascontiguousarray(table, dtype=PLANE_OP_DTYPE)
    n = int(table.shape[0])
    if n == 0:
        return
    device = torch.device(device)
    aligned = bool(((table["src"] | table["dst"]) & 15 == 0).all())
    with torch.cuda.device(device):
        # page-locked staging + asynchronous copy: a pageable upload would synchronise the stream (the caching host allocator keeps
        # the staging block until the copy has run)
        dev_table = torch.from_numpy(table.view(np.uint8).reshape(-1)).pin_memory().to(device, non_blocking=True)
        _lib.call("ofd_plane_ops", _ptr(dev_table), n, int(hw), int(hw if gray_stride is None else gray_stride), int(aligned),
                  _stream(device))  # dev_table is allocated and read on the current stream: its reuse is stream-ordered


def pack_u8(t, flag):
    """float32 CUDA tensor -> uint8 tensor of the same shape (ofd_pack_u8); `flag` (int32 CUDA tensor, one element, zeroed by the
    caller) is raised when some value is not exactly a uint8."""
    _check("t", t, dtype=torch.float32)
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        _lib.call("ofd_pack_u8", _ptr(t), _ptr(out), C.c_size_t(t.numel()), _ptr(flag), _stream(t.device))
    return out


def host_widen_u8(src, dst):
    """uint8 CPU tensor -> float32 CPU tensor (same element count, both contiguous), non-temporal stores on the calling thread; the
    native call releases the GIL, so several Python threads widen in parallel."""
    if src.is_cuda or dst.is_cuda or src.dtype != torch.uint8 or dst.dtype != torch.float32 or src.numel() != dst.numel() \
            or not src.is_contiguous() or not dst.is_contiguous():
        raise ValueError("host_widen_u8 needs contiguous CPU tensors: uint8 source, float32 destination, equal element counts")
    _lib.call("ofd_host_widen_u8", C.c_void_p(src.data_ptr()), C.c_size_t(src.numel()), C.c_void_p(dst.data_ptr()))


def scatter_channels_to_host(t, host, c0: int, stream=None, n_channels: Optional[int] = None):
    """Asynchronous D2H copy of t[B,c,H,W] (CUDA, contiguous; float32 or uint8) into channels [c0, c0+c) of the page-locked CPU tensor
    host[>=B,Ctot,H,W] of the same dtype (ofd_copy_rows_to_host: ONE strided DMA, no concatenation on the device).  Runs on `stream`
    (default: torch's current stream of t's device); the caller synchronises before reading `host`.  n_channels: only the first
    n_channels channels of every frame of t are copied (into [c0, c0+n_channels))."""
    _check("t", t, dtype=(torch.float32, torch.uint8))
    if t.dim() != 4 or host.dim() != 4:
        raise ValueError("t and host must be 4-D [B,C,H,W]")
    B, c, H, W = t.shape
    nc = c if n_channels is None else int(n_channels)
    if not 0 < nc <= c:
        raise ValueError(f"n_channels must be in 1..{c}")
    if host.is_cuda or host.dtype != t.dtype or not host.is_contiguous():
        raise ValueError("host must be a contiguous CPU tensor of t's dtype (page-locked for an asynchronous copy)")
    if host.shape[0] < B or tuple(host.shape[2:]) != (H, W) or not (0 <= c0 and c0 + nc <= host.shape[1]):
        raise ValueError(f"host{tuple(host.shape)} cannot take channels [{c0},{c0 + nc}) of a batch of {B} frames {H}x{W}")
    hw4 = H * W * t.element_size()
    st = C.c_void_p(stream.cuda_stream) if stream is not None else _stream(t.device)
    with torch.cuda.device(t.device):
        _lib.call("ofd_copy_rows_to_host", _ptr(t), C.c_size_t(c * hw4), C.c_void_p(host.data_ptr() + c0 * hw4),
                  C.c_size_t(host.shape[1] * hw4), C.c_size_t(nc * hw4), C.c_size_t(B), st)


def host_stream_fill(dst, value: float):
    """ofd_host_stream_fill: a contiguous float32 CPU tensor filled with `value` (sign of zero kept) by non-temporal stores on the
    calling thread (releases the GIL: ctypes call)."""
    if dst.is_cuda or dst.dtype != torch.float32 or not dst.is_contiguous():
        raise ValueError("dst must be a contiguous float32 CPU tensor")
    _lib.call("ofd_host_stream_fill", C.c_void_p(dst.data_ptr()), C.c_size_t(dst.numel()), C.c_float(value))


class PairPipeline:
    """Host-buffer front end (ofd_pair_pipeline_*): pinned CPU tensors in, pinned CPU tensors out.
    Every tensor is validated against the pipeline's own (H, W) before the native call (the C side copies B*C*H*W elements)."""

    def __init__(self, device: int, H: int, W: int, chunk_frames: int = 8):
        self._h = C.c_void_p(0)
        self.H, self.W = H, W
        _lib.call("ofd_pair_pipeline_create", int(device), H, W, int(chunk_frames), C.byref(self._h))

    def _validate(self, spec, B):
        for n, t, dt, shape in spec:
            if t is None:
                continue
            if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{n} must be a contiguous {dt} CPU tensor")
            if tuple(t.shape) != shape:
                raise ValueError(f"{n} must have shape {shape} for this pipeline (B={B}, H={self.H}, W={self.W}), got {tuple(t.shape)}")

    def run(self, img0, depth0, sBf, img1, depth1, back_flow, flow, valid, collision, keep_const_planes: bool = False):
        """keep_const_planes: the result buffers are recycled and their flow.y / back_flow.y planes already hold -0.0 / +0.0
        (OFD_PIPE_KEEP_CONST_PLANES) - the pipeline does not rewrite them."""
        if img0.dim() != 4:
            raise ValueError("img0 must be [B,3,H,W]")
        B, H, W, f = img0.shape[0], self.H, self.W, torch.float32
        self._validate((("img0", img0, f, (B, 3, H, W)), ("depth0", depth0, f, (B, 1, H, W)), ("sBf", sBf, f, (B,)),
                        ("img1", img1, f, (B, 3, H, W)), ("depth1", depth1, f, (B, 1, H, W)),
                        ("back_flow", back_flow, f, (B, 2, H, W)), ("flow", flow, f, (B, 2, H, W)),
                        ("valid", valid, f, (B, 1, H, W)), ("collision", collision, f, (B, 1, H, W))), B)
        _lib.call("ofd_pair_pipeline_run_flags", self._h, _ptr(img0), _ptr(depth0), _ptr(sBf), B, _ptr(img1), _ptr(depth1),
                  _ptr(back_flow), _ptr(flow), _ptr(valid), _ptr(collision),
                  C.c_uint(_lib.PIPE_KEEP_CONST_PLANES if keep_const_planes else 0))

    def run_u8(self, img0_u8, depth0, sBf, img1_u8, depth1, back_flow_x, flow_x, valid_u8, collision_u8):
        """Compact transport (ofd_pair_pipeline_run_u8): uint8 colour / masks, x planes only.  CPU tensors, contiguous."""
        if img0_u8.dim() != 4:
            raise ValueError("img0_u8 must be [B,3,H,W]")
        B, H, W, f, u = img0_u8.shape[0], self.H, self.W, torch.float32, torch.uint8
        self._validate((("img0_u8", img0_u8, u, (B, 3, H, W)), ("depth0", depth0, f, (B, 1, H, W)), ("sBf", sBf, f, (B,)),
                        ("img1_u8", img1_u8, u, (B, 3, H, W)), ("depth1", depth1, f, (B, 1, H, W)),
                        ("back_flow_x", back_flow_x, f, (B, 1, H, W)), ("flow_x", flow_x, f, (B, 1, H, W)),
                        ("valid_u8", valid_u8, u, (B, 1, H, W)), ("collision_u8", collision_u8, u, (B, 1, H, W))), B)
        _lib.call("ofd_pair_pipeline_run_u8", self._h, _ptr(img0_u8), _ptr(depth0), _ptr(sBf), B, _ptr(img1_u8), _ptr(depth1),
                  _ptr(back_flow_x), _ptr(flow_x), _ptr(valid_u8), _ptr(collision_u8))

    def close(self):
        if self._h:
            _lib.load().ofd_pair_pipeline_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
