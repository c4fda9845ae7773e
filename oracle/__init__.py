"""CPU oracle of the flow-synthesis hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this package;
the product (opticalflowfromdepth_b200/, dropin/) never does.  See ofd_oracle.c for the parity-pinning statement.

    oracle.build()                      compile liboracle.so (gcc) and, when /root/reference exists, oracle/_ref
    oracle.splat_literal(...)           literal restatement of alt_cuda/fw_cuda_kernel.cu:10-83
    oracle.fw_targets(flow)             alt_cuda/fw.py:27-42
    oracle.fw_forward(obj, flow, depth) FW.forward = targets + splat (+ winner map)
    oracle.disparity_pair(...)          preprocess.py:356-365 minus inpaint (multi-threaded over frames)
    oracle.flow / oracle.bilateral      torch / numpy restatements of preprocess.py, geometry.py, bilateral_filter.py
    oracle.load_ref_fw_cuda()           the reference's own compiled kernel (oracle/_ref/fw_cuda.so), GPU only
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Build liboracle.so (always) and oracle/_ref/fw_cuda.so (only where /root/reference exists)."""
    src = HERE / "ofd_oracle.c"
    if not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if ref and Path("/root/reference/alt_cuda/fw_cuda_kernel.cu").exists():
        from .build_ref import build_ref, stage_reference_python

        build_ref()
        stage_reference_python()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build(ref=False)
        L = C.CDLL(str(LIB))
        L.oracle_splat_literal.restype = C.c_int
        L.oracle_splat_literal.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.oracle_fw_targets_f32.restype = None
        L.oracle_fw_targets_f32.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p]
        L.oracle_fw_targets_f64.restype = None
        L.oracle_fw_targets_f64.argtypes = [_f64p, C.c_int, C.c_int, _f32p, _f32p]
        L.oracle_splat_frame.restype = None
        L.oracle_splat_frame.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _i32p, _f32p, C.c_void_p]
        L.oracle_disparity_pair.restype = C.c_int
        L.oracle_disparity_pair.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int]
        _lib = L
    return _lib


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def splat_literal(obj, safe_y, safe_x, depth):
    """Literal fw_cuda.forward_warping on numpy float32 arrays [B,C,H,W] / [B,1,H,W] -> (out, valid, collision, rc)."""
    obj, safe_y, safe_x, depth = _c(obj), _c(safe_y), _c(safe_x), _c(depth)
    B, Cc, H, W = obj.shape
    out = np.empty_like(obj)
    valid = np.empty_like(depth)
    coll = np.empty_like(depth)
    rc = lib().oracle_splat_literal(obj, safe_y, safe_x, depth, B, Cc, H, W, out, valid, coll)
    return out, valid, coll, rc


def fw_targets(flow):
    """alt_cuda/fw.py:27-42 on flow[2,H,W] (float32 or float64) -> (safe_x, safe_y) float32 [H,W]; NaN -> -1e30."""
    flow = np.ascontiguousarray(flow)
    _, H, W = flow.shape
    sx = np.empty((H, W), np.float32)
    sy = np.empty((H, W), np.float32)
    if flow.dtype == np.float64:
        lib().oracle_fw_targets_f64(flow, H, W, sx, sy)
    else:
        lib().oracle_fw_targets_f32(_c(flow), H, W, sx, sy)
    return sx, sy


def splat_frame(obj, safe_y, safe_x, depth):
    """Shared-LUT splat of one frame with winner map: obj[C,H,W] -> (out, valid[1,H,W], collision[1,H,W], winner[H,W], dropped)."""
    obj, safe_y, safe_x, depth = _c(obj), _c(safe_y), _c(safe_x), _c(depth)
    Cc, H, W = obj.shape
    out = np.empty_like(obj)
    valid = np.empty((1, H, W), np.float32)
    coll = np.empty((1, H, W), np.float32)
    winner = np.empty((H, W), np.int32)
    dlut = np.empty((H, W), np.float32)
    dropped = C.c_int64(0)
    lib().oracle_splat_frame(obj, safe_y.reshape(H, W), safe_x.reshape(H, W), depth.reshape(H, W), Cc, H, W, out, valid, coll,
                             winner, dlut, C.addressof(dropped))
    return out, valid, coll, winner, dropped.value


def fw_forward(obj, flow, depth):
    """FW.forward (alt_cuda/fw.py:19-59) on one frame: obj[C,H,W], flow[2,H,W] (f32|f64), depth[1,H,W]."""
    sx, sy = fw_targets(flow)
    return splat_frame(_c(obj), sy, sx, _c(depth))


def disparity_pair(img0, depth0, sBf, nthreads=1):
    """preprocess.py:356-365 minus inpaint on float32 batches; returns (img1, depth1, back_flow, flow, valid, collision)."""
    img0, depth0, sBf = _c(img0), _c(depth0), _c(sBf)
    B, _, H, W = img0.shape
    img1 = np.empty((B, 3, H, W), np.float32)
    depth1 = np.empty((B, 1, H, W), np.float32)
    back = np.empty((B, 2, H, W), np.float32)
    flow = np.empty((B, 2, H, W), np.float32)
    valid = np.empty((B, 1, H, W), np.float32)
    coll = np.empty((B, 1, H, W), np.float32)
    rc = lib().oracle_disparity_pair(img0, depth0, sBf, B, H, W, img1, depth1, back, flow, valid, coll, int(nthreads))
    if rc:
        raise RuntimeError(f"oracle_disparity_pair rc={rc}")
    return img1, depth1, back, flow, valid, coll


def load_ref_fw_class():
    """The reference's own `FW` module class (alt_cuda/fw.py, staged unmodified in baseline/_ref) bound to the reference's own
    compiled kernel (oracle/_ref/fw_cuda.so): `FW(device).forward(obj, flow, depth)` is then the reference's stock forward-warp
    call, prologue included.  TEST / BENCH INFRASTRUCTURE only."""
    import importlib.util
    import sys

    path = HERE.parent / "baseline" / "_ref" / "alt_cuda" / "fw.py"
    if not path.exists():
        raise FileNotFoundError(f"{path} missing: run __graft_entry__.build() in the build container")
    saved = sys.modules.get("fw_cuda")
    sys.modules["fw_cuda"] = load_ref_fw_cuda()
    try:
        spec = importlib.util.spec_from_file_location("ref_alt_cuda_fw", str(path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["fw_cuda"] = saved
        else:
            sys.modules.pop("fw_cuda", None)
    return mod.FW


def load_ref_fw_cuda():
    """Import the reference's own compiled extension (oracle/_ref/fw_cuda.so) as a module; needs a GPU to run."""
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    path = HERE / "_ref" / "fw_cuda.so"
    if not path.exists():
        raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` in the build container")
    spec = importlib.util.spec_from_file_location("fw_cuda", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
