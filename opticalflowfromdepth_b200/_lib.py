"""ctypes binding of libofd_b200.so (include/ofd_b200.h).  There is NO fallback: if the library is missing the
import of any operator fails loudly, and every non-zero return code raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# OFD_LIB_PATH selects another build of the same ABI (kernel tuning variants); the default is the in-tree library
LIB_PATH = Path(os.environ["OFD_LIB_PATH"]) if os.environ.get("OFD_LIB_PATH") else PKG / "libofd_b200.so"

F32, F64 = 0, 1
EPI_NONE, EPI_CONCAT, EPI_BACK = 0, 1, 2
SRC_RELDEPTH, SRC_DISPARITY = 0, 1
MAX_CHANNELS = 8
PIPE_KEEP_CONST_PLANES = 1
PLANE_COPY, PLANE_SCALE, PLANE_ADD, PLANE_GRAY = 0, 1, 2, 3
CNT_HIT, CNT_HOLE, CNT_COLLISION, CNT_DROPPED, CNT_TIE_SRC, CNT_FRAMES, CNT_PAIRS, CNT_SLOTS = 0, 1, 2, 3, 4, 5, 6, 8

_p, _i, _sz, _f, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes): one entry per symbol declared in include/ofd_b200.h
SIGNATURES = {
    "ofd_version": (_i, []),
    "ofd_last_error_string": (C.c_char_p, []),
    "ofd_workspace_bytes": (_sz, [_i, _i, _i]),
    "ofd_workspace_reset": (_i, [_p, _sz, _p]),
    "ofd_splat_targets": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ofd_splat_flow": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _sz, _p]),
    "ofd_splat_flow_rows": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p]),
    "ofd_disparity_flow": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "ofd_disparity_pair": (_i, [_p, _p, _i, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ofd_disparity_pair_ragged": (_i, [_p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ofd_reproject_flow": (_i, [_p, _i, _p, _f, _i, _i, _i, _p, _p]),
    "ofd_backproject": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "ofd_project": (_i, [_p, _p, _f, _i, _i, _i, _p, _p, _p]),
    "ofd_frame_splat": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ofd_concat_frame_splat": (_i, [_p] * 7 + [_i, _i, _i] + [_p] * 9 + [_sz, _p]),
    "ofd_frame_splat_f64": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ofd_reproject_pair": (_i, [_p, _p, _p, _f, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ofd_normalize_depth": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "ofd_normalize_depth_ragged": (_i, [_p, _i, _i, _p, _p, _p, _p, _p]),
    "ofd_depth_from_png": (_i, [_p, _i, _i, _sz, _p, _i, _p]),
    "ofd_fix_warped_depth": (_i, [_p, _sz, _p]),
    "ofd_inpaint_mask": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "ofd_special_flow": (_i, [_i, _p, _i, _i, _p, _p, _p]),
    "ofd_special_flow_batch": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "ofd_augment_pairs": (_i, [_p] * 8 + [_i, _i, _i] + [_p] * 17 + [_sz, _p]),
    "ofd_bilateral_iter": (_i, [_p, _p, _i, _i, _i, _i, _d, _p, _p]),
    "ofd_bilateral_iter_masked": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _d, _p, _p]),
    "ofd_bilateral_iter_batch": (_i, [_p, _p, _i, _i, _p, _p, _p, _i, _d, _p, _p]),
    "ofd_pair_pipeline_create": (_i, [_i, _i, _i, _i, C.POINTER(_p)]),
    "ofd_pair_pipeline_run": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p]),
    "ofd_pair_pipeline_run_flags": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, C.c_uint]),
    "ofd_jpeg_decoder_create": (_i, [_i, C.POINTER(_p)]),
    "ofd_jpeg_info": (_i, [_p, _p, _sz, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ofd_jpeg_decode": (_i, [_p, _p, _sz, _p, _i, _i, _p]),
    "ofd_jpeg_decoder_destroy": (None, [_p]),
    "ofd_resize_bilinear_aa": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ofd_inpaint_workspace_bytes": (_sz, [_i, _i, _i]),
    "ofd_inpaint_telea": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _sz, _p, _p]),
    "ofd_copy_rows_to_host": (_i, [_p, _sz, _p, _sz, _sz, _sz, _p]),
    "ofd_plane_ops": (_i, [_p, _i, _sz, _sz, _i, _p]),
    "ofd_pack_u8": (_i, [_p, _p, _sz, _p, _p]),
    "ofd_host_widen_u8": (_i, [_p, _sz, _p]),
    "ofd_host_stream_fill": (_i, [_p, _sz, _f]),
    "ofd_pair_pipeline_run_u8": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p]),
    "ofd_pair_pipeline_destroy": (None, [_p]),
}


class OfdError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the shared library once.  Raises if it has not been built (python -m opticalflowfromdepth_b200._build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m opticalflowfromdepth_b200._build` "
                "(or __graft_entry__.build()); there is no CPU or PyTorch fallback for this path"
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise OfdError on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise OfdError(name, rc, lib.ofd_last_error_string().decode("utf-8", "replace"))
