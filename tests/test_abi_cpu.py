"""The C-ABI library loads without a GPU, exports every symbol include/ofd_b200.h declares, and rejects bad
arguments before touching CUDA."""
import ctypes as C
import re
from pathlib import Path

import pytest

from opticalflowfromdepth_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "ofd_b200.h").read_text()


def declared_symbols():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(ofd_[a-z0-9_]+)\s*\(", body)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ofd_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_workspace_bytes():
    lib = _lib.load()
    assert lib.ofd_version() == 100
    assert lib.ofd_workspace_bytes(0, 480, 640) == 0
    n = lib.ofd_workspace_bytes(2, 480, 640)
    assert n >= 2 * 480 * 640 * 8 and n % 256 == 0


@pytest.mark.parametrize("name,args,code", [
    ("ofd_splat_flow", (1, 1, 0, 1, 1, 9, 4, 4, 1, 1, 1, None, 0, None, None, 1, 1 << 20, None), -2),   # C = 9
    ("ofd_splat_flow", (1, 1, 7, 1, 1, 2, 4, 4, 1, 1, 1, None, 0, None, None, 1, 1 << 20, None), -3),   # flow dtype
    ("ofd_splat_flow", (1, 1, 0, 1, 1, 2, 4, 4, 1, 1, 1, None, 5, None, None, 1, 1 << 20, None), -4),   # epilogue
    ("ofd_splat_flow", (1, 1, 0, 1, 1, 2, 4, 4, 1, 1, 1, None, 1, None, None, 1, 1 << 20, None), -1),   # concat w/o aux
    ("ofd_splat_flow", (1, 1, 0, 1, 1, 2, 4, 4, 1, 1, 1, None, 0, None, None, 8, 16, None), -5),        # ws too small
    ("ofd_splat_flow", (1, 1, 0, 1, 1, 2, 65536, 32768, 1, 1, 1, None, 0, None, None, 8, 1 << 20, None), -2),  # H*W = 2^31
    ("ofd_splat_flow", (1, 1, 0, 1, 1, 2, 600000, 8, 1, 1, 1, None, 0, None, None, 8, 1 << 20, None), -2),     # H / 8 > 65535 grid rows
    ("ofd_disparity_pair", (1, 1, 0, 1, 1, 65536, 32768, 1, 1, 1, 1, 1, 1, None, None), -2),                   # H*W = 2^31
    ("ofd_augment_pairs", (1, 1, 1, 1, 1, 1, None, None, 2, 8, 8) + (1,) * 15 + (None, 8, 1 << 20, None), -1), # NULL kinds
    ("ofd_bilateral_iter_batch", (1, 1, 0, 1, None, None, None, 5, 0.04, 1, None), -1),                        # NULL tables
    ("ofd_disparity_pair_ragged", (1, 1, 0, 1, 2, None, None, None, 1, 1, 1, 1, 1, 1, None, None), -1),        # NULL shape tables
    ("ofd_disparity_pair_ragged", (1, 1, 0, 1, -1, None, None, None, 1, 1, 1, 1, 1, 1, None, None), -2),       # negative n_images
    ("ofd_disparity_pair_ragged", (1, 1, 7, 1, 1, None, None, None, 1, 1, 1, 1, 1, 1, None, None), -3),        # bad depth dtype
    ("ofd_bilateral_iter_masked", (1, 1, 1, 0, 0, 8, 8, 9, 0.04, 1, None), -4),                                # window 9 with a mask
    ("ofd_bilateral_iter_masked", (1, 1, None, 0, 0, 8, 8, 5, 0.04, 1, None), -1),                             # NULL mask
    ("ofd_depth_from_png", (1, 12, 0, 16, 1, 1, None), -3),                                                    # 12-bit source
    ("ofd_depth_from_png", (1, 8, 7, 16, 1, 1, None), -4),                                                     # unknown kind
    ("ofd_splat_targets", (1, 1, 1, 1, 5, 1, 2, 4, 4, 1, 1, 1, None, None, 8, 1 << 20, None), -3),      # bad dtype
    ("ofd_splat_targets", (1, 1, 1, 1, 1, 1, 2, 4, 4, 1, 1, 1, None, None, 8, 200, None), -5),          # f64: 2 key planes
    ("ofd_splat_targets", (None, 1, 1, 1, 0, 1, 2, 4, 4, 1, 1, 1, None, None, 8, 1 << 20, None), -1),   # NULL obj
    ("ofd_disparity_pair", (1, 1, 3, 1, 1, 4, 4, 1, 1, 1, 1, 1, 1, None, None), -3),
    ("ofd_bilateral_iter", (1, 1, 0, 8, 8, 4, 0.04, 1, None), -4),                                      # even window
    ("ofd_bilateral_iter", (1, 1, 0, 2, 8, 3, 0.04, 1, None), -2),                                      # H < 3
    ("ofd_special_flow", (4, None, 8, 8, 1, 1, None), -4),
    ("ofd_normalize_depth", (1, 2, 1, 4, 4, 1, 8, None), -3),
])
def test_bad_arguments_are_rejected_without_a_gpu(name, args, code):
    lib = _lib.load()
    rc = getattr(lib, name)(*args)
    assert rc == code, lib.ofd_last_error_string()
    assert lib.ofd_last_error_string().decode().startswith(name)
    with pytest.raises(_lib.OfdError):
        _lib.call(name, *args)


def test_empty_inputs_are_a_no_op():
    lib = _lib.load()
    assert lib.ofd_splat_flow(None, None, 0, None, 0, 2, 4, 4, None, None, None, None, 0, None, None, None, 0, None) == 0
    assert lib.ofd_disparity_pair(None, None, 0, None, 3, 0, 7, None, None, None, None, None, None, None, None) == 0
    assert lib.ofd_workspace_reset(None, 0, None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()


@pytest.mark.parametrize("name,args,code", [
    ("ofd_splat_flow_rows", (1, 1, 1, 1, 6, 4, 4, 1, 1, None, 0, None, None, None), -2),                 # C != 2
    ("ofd_splat_flow_rows", (1, 1, 1, 1, 2, 4, 4096, 1, 1, None, 0, None, None, None), -2),              # W > 2048
    ("ofd_splat_flow_rows", (1, 1, 1, 1, 2, 4, 4, 1, 1, None, 1, None, None, None), -1),                 # concat without aux
    ("ofd_inpaint_telea", (1, 1, 1, 8, 8, 9, 1, 256, 1 << 20, None, None), -4),                          # range 9
    ("ofd_inpaint_telea", (1, 1, 1, 8, 8, 3, 1, 256, 16, None, None), -5),                               # workspace too small
    ("ofd_inpaint_telea", (None, 1, 1, 8, 8, 3, 1, 256, 1 << 20, None, None), -1),                       # NULL image
    ("ofd_copy_rows_to_host", (1, 8, 1, 16, 32, 2, None), -4),                                           # pitch < width
    ("ofd_resize_bilinear_aa", (1, 5, 1, 4, 4, 8, 8, 1, 1, None), -3),                                   # bad dtype
    ("ofd_resize_bilinear_aa", (1, 0, 1, 4, 4, 8, 8, 1, None, None), -1),                                # two-axis resize without tmp
    ("ofd_resize_bilinear_aa", (1, 0, 1, 4, 4, 0, 8, 1, 1, None), -2),                                   # empty output
    ("ofd_pair_pipeline_run_flags", (None, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0), -1),                        # NULL pipeline
    ("ofd_plane_ops", (8, -1, 16, 16, 1, None), -2),                                                     # negative table size
    ("ofd_plane_ops", (None, 3, 16, 16, 1, None), -1),                                                   # NULL table
    ("ofd_plane_ops", (4, 3, 16, 16, 1, None), -4),                                                      # table not 8-byte aligned
    ("ofd_concat_frame_splat", (16, 16, 16, 16, None, 16, 16, 1, 4, 6, 16, 16, 16, 16, 16, 16, None, None, 256, 1 << 20, None), -2),   # W % 4 != 0
    ("ofd_concat_frame_splat", (16, 16, 16, 16, None, 16, 16, 1, 4, 4096, 16, 16, 16, 16, 16, 16, None, None, 256, 1 << 20, None), -2),  # W > 2048
    ("ofd_concat_frame_splat", (16, 16, 16, None, None, 16, 16, 1, 4, 8, 16, 16, 16, 16, 16, 16, None, None, 256, 1 << 20, None), -1),  # NULL flowAB
    ("ofd_concat_frame_splat", (16, 16, 16, 16, None, 16, 16, 1, 4, 8, 16, 16, 16, 16, 16, 16, None, None, 256, 16, None), -5),         # workspace too small
    ("ofd_concat_frame_splat", (16, 20, 16, 16, None, 16, 16, 1, 4, 8, 16, 16, 16, 16, 16, 16, None, None, 256, 1 << 20, None), -4),    # plane not 16-byte aligned
])
def test_round2_entry_points_reject_bad_arguments(name, args, code):
    """The entry points added in round 2 validate before touching CUDA (no GPU needed)."""
    lib = _lib.load()
    fn = getattr(lib, name)
    conv = []
    for a, t in zip(args, fn.argtypes):
        conv.append(None if a is None and t in (C.c_void_p,) else a)
    assert fn(*conv) == code, lib.ofd_last_error_string()
