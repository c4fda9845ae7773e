"""opticalflowfromdepth_b200 — B200-native (sm_100a) flow-synthesis hot path of AegeanKI/OpticalFlowFromDepth.

Product layout (only what the path needs):
    csrc/                hand-written CUDA kernels + the C ABI (include/ofd_b200.h) -> libofd_b200.so
    _lib.py / ops.py     ctypes binding and tensor-level operators (no fallback: a missing library is an error)
    fw_cuda.py, fw.py    drop-ins for the reference's `fw_cuda` extension and alt_cuda/fw.py `FW`
    geometry.py          BackprojectDepth / Project3D / transformation_from_parameters + fused depth_to_flow
    bilateral_filter.py  sparse_bilateral_filtering
    synthesis.py         Plausible / Convert / ConcatFlow / BackFlow / SpecialFlow / augment_flow, batched
                         synthesize_pairs / synthesize_group
    sweep.py             multi-GPU sharding of frames by image index + end-of-run counter reduction, host-in / host-out sweep
    preprocess.py        the reference's driver (PreprocessPlusAugment, CLI) over the fused path, asynchronous .npz writer
    dataloader.py        the training-side reader of those files (host code)
    inloop.py            in-loop alternative to writing / reading those files: frames -> training samples per batch on the GPU (cfg4)
    synthetic.py         seeded DIML- / ReDWeb-shaped synthetic frames for tests and benches
"""
from . import _lib, ops  # noqa: F401
from .fw import FW  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
