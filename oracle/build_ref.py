"""Compile the reference's own forward-warp CUDA extension, unmodified, into oracle/_ref/ (TEST INFRASTRUCTURE).

Sources are compiled where they lie under /root/reference/alt_cuda (fw_cuda.cpp:1-30, fw_cuda_kernel.cu:1-83);
nothing is copied into the repository.  The result `oracle/_ref/fw_cuda.so` is a torch extension module named
`fw_cuda` holding an sm_100a cubin; it is git-ignored but travels to the GPU box, where tests and bench.py load it
to (a) check the product bit-for-bit against the real reference kernel and (b) time "the reference's fw_cuda on
the same B200".  Needs /root/reference, so it only runs in the build container.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/alt_cuda")
OUT = HERE / "_ref"


def build_ref(verbose: bool = False) -> Path | None:
    target = OUT / "fw_cuda.so"
    srcs = [REF_SRC / "fw_cuda.cpp", REF_SRC / "fw_cuda_kernel.cu"]
    if not all(s.exists() for s in srcs):
        return target if target.exists() else None
    if target.exists() and all(target.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return target
    OUT.mkdir(exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension

    cpp_extension.load(
        name="fw_cuda",
        sources=[str(s) for s in srcs],
        build_directory=str(OUT),
        extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"],
        verbose=verbose,
        is_python_module=False,  # do not import here: just build
    )
    return target if target.exists() else None


REF_ROOT = Path("/root/reference")
STAGE = HERE.parent / "baseline" / "_ref"
# the reference's own Python on the flow-synthesis path, staged UNMODIFIED for the GPU box
REF_PY = ["preprocess.py", "utils.py", "dataloader.py", "geometry.py", "bilateral_filter.py", "flow_colors.py",
          "alt_cuda/__init__.py", "alt_cuda/fw.py"]


def stage_reference_python() -> Path | None:
    """Copy the reference's Python modules of the path, byte for byte, into git-ignored `baseline/_ref/` (the base contract's
    place for the unmodified reference; it travels to the GPU box, where /root/reference does not exist).  Used by
    tools/run_reference_on_dropin.py (the reference's preprocess.py running on top of dropin/) and bench.py's ref_fw_cuda leg (the
    reference's FW.forward driving its own kernel).  Nothing under baseline/_ref is imported by the product."""
    import shutil

    if not REF_ROOT.exists():
        return STAGE if (STAGE / "preprocess.py").exists() else None
    for rel in REF_PY:
        src, dst = REF_ROOT / rel, STAGE / rel
        if not src.exists():
            continue
        dst.parent.mkdir(parents=True, exist_ok=True)
        if not dst.exists() or dst.read_bytes() != src.read_bytes():
            shutil.copyfile(src, dst)
    return STAGE


if __name__ == "__main__":
    print(build_ref(verbose="-v" in sys.argv))
    print(stage_reference_python())
