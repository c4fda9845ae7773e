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


if __name__ == "__main__":
    print(build_ref(verbose="-v" in sys.argv))
