"""Build libofd_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI).

The .so is written next to this file so it travels to the GPU box with the repo snapshot; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libofd_b200.so"
STAMP = PKG / ".libofd_b200.stamp"
SOURCES = ["ofd_abi.cu", "ofd_splat.cu", "ofd_splat_f64.cu", "ofd_pair.cu", "ofd_flow.cu", "ofd_bilateral.cu", "ofd_host.cu", "ofd_inpaint.cu", "ofd_decode.cu", "ofd_assemble.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--compiler-options", "-fPIC",
    "--expt-relaxed-constexpr",
    # IEEE division / sqrt and NO fast-math: the kernels must round like the reference's torch ops
    "-prec-div=true", "-prec-sqrt=true", "-fmad=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        h.update(name.encode())
        h.update((CSRC / name).read_bytes())
    h.update((PKG.parent / "include" / "ofd_b200.h").read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name: str, defines: list[str]) -> Path:
    """Tuning helper: build csrc/ with extra -D flags into build/variants/<name>.so (load it with OFD_LIB_PATH)."""
    vdir = PKG / "build" / "variants"
    vdir.mkdir(parents=True, exist_ok=True)
    out = vdir / f"{name}.so"
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", str(out), *[str(CSRC / s) for s in SOURCES], "-lcudart", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout)
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu of csrc/ into libofd_b200.so (skipped when sources are unchanged)."""
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(str(obj))
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
