"""Sweep the L2 chunk size of the general splat path on the GPU (CUDA events, inputs larger than L2)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def case(H, W, B, pool=4):
    frames = [synthetic.diml_frame(k, H, W) for k in range(pool)]
    idx = torch.arange(B, device=dev) % pool
    img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)[idx].contiguous()
    depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))[idx].contiguous()
    K, invK = synthesis.Plausible.K((H, W))
    cams = []
    for k in range(B):
        torch.manual_seed(k)
        cams.append(geometry.camera_constants(K, invK, synthesis.Plausible.random_motion(1 / 36, 1 / 36, .1, .1)[0]))
    cam = torch.cat(cams).to(dev)
    vin = torch.ones(B, 1, H, W, device=dev)
    flow = ops.reproject_flow(depth, cam)
    px = B * H * W
    for mode in ("two-launch", "pipe D=auto", "pipe D=1", "pipe D=2", "pipe D=3", "pipe D=8"):
        chunk = mode
        os.environ["OFD_SPLAT_PIPELINE"] = "0" if mode == "two-launch" else "1"
        os.environ.pop("OFD_SPLAT_PIPE_D", None)
        if "D=" in mode and "auto" not in mode:
            os.environ["OFD_SPLAT_PIPE_D"] = mode.split("=")[1]
        t_plane = timeit(lambda: ops.frame_splat(img, depth, flow, vin))
        t_fused = timeit(lambda: ops.reproject_pair(img, depth, cam, vin))
        t_c2 = timeit(lambda: ops.splat_flow(flow, flow, depth, epilogue=ops.EPI_BACK))
        print(f"{H}x{W} B={B} {chunk:12s}: frame_splat(flow plane, 72 B/px) {t_plane*1e6:8.1f} us {72*px/t_plane/1e9:6.0f} GB/s | "
              f"reproject_pair (64 B/px) {t_fused*1e6:8.1f} us {64*px/t_fused/1e9:6.0f} GB/s {B/t_fused:9.0f} fr/s | "
              f"backflow C=2 (36 B/px) {t_c2*1e6:8.1f} us {36*px/t_c2/1e9:6.0f} GB/s", flush=True)


if __name__ == "__main__":
    case(480, 640, 128)
    case(1080, 1920, 16)
