"""Run the two-launch splat over 64 frames of 480x640 in chunks (OFD_SPLAT_CHUNK_FRAMES) for an ncu traffic capture."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
H, W, B, pool = 480, 640, 64, 4
frames = [synthetic.diml_frame(k, H, W) for k in range(pool)]
idx = torch.arange(B, device=dev) % pool
img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)[idx].contiguous()
depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))[idx].contiguous()
flow = torch.randn(B, 2, H, W, device=dev) * 20
for _ in range(3):
    ops.frame_splat(img, depth, flow, None)
torch.cuda.synchronize()
print("ok")
