"""ofd_inpaint_telea: time per batch against the number of resident blocks per SM (the grid-wide barrier cost grows with the grid)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic  # noqa: E402

dev = torch.device("cuda:0")
H, W, B = 480, 640, 9
frames = [synthetic.diml_frame(k, H, W) for k in range(B)]
img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)
depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))
pair = synthesis.synthesize_pairs(img, depth, torch.full((B,), 47.0, device=dev))
Kc, invK = synthesis.Plausible.K((H, W))
cams = []
for k in range(B):
    torch.manual_seed(12345 + k)
    cams.append(geometry.camera_constants(Kc, invK, synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)[0]))
six = ops.reproject_pair(pair["img1"], pair["depth1"], torch.cat(cams).to(dev), pair["valid"])
cases = {"stereo (46 layers)": (pair["img1"], ops.inpaint_mask(pair["valid"], pair["collision"])),
         "6-DoF (211 layers)": (six[0], ops.inpaint_mask(six[4], six[5]))}
for bps in ("0", "1", "2", "3", "4"):
    if bps == "0":
        os.environ.pop("OFD_TELEA_BLOCKS_PER_SM", None)
    else:
        os.environ["OFD_TELEA_BLOCKS_PER_SM"] = bps
    for name, (im, mask) in cases.items():
        ops.inpaint_telea(im, mask, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.inpaint_telea(im, mask, 3)
        e1.record()
        torch.cuda.synchronize()
        print(f"blocks/SM {bps if bps != '0' else 'occupancy'}: {name}: {e0.elapsed_time(e1) / 5:.2f} ms per batch of {B}", flush=True)
