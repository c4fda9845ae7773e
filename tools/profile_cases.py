"""Short, single-purpose workloads for ncu (one kernel family per case, a few launches):
    python tools/profile_cases.py telea | bilateral | group | sixdof
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import bilateral_filter as bfm  # noqa: E402
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic  # noqa: E402

dev = torch.device("cuda:0")
case = sys.argv[1] if len(sys.argv) > 1 else "telea"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 480, 640


def frames(B, h=H, w=W, seed0=0):
    fr = [synthetic.diml_frame(seed0 + k, h, w) for k in range(B)]
    img = torch.from_numpy(np.stack([f[0] for f in fr])).to(dev)
    depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in fr])).to(dev))
    return img, depth


def cams(B, h=H, w=W):
    Kc, invK = synthesis.Plausible.K((h, w))
    out = []
    for k in range(B):
        torch.manual_seed(12345 + k)
        out.append(geometry.camera_constants(Kc, invK, synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)[0]))
    return torch.cat(out).to(dev)


if case == "telea":
    B = 9
    img, depth = frames(B)
    pair = synthesis.synthesize_pairs(img, depth, torch.full((B,), 47.0, device=dev))
    six = ops.reproject_pair(pair["img1"], pair["depth1"], cams(B), pair["valid"])
    mask = ops.inpaint_mask(six[4], six[5])
    for _ in range(reps):
        ops.inpaint_telea(six[0], mask, 3)
elif case == "bilateral":
    sizes = synthetic.redweb_sizes(16, seed=1)
    deps = [torch.from_numpy(synthetic.redweb_frame(k, hh, ww)[1]).to(dev)[0] for k, (hh, ww) in enumerate(sizes)]
    for _ in range(reps):
        bfm.sparse_bilateral_filtering_batch(deps, [7, 7, 5, 5, 5], depth_threshold=0.04, num_iter=5, normalize=True)
elif case == "group":
    B = 128
    img, depth = frames(16)
    idx = torch.arange(B, device=dev) % 16
    img, depth = img[idx].contiguous(), depth[idx].contiguous()
    sBf = torch.full((B,), 47.0, device=dev)
    cam = cams(B)
    for _ in range(reps):
        synthesis.synthesize_group(img, depth, sBf, cam)
elif case == "sixdof":
    B, h, w = 32, 1080, 1920
    img, depth = frames(2, h, w, 100)
    idx = torch.arange(B, device=dev) % 2
    img, depth = img[idx].contiguous(), depth[idx].contiguous()
    vin = torch.ones(B, 1, h, w, device=dev)
    cam = cams(B, h, w)
    for _ in range(reps):
        ops.reproject_pair(img, depth, cam, vin)
else:
    raise SystemExit(f"unknown case {case}")
torch.cuda.synchronize()
print("done", case)
