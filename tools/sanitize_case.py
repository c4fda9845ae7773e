"""One tiny invocation of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import bilateral_filter, fw_cuda, geometry, ops, synthesis, synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, H, W) in ((2, 24, 40), (3, 17, 33)):
    fr = [synthetic.diml_frame(k, max(H, 16), max(W, 16)) for k in range(B)]
    img = torch.from_numpy(np.stack([f[0][:, :H, :W] for f in fr])).contiguous().to(dev)
    depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1][:, :H, :W] for f in fr])).contiguous().to(dev))
    sBf = torch.full((B,), 47.0, device=dev)
    cnt = ops.new_counters(dev)
    ops.disparity_pair(img, depth, sBf, counters=cnt)                       # persistent TMA kernel (W % 4 == 0) / one-row kernel
    ops.disparity_pair(img, depth.double(), sBf)
    flow = ops.disparity_flow(depth, sBf)
    K, invK = synthesis.Plausible.K((H, W))
    T1, _, _ = synthesis.Plausible.random_motion(1 / 36, 1 / 36, .1, .1)
    cam = geometry.camera_constants(K, invK, T1).repeat(B, 1).to(dev)
    f6 = ops.reproject_flow(depth, cam)
    obj = torch.cat((img, depth, f6 * -1.0), 1).contiguous()
    ops.splat_flow(obj, f6, depth, want_winner=True, counters=cnt)
    ops.splat_flow(obj, f6.double(), depth)
    ops.splat_flow(f6, f6, depth, epilogue=ops.EPI_BACK)
    ops.splat_flow(f6, flow, depth, epilogue=ops.EPI_CONCAT, aux=flow)
    ops.frame_splat(img, depth, f6, torch.ones_like(depth), want_raw_valid=True, counters=cnt)
    ops.reproject_pair(img, depth, cam, None)
    gx, gy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    sx = gx.float().to(dev).expand(B, 1, H, W).contiguous()
    sy = gy.float().to(dev).expand(B, 1, H, W).contiguous()
    fw_cuda.forward_warping(obj, sy, sx, depth)
    fw_cuda.forward_warping(obj.double(), sy.double(), sx.double(), depth.double())
    for kind in (5, 6, 7):
        synthesis.SpecialFlow(dev)((H, W), float(kind))
    bilateral_filter.sparse_bilateral_filtering(depth[0, 0].contiguous(), None, [5, 3], num_iter=2)
    bp = geometry.BackprojectDepth(B, H, W, dev)
    pj = geometry.Project3D(B, H, W)
    pj(bp(depth, invK.repeat(B, 1, 1).to(dev)), K.repeat(B, 1, 1).to(dev), T1.repeat(B, 1, 1).to(dev))
    synthesis.fix_warped_depth(depth.clone())
# ragged batch through the one-launch pair kernel: units on every 16-byte phase, head / tail scalar stores, float64 depth
sizes = [(8, 1025), (10, 1026), (12, 682), (4, 7), (16, 32), (2, 2)]
offs = [0]
for h, w in sizes[:-1]:
    offs.append(offs[-1] + h * w)
P = offs[-1] + sizes[-1][0] * sizes[-1][1]
rimg = torch.randint(0, 256, (3 * P,), device=dev).float()
rdep = torch.randint(1, 60, (P,), device=dev).float()
rs = torch.full((len(sizes),), 47.0, device=dev)
ops.disparity_pair_ragged(rimg, rdep, rs, sizes, offs, counters=cnt)
ops.disparity_pair_ragged(rimg, rdep.double(), rs, sizes, offs)
# host pipeline with byte-packed masks
pipe = ops.PairPipeline(0, 24, 40, chunk_frames=2)
himg = torch.randint(0, 256, (5, 3, 24, 40)).float().pin_memory()
hdep = torch.randint(1, 60, (5, 1, 24, 40)).float().pin_memory()
houts = [torch.empty((5, c, 24, 40)).pin_memory() for c in (3, 1, 2, 2, 1, 1)]
pipe.run(himg, hdep, torch.full((5,), 47.0), *houts)
pipe.close()
torch.cuda.synchronize()
print("sanitize_case ok", cnt.tolist())
