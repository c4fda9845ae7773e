"""Per-call cost of the zero-edit drop-in (`FW.forward`, B = 1) at 480x640: wall clock per call with and without a synchronize,
next to the reference's own extension (oracle/_ref) driven by the reference's prologue."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import ops, synthetic  # noqa: E402
from opticalflowfromdepth_b200.fw import FW  # noqa: E402

dev = torch.device("cuda:0")
H, W = 480, 640
img, raw = synthetic.diml_frame(0, H, W)
img = torch.from_numpy(img).to(dev)
depth = ops.normalize_depth(torch.from_numpy(raw).to(dev)[None])[0]
flow = ops.disparity_flow(depth[None], torch.tensor([47.0], device=dev))[0]
fw = FW(dev)
for C, obj in ((6, torch.cat((img, depth, flow * -1.0))), (2, flow.clone()), (4, torch.cat((img, depth)))):
    for _ in range(20):
        fw(obj, flow, depth)
    torch.cuda.synchronize()
    n = 500
    t0 = time.perf_counter()
    for _ in range(n):
        fw(obj, flow, depth)
    t_issue = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fw(obj, flow, depth)
    e1.record()
    torch.cuda.synchronize()
    print(f"FW.forward C={C} 480x640: host issue {t_issue*1e6:6.1f} us/call, wall {t_total*1e6:6.1f} us/call, device {e0.elapsed_time(e1)/n*1e3:6.1f} us/call", flush=True)
try:
    import oracle
    ref = oracle.load_ref_fw_cuda()
    obj = torch.cat((img, depth, flow * -1.0))[None].contiguous()
    gx, gy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    p0 = torch.stack((gx, gy), 0).float()[None].to(dev)
    p1 = p0 + flow[None]
    sy = torch.clamp(p1[:, 1:2], min=0, max=H - 1).contiguous().long().float()
    sx = torch.clamp(p1[:, 0:1], min=0, max=W - 1).contiguous().long().float()
    ref.forward_warping(obj, sy, sx, depth[None])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ref.forward_warping(obj, sy, sx, depth[None])
    torch.cuda.synchronize()
    print(f"reference fw_cuda.forward_warping C=6 480x640: {(time.perf_counter()-t0)/3*1e3:.1f} ms/call")
except Exception as e:  # noqa: BLE001
    print("reference extension unavailable:", e)
