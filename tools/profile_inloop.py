"""Host-time breakdown of one inloop.InLoopSampler step (synchronised sections) and its throughput against the batch size."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import inloop, ops, synthesis, synthetic  # noqa: E402

dev = torch.device("cuda:0")
H, W = 368, 496
for B in (8, 32, 128):
    fr = [synthetic.diml_frame(200 + k % 16, H, W) for k in range(B)]
    img = torch.from_numpy(np.stack([f[0] for f in fr])).to(dev)
    dep = torch.from_numpy(np.stack([f[1] for f in fr])).to(dev)
    s = inloop.InLoopSampler(dev, seed=4)
    for _ in range(5):
        s(img, dep).raft_tuple()
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        plan = s.draw(B, (H, W))
    t_draw = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        s(img, dep, plan).raft_tuple()
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / n
    d0 = ops.normalize_depth(dep)
    sBf, cam = plan.sBf.to(dev), plan.cam.to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        synthesis.synthesize_group(img, d0, sBf, cam)
    torch.cuda.synchronize()
    t_grp = (time.perf_counter() - t0) / n
    print(f"B={B:4d}: draw {1e3 * t_draw:6.2f} ms, step with a given plan {1e3 * t_dev:6.2f} ms (of which the 5-pair group {1e3 * t_grp:6.2f} ms) "
          f"-> {B / (t_draw + t_dev):8.0f} samples/s")
