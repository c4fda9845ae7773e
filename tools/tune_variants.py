"""Time the two-launch splat path for every kernel build variant in opticalflowfromdepth_b200/build/variants/."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic
dev = torch.device("cuda:0")
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
H, W, B, pool = 480, 640, 128, 4
frames = [synthetic.diml_frame(k, H, W) for k in range(pool)]
idx = torch.arange(B, device=dev) %% pool
img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)[idx].contiguous()
depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))[idx].contiguous()
K, invK = synthesis.Plausible.K((H, W))
cams = []
for k in range(B):
    torch.manual_seed(k)
    cams.append(geometry.camera_constants(K, invK, synthesis.Plausible.random_motion(1/36, 1/36, .1, .1)[0]))
cam = torch.cat(cams).to(dev)
vin = torch.ones(B, 1, H, W, device=dev)
flow = ops.reproject_flow(depth, cam)
obj = torch.cat((img, depth, flow * -1.0), 1).contiguous()
px = B * H * W
t7 = timeit(lambda: ops.frame_splat(img, depth, flow, vin))
tf = timeit(lambda: ops.reproject_pair(img, depth, cam, vin))
t6 = timeit(lambda: ops.splat_flow(obj, flow, depth))
t2 = timeit(lambda: ops.splat_flow(flow, flow, depth, epilogue=ops.EPI_BACK))
print(f"frame C=7 {t7*1e6:7.1f} us {72*px/t7/1e9:5.0f} GB/s | reproject_pair {tf*1e6:7.1f} us {B/tf:8.0f} fr/s | FW C=6 (68 B/px) {t6*1e6:7.1f} us {68*px/t6/1e9:5.0f} GB/s | back C=2 {t2*1e6:7.1f} us {36*px/t2/1e9:5.0f} GB/s")
''' % str(ROOT)

if __name__ == "__main__":
    vdir = ROOT / "opticalflowfromdepth_b200" / "build" / "variants"
    for so in sorted(vdir.glob("*.so")):
        env = dict(os.environ, OFD_LIB_PATH=str(so))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        print(f"{so.stem:10s}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else 'no output'}", flush=True)
