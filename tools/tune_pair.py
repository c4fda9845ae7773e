"""Time the fused pair kernel for every build variant in opticalflowfromdepth_b200/build/variants/pair_*.so."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys
import numpy as np, torch
sys.path.insert(0, %r)
from opticalflowfromdepth_b200 import ops, synthetic
dev = torch.device("cuda:0")
H, W, B, pool = 480, 640, 256, 8
frames = [synthetic.diml_frame(k, H, W) for k in range(pool)]
idx = torch.arange(B, device=dev) %% pool
img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)[idx].contiguous()
depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))[idx].contiguous()
sBf = torch.linspace(40, 55, B, device=dev)
out = ops.disparity_pair(img, depth, sBf)
def run(): ops.disparity_pair(img, depth, sBf, out=out)
for _ in range(5): run()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): run()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30 * 1e-3)
print(f"{best*1e6:7.1f} us/step  {B/best:9.0f} pairs/s  {56*B*H*W/best/1e9:6.0f} GB/s  {56*B*H*W/best/1e9/6553.6:.3f} of measured peak")
''' % str(ROOT)

if __name__ == "__main__":
    vdir = ROOT / "opticalflowfromdepth_b200" / "build" / "variants"
    for so in [None] + sorted(vdir.glob("pair_*.so")):
        env = dict(os.environ)
        if so is not None:
            env["OFD_LIB_PATH"] = str(so)
        else:
            env.pop("OFD_LIB_PATH", None)
            so = Path("default")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        if os.environ.get("OFD_DEBUG"):
            print("\n".join(sorted(set(l for l in r.stdout.splitlines() if l.startswith("[ofd]")))))
        print(f"{so.stem:16s}: {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else 'no output'}", flush=True)
