"""A/B of the host pipeline's knobs on one GPU (bench.py's e2e shape: 128 frames 480x640 per call, chunk 8):
OFD_HOST_SYNC spin|block x OFD_HOST_WORKERS, with and without torch's CPU thread pool having just run."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import ops  # noqa: E402

H, W, Fe = 480, 640, 128
rng = np.random.default_rng(0)
h_img = torch.from_numpy(rng.integers(0, 256, (Fe, 3, H, W)).astype(np.float32)).pin_memory()
h_dep = torch.from_numpy((rng.random((Fe, 1, H, W)) * 98 + 1).astype(np.float32)).pin_memory()
h_s = torch.full((Fe,), 47.0)
h_out = [torch.empty((Fe, c, H, W), dtype=torch.float32).pin_memory() for c in (3, 1, 2, 2, 1, 1)]


def run(label, env, K=10, stir=False):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    pipe = ops.PairPipeline(0, H, W, chunk_frames=8)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    for _ in range(2):
        pipe.run(h_img, h_dep, h_s, *h_out)
    if stir:  # what bench.py does right before its e2e leg: a multi-threaded torch CPU op (the OpenMP pool then spins for a while)
        torch.rand(64, 3, H, W).mul_(2.0).sum()
    t0 = time.perf_counter()
    for _ in range(K):
        pipe.run(h_img, h_dep, h_s, *h_out)
    dt = (time.perf_counter() - t0) / K
    pipe.close()
    print(f"{label:60s} {Fe / dt:7.0f} pairs/s", flush=True)


for rep in range(2):
    for img_bytes in ("1", "0"):
        for w in ("1", "2", "3", "4", "6"):
            run(f"img1 as {'verified bytes' if img_bytes == '1' else 'float planes'} workers={w}", {"OFD_HOST_IMG_BYTES": img_bytes, "OFD_HOST_WORKERS": w})
    run("sync=block img bytes workers=3", {"OFD_HOST_SYNC": "block", "OFD_HOST_WORKERS": "3"})
