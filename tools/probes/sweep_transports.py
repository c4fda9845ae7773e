"""cfg5 sweep end to end (sweep.run_sweep + PinnedGroupSink) with the sink's transports switched one by one, same box, same run:
frames/s and D2H GB/s for float planes / image bytes / image bytes + host-filled constant planes."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from opticalflowfromdepth_b200 import sweep, synthetic  # noqa: E402

dev = torch.device("cuda:0")
sweep.bind_rank_cores(0, 1)
H, W, POOL, N, B = 480, 640, 16, 4096, 32
pool = [synthetic.diml_frame(k, H, W) for k in range(POOL)]
for rep in range(2):
    for label, kw in (("float planes", dict(byte_images=False, const_planes=False)), ("image bytes", dict(byte_images=True, const_planes=False)),
                      ("image bytes + constant planes host-filled", dict(byte_images=True, const_planes=True)),
                      ("constant planes host-filled only", dict(byte_images=False, const_planes=True))):
        sink = sweep.PinnedGroupSink(**kw)
        sweep.run_sweep(range(3 * B), lambda i: pool[i % POOL], dev, batch=B, dataset_len=N, sink=sink)
        torch.cuda.synchronize()
        sink.frames = sink.bytes = 0
        t0 = time.perf_counter()
        sweep.run_sweep(range(N), lambda i: pool[i % POOL], dev, batch=B, dataset_len=N, sink=sink)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{label:45s} {N / dt:7.0f} frames/s   D2H {sink.bytes / dt / 1e9:5.1f} GB/s   {sink.bytes // sink.frames / (H * W):5.0f} B/px", flush=True)
