"""Fused pair kernel across row widths; arguments = rows per work unit to force (OFD_PAIR_GROUP), default = the launcher's choice (shown as a). CUDA events, 56 B/px."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


if __name__ == "__main__":
    targets = [int(a) for a in sys.argv[1:]] or [0]
    for (H, W, B) in ((480, 640, 256), (480, 642, 256), (480, 641, 256), (1080, 1920, 32), (1080, 1922, 32), (368, 496, 256),
                      (300, 1000, 128), (300, 1002, 128), (512, 384, 256), (1080, 1001, 32), (1080, 2562, 16), (375, 1242, 64), (481, 641, 64)):
        img = torch.rand(B, 3, H, W, device=dev) * 255
        depth = torch.rand(B, 1, H, W, device=dev) * 98 + 1
        s = torch.full((B,), 47.0, device=dev)
        out = ops.disparity_pair(img, depth, s)
        line = f"{H}x{W} B={B}:"
        for tp in targets:
            if tp:
                os.environ["OFD_PAIR_GROUP"] = str(tp)
            t = timeit(lambda: ops.disparity_pair(img, depth, s, out=out))
            line += f"  rows/unit {tp or chr(97)}: {t*1e6:7.1f} us {56*B*H*W/t/1e9/6553.6*100:5.1f}%"
        os.environ.pop("OFD_PAIR_GROUP", None)
        print(line, flush=True)
        del img, depth, out
