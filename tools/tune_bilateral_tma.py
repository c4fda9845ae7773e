"""VERDICT r1 next #7b: the bilateral raw tile by TMA tensor-map copies (OFD_BIL_TMA=1, cp.async.bulk.tensor.2d / UTMALDG) against the default
per-thread loads, dense single frames, 5 iterations [7,7,5,5,5], CUDA events."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import bilateral_filter as bfm  # noqa: E402
from opticalflowfromdepth_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
for (h, w) in ((480, 640), (1080, 1920), (2160, 3840)):
    _, depth = synthetic.redweb_frame(1, h, w)
    d = ops.normalize_depth(torch.from_numpy(depth).to(dev)[None])[0, 0].contiguous()
    res = {}
    for mode in ("0", "1", "0", "1"):
        os.environ["OFD_BIL_TMA"] = mode
        out = bfm.sparse_bilateral_filtering(d, None, [7, 7, 5, 5, 5], depth_threshold=0.04, num_iter=5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            bfm.sparse_bilateral_filtering(d, None, [7, 7, 5, 5, 5], depth_threshold=0.04, num_iter=5)
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(mode, []).append(e0.elapsed_time(e1) / 20)
        res["out" + mode] = out
    same = torch.equal(res["out0"], res["out1"])
    print(f"{h}x{w}: default {min(res['0']):.4f} ms, TMA tile load {min(res['1']):.4f} ms per 5 iterations ({h * w * 5 / min(res['0']) / 1e6:.1f} vs "
          f"{h * w * 5 / min(res['1']) / 1e6:.1f} Gpx/s), identical results: {same}", flush=True)
os.environ.pop("OFD_BIL_TMA", None)
