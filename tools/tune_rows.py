import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
def child():
    import numpy as np, torch
    sys.path.insert(0, str(ROOT))
    from opticalflowfromdepth_b200 import ops, synthesis, synthetic
    dev = torch.device("cuda:0")
    for (B, h, w) in ((128, 480, 640), (64, 368, 496)):
        fr = [synthetic.diml_frame(k, h, w) for k in range(16)]
        img = torch.from_numpy(np.stack([f[0] for f in fr])).to(dev)
        depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in fr])).to(dev))
        idx = torch.arange(B, device=dev) % 16
        img, depth = img[idx].contiguous(), depth[idx].contiguous()
        pair = synthesis.synthesize_pairs(img, depth, torch.full((B,), 47.0, device=dev))
        fl = torch.randn(B, 2, h, w, device=dev)
        def f():
            ops.splat_flow(fl, pair["back_flow"], pair["depth1"], epilogue=ops.EPI_CONCAT, aux=pair["flow"], want_collision=False, horizontal=True)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(20):
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        m = sum(ts) / len(ts)
        print(f"rows concat {B}x{h}x{w}: best {min(ts):.4f} mean {m:.4f} ms  {40 * B * h * w / (m * 1e-3) / 1e9:.0f} GB/s on 40 B/px")
if len(sys.argv) > 1: child()
else:
    # the row-local splat: four pixels per thread (default when the planes are 16-byte aligned and W % 4 == 0) against the scalar kernel
    for scalar in ("1", "0"):
        env = dict(os.environ, OFD_ROWS_SCALAR=scalar)
        print("== OFD_ROWS_SCALAR=" + scalar + (" (one pixel per thread)" if scalar == "1" else " (four pixels per thread)"), flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
