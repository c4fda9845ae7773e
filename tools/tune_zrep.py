"""A/B of the hand-scheduled 6-DoF z-test (ztest_reproject_kernel) against the generic one (OFD_ZREP_GENERIC=1), CUDA events:
    python tools/tune_zrep.py            # runs both variants in child processes and prints one table
cases: cfg3 (32 x 1080x1920, C=7 + valid_in), one reprojection pair 128 x 480x640, the whole cfg5 group 128 x 480x640."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def child():
    import numpy as np
    import torch
    sys.path.insert(0, str(ROOT))
    from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic
    dev = torch.device("cuda:0")

    def frames(B, h, w, seed0, pool):
        fr = [synthetic.diml_frame(seed0 + k, h, w) for k in range(pool)]
        img = torch.from_numpy(np.stack([f[0] for f in fr])).to(dev)
        depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in fr])).to(dev))
        idx = torch.arange(B, device=dev) % pool
        return img[idx].contiguous(), depth[idx].contiguous()

    def cams(B, h, w):
        Kc, invK = synthesis.Plausible.K((h, w))
        out = []
        for k in range(B):
            torch.manual_seed(12345 + k)
            out.append(geometry.camera_constants(Kc, invK, synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)[0]))
        return torch.cat(out).to(dev)

    def timeit(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        tot = 0.0
        for _ in range(reps):
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = min(best, t)
            tot += t
        return best, tot / reps

    B, h, w = 32, 1080, 1920
    img, depth = frames(B, h, w, 100, 2)
    vin = torch.ones(B, 1, h, w, device=dev)
    cam = cams(B, h, w)
    b, m = timeit(lambda: ops.reproject_pair(img, depth, cam, vin))
    print(f"cfg3_1080p_b32_ms best {b:.4f} mean {m:.4f}  frac {64 * B * h * w / (m * 1e-3) / 6553.6e9:.3f}")
    del img, depth, vin
    B, h, w = 128, 480, 640
    img, depth = frames(B, h, w, 0, 16)
    cam = cams(B, h, w)
    b, m = timeit(lambda: ops.reproject_pair(img, depth, cam, None))
    print(f"pair_480x640_b128_ms best {b:.4f} mean {m:.4f}")
    sBf = torch.full((B,), 47.0, device=dev)
    b, m = timeit(lambda: synthesis.synthesize_group(img, depth, sBf, cam))
    print(f"group_480x640_b128_ms best {b:.4f} mean {m:.4f}  frac {356 * B * h * w / (m * 1e-3) / 6553.6e9:.3f}")
    b, m = timeit(lambda: ops.reproject_pair(img[:1], depth[:1], cam[:1], None), reps=200)
    print(f"pair_480x640_b1_ms best {b:.4f} mean {m:.4f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for gen in ("1", "0"):
            env = dict(os.environ, OFD_ZREP_GENERIC=gen)
            print(f"== OFD_ZREP_GENERIC={gen} ({'generic ztest_rows_kernel<ProdReproject>' if gen == '1' else 'ztest_reproject_kernel'})", flush=True)
            subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
        # compile-time variants built beforehand with _build.build_variant (opticalflowfromdepth_b200/build/variants/*.so)
        vdir = ROOT / "opticalflowfromdepth_b200" / "build" / "variants"
        for lib in sorted(vdir.glob("*.so")) if vdir.exists() else []:
            print(f"== variant {lib.name}", flush=True)
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, OFD_ZREP_GENERIC="0", OFD_LIB_PATH=str(lib)), check=False)
