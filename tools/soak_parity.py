"""Randomised parity soak (not part of the test suite): random shapes / batches / flows through the fused pair kernel, the
general splat and the ragged bilateral, each checked bit-exactly against the CPU oracle.  `python tools/soak_parity.py SECONDS`."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from oracle import bilateral as obil  # noqa: E402
from opticalflowfromdepth_b200 import bilateral_filter, ops  # noqa: E402

DEV = torch.device("cuda:0")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def eq(t, a):
    return np.array_equal(t.cpu().numpy(), a, equal_nan=True)


def depth_field(rng, h, w, quant):
    y, x = np.mgrid[0:h, 0:w]
    d = 20 + 15 * y / max(h, 1) + 3 * np.sin(2 * np.pi * x / max(w, 1))
    for _ in range(rng.integers(0, 6)):
        r0, c0 = rng.integers(0, h), rng.integers(0, w)
        d[r0:r0 + rng.integers(1, h + 1), c0:c0 + rng.integers(1, w + 1)] = rng.uniform(1, 99)
    d = d + rng.normal(0, 0.5, d.shape)
    if quant:
        d = np.round(d)
    return np.clip(d, 1, 99).astype(np.float32)


def trial_pair(rng):
    h, w, b = int(rng.integers(1, 90)), int(rng.choice([rng.integers(1, 700), rng.integers(1, 60), 4 * rng.integers(1, 300)])), int(rng.integers(1, 5))
    img = rng.integers(0, 256, (b, 3, h, w)).astype(np.float32)
    depth = np.stack([depth_field(rng, h, w, rng.random() < 0.5) for _ in range(b)])[:, None]
    if rng.random() < 0.3:
        depth[0, 0, rng.integers(0, h), rng.integers(0, w)] = 1000.0
    sBf = rng.uniform(40, 55, b).astype(np.float32)
    got = ops.disparity_pair(cu(img), cu(depth), cu(sBf))
    want = oracle.disparity_pair(img, depth, sBf, nthreads=4)
    ok = all(eq(g, wv) for g, wv in zip(got, want))
    return ok, f"pair {b}x{h}x{w}"


def trial_splat(rng):
    h, w, b, c = int(rng.integers(1, 70)), int(rng.integers(1, 200)), int(rng.integers(1, 4)), int(rng.integers(1, 9))
    obj = rng.normal(0, 50, (b, c, h, w)).astype(np.float32)
    flow = rng.normal(0, rng.uniform(0.3, 40), (b, 2, h, w)).astype(np.float32)
    depth = np.stack([depth_field(rng, h, w, rng.random() < 0.7) for _ in range(b)])[:, None]
    out, valid, coll, win = ops.splat_flow(cu(obj), cu(flow), cu(depth), want_winner=True)
    ok = True
    for k in range(b):
        o, v, cc, wm, _ = oracle.fw_forward(obj[k], flow[k], depth[k])
        ok &= eq(out[k], o) and eq(valid[k], v) and eq(coll[k], cc) and eq(win[k, 0], wm)
    return ok, f"splat {b}x{c}x{h}x{w}"


def trial_bilateral(rng):
    n = int(rng.integers(1, 5))
    depths = []
    for _ in range(n):
        h, w = int(rng.integers(3, 80)), int(rng.integers(3, 120))
        d = depth_field(rng, h, w, rng.random() < 0.5)
        if rng.random() < 0.3:
            d[rng.integers(0, h), rng.integers(0, w)] = 0
        depths.append(d)
    fs = [int(rng.choice([3, 5, 7, 9])) for _ in range(3)]
    got = bilateral_filter.sparse_bilateral_filtering_batch([cu(d) for d in depths], fs, 0.04, 3)
    ok = True
    with np.errstate(divide="ignore", invalid="ignore"):
        for d, g in zip(depths, got):
            ok &= eq(g, obil.sparse_bilateral_filtering(d.copy(), fs, 0.04, 3))
    return ok, f"bilateral {[d.shape for d in depths]} windows {fs}"


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0, counts, fails = time.time(), {}, []
    trials = (trial_pair, trial_splat, trial_bilateral)
    k = 0
    while time.time() - t0 < budget:
        fn = trials[k % len(trials)]
        k += 1
        ok, desc = fn(rng)
        counts[fn.__name__] = counts.get(fn.__name__, 0) + 1
        if not ok:
            fails.append(desc)
            print("MISMATCH:", desc, flush=True)
    print(f"soak: {counts} trials in {time.time() - t0:.0f} s, {len(fails)} mismatches")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
