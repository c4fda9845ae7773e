"""Randomised parity soak (not part of the test suite): random shapes / batches / flows through the fused pair kernel, the
general splat, the row-local splat, the ragged bilateral, the fused 6-DoF pair, the batched augmentation, the Telea fill and the
host-buffer pipeline, each checked against the CPU oracle
(bit-exact; the 6-DoF flow within the path's 1e-5 coordinate-relative tolerance).  `python tools/soak_parity.py SECONDS`."""
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


def trial_pair_ragged(rng):
    """Mixed-resolution batch through the one-launch ragged kernel (any H and W: units and planes on every 16-byte phase),
    float32 or float64 depth, each frame against the oracle."""
    n = int(rng.integers(1, 7))
    sizes = []
    for _ in range(n):
        w = int(rng.choice([rng.integers(1, 2400), rng.integers(1, 80), 2 * rng.integers(1, 700)]))
        q = 4 // np.gcd(w, 4) if rng.random() < 0.5 else 1  # half of the frames keep H*W a multiple of 4 (aligned specialisation)
        sizes.append((int(q * rng.integers(1, max(2, 40 // q))), w))
    f64 = rng.random() < 0.3
    imgs = [rng.integers(0, 256, (3, h, w)).astype(np.float32) for h, w in sizes]
    deps = [depth_field(rng, h, w, rng.random() < 0.5)[None] for h, w in sizes]
    if f64:
        deps = [d.astype(np.float64) + rng.uniform(0, 1e-3, d.shape) for d in deps]
    sBf = rng.uniform(40, 55, n).astype(np.float32)
    offs = [0]
    for h, w in sizes[:-1]:
        offs.append(offs[-1] + h * w)
    got = ops.disparity_pair_ragged(torch.cat([cu(i).reshape(-1) for i in imgs]), torch.cat([cu(d).reshape(-1) for d in deps]),
                                    cu(sBf), sizes, offs)
    views = [ops.ragged_views(t, c, sizes, offs) for t, c in zip(got, (3, 1, 2, 2, 1, 1))]
    ok = True
    for i in range(n):
        if f64:  # the oracle's pair is float32: compare with the single-frame float64 kernel path (itself pinned by the tests)
            want = [t[0].cpu().numpy() for t in ops.disparity_pair(cu(imgs[i])[None], cu(deps[i])[None], cu(sBf[i:i + 1]))]
        else:
            want = [t[0] for t in oracle.disparity_pair(imgs[i][None], deps[i][None], sBf[i:i + 1], nthreads=4)]
        ok = ok and all(eq(views[k][i], want[k]) for k in range(6))
    return ok, f"pair_ragged {sizes} f64={f64}"


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


def trial_bilateral_masked(rng):
    """The binary-mask path of the gated median (uint8 / float64 masks: float32 / float64 rank rule)."""
    h, w = int(rng.integers(3, 90)), int(rng.integers(3, 130))
    dt = np.float32 if rng.random() < 0.7 else np.float64
    d = depth_field(rng, h, w, rng.random() < 0.5).astype(dt)
    if rng.random() < 0.4:
        d[rng.integers(0, h), rng.integers(0, w)] = 0
    mask = (rng.random((h, w)) > rng.uniform(0.0, 0.6)).astype(np.uint8 if rng.random() < 0.5 else np.float64)
    fs = [int(rng.choice([3, 5, 7])) for _ in range(2)]
    got = bilateral_filter.sparse_bilateral_filtering(d.copy(), None, fs, depth_threshold=0.04, mask=mask, num_iter=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = obil.sparse_bilateral_filtering(d.copy(), fs, 0.04, 2, mask=mask)
    return bool(np.array_equal(got, want, equal_nan=True)), f"bilateral masked {h}x{w} {dt.__name__} mask {mask.dtype} windows {fs}"


def trial_reproject_pair(rng):
    """Fused 6-DoF pair: flow within the path's tolerance of the torch restatement, splat bit-exact given OUR flow."""
    from oracle import flow as oflow
    from opticalflowfromdepth_b200 import geometry, synthesis

    h, w, b = int(rng.integers(2, 70)), int(rng.integers(2, 160)), int(rng.integers(1, 4))
    img = rng.integers(0, 256, (b, 3, h, w)).astype(np.float32)
    depth = np.stack([depth_field(rng, h, w, rng.random() < 0.5) for _ in range(b)])[:, None]
    vin = (rng.random((b, 1, h, w)) > 0.1).astype(np.float32)
    K, invK = synthesis.Plausible.K((h, w))
    cams, poses = [], []
    for _ in range(b):
        torch.manual_seed(int(rng.integers(0, 1 << 30)))
        T1 = synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)[0]
        poses.append(T1)
        cams.append(geometry.camera_constants(K, invK, T1))
    io, do, bo, fo, vo, co, raw = ops.reproject_pair(cu(img), cu(depth), torch.cat(cams).to(DEV), cu(vin), want_raw_valid=True)
    ok = True
    for k in range(b):
        ref = oflow.reproject_flow(torch.from_numpy(depth[k]), poses[k]).numpy()
        got = fo[k].cpu().numpy()
        yy, xx = np.mgrid[0:h, 0:w]
        tol_x = 1e-5 * np.maximum(np.abs(ref[0] + xx), max(w - 1, 1))
        tol_y = 1e-5 * np.maximum(np.abs(ref[1] + yy), max(h - 1, 1))
        ok &= bool((np.abs(got[0] - ref[0]) <= tol_x).all() and (np.abs(got[1] - ref[1]) <= tol_y).all())
        obj = np.concatenate([img[k], depth[k], got * -1.0, vin[k]])
        o, v, cc, _, _ = oracle.fw_forward(obj, got, depth[k])
        v2 = v * o[6:7]
        ok &= eq(raw[k], v) and eq(vo[k], v2) and eq(co[k], cc) and eq(io[k], o[0:3] * v2) and eq(bo[k], o[4:6] * v2)
        ok &= eq(do[k], oflow.fix_warped_depth(torch.from_numpy(o[3:4] * v2)).numpy())
    return ok, f"reproject_pair {b}x{h}x{w}"


def trial_augment(rng):
    """ofd_augment_pairs: every one of the six splats against the oracle fed with the same special flow."""
    from oracle import flow as oflow
    from opticalflowfromdepth_b200 import synthesis

    h, w, b = int(rng.integers(4, 60)), int(rng.integers(4, 120)), int(rng.integers(1, 4))
    img = rng.integers(0, 256, (b, 3, h, w)).astype(np.float32)
    depth = np.stack([depth_field(rng, h, w, True) for _ in range(b)])[:, None]
    p01 = ops.disparity_pair(cu(img), cu(depth), cu(rng.uniform(40, 55, b).astype(np.float32)))
    img1, d1, back, flow = p01[0], p01[1], p01[2], p01[3]
    kinds = [int(rng.integers(5, 8)) for _ in range(b)]
    params = synthesis.sample_special_params(kinds, (h, w), torch.Generator().manual_seed(int(rng.integers(0, 1 << 30))))
    r = ops.augment_pairs(cu(img), cu(depth), img1, d1, flow, back, kinds, params)
    n = lambda t: t.cpu().numpy()  # noqa: E731
    ok = True
    for k in range(b):
        sf, bsf = n(r["special_flow"][k]), n(r["back_special_flow"][k])
        o, v, _, _, _ = oracle.fw_forward(n(flow[k]), sf, depth[k])
        ok &= eq(r["aug0_flow"][k], (o + bsf) * v)
        o, v, _, _, _ = oracle.fw_forward(sf, n(back[k]), n(d1[k]))
        ok &= eq(r["aug1_flow"][k], (o + n(flow[k])) * v)
        for tag, im, dp in (("0", img[k], depth[k]), ("1", n(img1[k]), n(d1[k]))):
            o, v, cc, _, _ = oracle.fw_forward(np.concatenate([im, dp]), sf, dp)
            ok &= eq(r["aug_img" + tag][k], o[0:3]) and eq(r["valid_img" + tag][k], v) and eq(r["collision_img" + tag][k], cc)
            ok &= eq(r["aug_depth" + tag][k], oflow.fix_warped_depth(torch.from_numpy(o[3:4].copy())).numpy())
        a0 = n(r["aug0_flow"][k])
        o, v, _, _, _ = oracle.fw_forward(a0, a0, n(r["aug_depth0"][k]))
        ok &= eq(r["back_aug0_flow"][k], (o * -1.0) * v)
        a1 = n(r["aug1_flow"][k])
        o, v, _, _, _ = oracle.fw_forward(a1, a1, depth[k])
        ok &= eq(r["back_aug1_flow"][k], (o * -1.0) * v)
    return ok, f"augment {b}x{h}x{w} kinds {kinds}"


def trial_splat_rows(rng):
    """Row-local splat for horizontal warp flows against the oracle's FW.forward (ConcatFlow / BackFlow / plain), C = 2."""
    h, w, b = int(rng.integers(1, 60)), int(rng.choice([rng.integers(1, 2049), rng.integers(1, 70)])), int(rng.integers(1, 4))
    obj = rng.normal(0, 30, (b, 2, h, w)).astype(np.float32)
    aux = rng.normal(0, 30, (b, 2, h, w)).astype(np.float32)
    fx = rng.normal(0, rng.choice([2.0, 40.0]), (b, 1, h, w)).astype(np.float32)
    if rng.random() < 0.5:
        fx[rng.random(fx.shape) < 0.02] = np.nan
    flow = np.concatenate([fx, np.where(rng.random(fx.shape) < 0.5, np.float32(-0.0), np.float32(0.0)).astype(np.float32)], 1)
    depth = np.stack([depth_field(rng, h, w, rng.random() < 0.7) for _ in range(b)])[:, None]
    if rng.random() < 0.4:
        depth[rng.random(depth.shape) < 0.02] = 1000.0
    epi = int(rng.integers(0, 3))
    got = ops.splat_flow(cu(obj), cu(flow), cu(depth), epilogue=epi, aux=cu(aux) if epi == ops.EPI_CONCAT else None, horizontal=True)
    ok = True
    for k in range(b):
        o, v, c, _, _ = oracle.fw_forward(obj[k], flow[k], depth[k])
        want = (o + aux[k]) * v if epi == ops.EPI_CONCAT else ((o * -1.0) * v if epi == ops.EPI_BACK else o)
        ok &= eq(got[0][k], want) and eq(got[1][k], v) and eq(got[2][k], c)
    return ok, f"splat_rows {b}x{h}x{w} epilogue {epi}"


def trial_telea(rng):
    """Telea fill kernel against the layer-order restatement (small frames: the restatement is pure Python)."""
    from oracle import inpaint as oinp

    h, w, b = int(rng.integers(2, 28)), int(rng.integers(2, 36)), int(rng.integers(1, 4))
    imgs = rng.integers(0, 256, (b, h, w, 3)).astype(np.uint8)
    masks = (rng.random((b, h, w)) < rng.choice([0.05, 0.3, 0.6])).astype(np.uint8)
    for k in range(b):
        if rng.random() < 0.5:
            r0, c0 = rng.integers(0, h), rng.integers(0, w)
            masks[k, r0:r0 + rng.integers(1, h + 1), c0:c0 + rng.integers(1, w + 1)] = 1
    radius = int(rng.choice([1, 2, 3, 3, 3, 4]))
    got = ops.inpaint_telea(cu(np.ascontiguousarray(imgs.transpose(0, 3, 1, 2)).astype(np.float32)), cu(masks[:, None]), radius)
    ok = True
    for k in range(b):
        want = oinp.telea(imgs[k], masks[k], radius, order="layer").transpose(2, 0, 1).astype(np.float32)
        ok &= eq(got[k], want)
    return ok, f"telea {b}x{h}x{w} radius {radius}"


_pipes = {}


def trial_pipeline(rng):
    """Host-buffer pipeline (byte masks, verified image bytes with float fallback) against the oracle."""
    h, w = int(rng.choice([8, 12, 20])), int(rng.choice([16, 24, 40]))
    b = int(rng.integers(1, 9))
    img = rng.integers(0, 256, (b, 3, h, w)).astype(np.float32)
    if rng.random() < 0.3:
        img[rng.integers(0, b)] += np.float32(rng.choice([0.5, 300.0, -2.0]))  # one frame a byte cannot carry
    depth = np.stack([depth_field(rng, h, w, True) for _ in range(b)])[:, None]
    sBf = rng.uniform(40, 55, b).astype(np.float32)
    key = (h, w, int(rng.integers(0, 2)))
    if key not in _pipes:
        _pipes[key] = ops.PairPipeline(0, h, w, chunk_frames=int(rng.integers(1, 4)))
    outs = [torch.full((b, c, h, w), 7.0).pin_memory() for c in (3, 1, 2, 2, 1, 1)]
    _pipes[key].run(torch.from_numpy(img).pin_memory(), torch.from_numpy(depth).pin_memory(), torch.from_numpy(sBf), *outs)
    want = oracle.disparity_pair(img, depth, sBf, nthreads=2)
    return all(np.array_equal(o.numpy(), wv) for o, wv in zip(outs, want)), f"pipeline {b}x{h}x{w}"


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0, counts, fails = time.time(), {}, []
    trials = (trial_pair, trial_pair_ragged, trial_splat, trial_bilateral, trial_bilateral_masked, trial_reproject_pair, trial_augment,
              trial_splat_rows, trial_telea, trial_pipeline)
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
