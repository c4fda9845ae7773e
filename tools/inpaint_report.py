"""profiles/r2/inpaint_report.json: ofd_inpaint_telea (layer-ordered Telea fill on the device) against cv2.inpaint (OpenCV's serial heap
order, what the reference's utils.inpaint calls) on real pipeline masks - per case the hole fraction, mean |difference| in grey levels,
the fraction of hole bytes differing by more than one level, the maximum, and the time of both (device: CUDA events; cv2: host threads
as synthesis.inpaint(backend="cv2") runs it)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    H, W, B = 480, 640, 9
    rep = {"device": torch.cuda.get_device_name(0), "cases": {}}
    yy, xx = np.mgrid[0:H, 0:W]
    smooth = np.stack([(xx * 3 + yy) % 256, (xx + 2 * yy) % 256, 128 + 40 * np.sin(xx / 15.0) + 30 * np.cos(yy / 17.0)]).astype(np.uint8).astype(np.float32)
    frames = [synthetic.diml_frame(k, H, W) for k in range(B)]
    depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))
    for name, img in (("noise_image", torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)),
                      ("smooth_image", torch.from_numpy(np.repeat(smooth[None], B, 0)).to(dev))):
        sBf = torch.full((B,), 47.0, device=dev)
        Kc, invK = synthesis.Plausible.K((H, W))
        cams = []
        for k in range(B):
            torch.manual_seed(12345 + k)
            cams.append(geometry.camera_constants(Kc, invK, synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)[0]))
        cam = torch.cat(cams).to(dev)
        pair = synthesis.synthesize_pairs(img, depth, sBf)
        six = ops.reproject_pair(pair["img1"], pair["depth1"], cam, pair["valid"])
        kinds = [5 + k % 3 for k in range(B)]
        aug = ops.augment_pairs(img, depth, pair["img1"], pair["depth1"], pair["flow"], pair["back_flow"], kinds,
                                synthesis.sample_special_params(kinds, (H, W), torch.Generator().manual_seed(1)))
        for tag, (im, v, c) in {"stereo_pair": (pair["img1"], pair["valid"], pair["collision"]),
                                "sixdof_pair": (six[0], six[4], six[5]),
                                "augment_flip_rotate_shear": (aug["aug_img0"], aug["valid_img0"], aug["collision_img0"])}.items():
            mask = ops.inpaint_mask(v, c)
            ops.inpaint_telea(im, mask, 3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            got, (layers, filled) = ops.inpaint_telea(im, mask, 3, want_stats=True)
            e1.record()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = synthesis.inpaint(im, v, c, backend="cv2")
            t_cv = time.perf_counter() - t0
            hole = (mask != 0).expand_as(got)
            d = (got - ref).abs()
            assert bool((d[~hole] == 0).all())
            dh = d[hole]
            rep["cases"][f"{name}/{tag}"] = {
                "frames": B, "hole_fraction": float((mask != 0).float().mean()), "layers": layers, "filled_pixels": filled,
                "mean_abs_diff_levels": float(dh.mean()), "frac_hole_bytes_diff_gt_1": float((dh > 1).float().mean()),
                "frac_hole_bytes_diff_gt_8": float((dh > 8).float().mean()), "max_abs_diff": float(dh.max()),
                "device_ms_per_batch": e0.elapsed_time(e1), "cv2_host_ms_per_batch": 1e3 * t_cv, "cv2_host_threads": min(B, __import__("os").cpu_count())}
            print(name, tag, rep["cases"][f"{name}/{tag}"], flush=True)
    out = ROOT / "profiles" / "r2" / "inpaint_report.json"
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(rep, indent=1))
    print("wrote", out)


if __name__ == "__main__":
    main()
