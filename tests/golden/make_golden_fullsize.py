"""Full-size reference fixtures: run the REFERENCE's own Python (imported from /root/reference) at the
BASELINE sizes - 480x640 (cfg1/cfg5), 368x496 (cfg4), one ReDWeb-shaped frame (cfg2) and one 1080p frame (cfg3) - and commit

  * `fullsize_digests.json`: SHA-256 of every plane that must be BIT-EXACT on the GPU: utils.normalize_depth, the disparity flow,
    the FW prologue's safe_x / safe_y (alt_cuda/fw.py:27-43, unmodified), the splat's output / valid / collision / winner map
    (the reference's kernel loop restated literally, oracle.splat_literal, behind the reference's own FW.forward), the ConcatFlow /
    BackFlow results, and sparse_bilateral_filtering (bilateral_filter.py:13-60, unmodified numpy, 5 iterations [7,7,5,5,5]);
  * `fullsize_flows.npz`: the reference's 6-DoF flow (Convert.depth_to_random_flow, preprocess.py:265-298, torch on the CPU) on a
    strided pixel sample (float32) with its pose T1 - the tolerance planes.

Inputs come from the repo's seeded generators (opticalflowfromdepth_b200/synthetic.py, pure numpy), so the GPU box regenerates
them bit for bit and only digests travel.  Runs only in the build container (needs /root/reference); about 3 minutes, most of it
the reference's per-pixel Python bilateral loop.
"""
from __future__ import annotations

import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import make_golden as mg  # noqa: E402  (reference loader + stand-in fw_cuda)
import oracle  # noqa: E402
from opticalflowfromdepth_b200 import synthetic  # noqa: E402

# (tag, kind, generator index, H, W, flow sample stride)
CASES = [
    ("cfg1_480x640", "diml", 0, 480, 640, 2),
    ("cfg4_368x496", "diml", 200, 368, 496, 2),
    ("cfg2_redweb", "redweb", 3, None, None, 3),   # size = synthetic.redweb_sizes(1, seed=3)[0]
    ("cfg3_1080p", "diml", 100, 1080, 1920, 4),
]
FILTER = [7, 7, 5, 5, 5]


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def case_inputs(kind, idx, h, w):
    if kind == "redweb":
        (h, w), = synthetic.redweb_sizes(1, seed=3)
        img, raw = synthetic.redweb_frame(idx, h, w)
    else:
        img, raw = synthetic.diml_frame(idx, h, w)
    return img, raw, h, w


def main():
    torch.set_num_threads(8)
    pp, ref_utils, ref_geo, ref_bil, RefFW = mg.install_reference()
    fw = RefFW(device="cpu")
    digests, flows = {}, {}
    for tag, kind, idx, h, w, stride in CASES:
        t0 = time.time()
        img, raw, h, w = case_inputs(kind, idx, h, w)
        d = {"H": h, "W": w, "kind": kind, "index": idx, "seed": 12345 + idx}
        d["input_img"], d["input_raw_depth"] = sha(img), sha(raw)
        # utils.normalize_depth (utils.py:102-116), reference torch ops on the CPU
        depth = ref_utils.normalize_depth(torch.from_numpy(raw.copy()))
        d["normalize_depth"] = sha(depth.numpy())
        # Convert.depth_to_disparity + disparity_to_flow (preprocess.py:239-254), seeded like the driver (:555)
        ref_utils.set_seed(12345 + idx)
        disp = pp.Convert.depth_to_disparity(depth)
        flow01 = pp.Convert.disparity_to_flow(disp, device="cpu", random_sign=False)
        d["sBf"] = float(np.float32((disp * depth).max()))  # informative only; the test redraws the scale from the seed
        d["flow01"] = sha(flow01.numpy())
        # the first splat of the pipeline (preprocess.py:358-365) through the reference's own FW.forward
        obj = torch.cat((torch.from_numpy(img), depth, flow01 * -1.0), 0)
        out, valid, coll = fw(obj, flow01, depth)
        d["fw01_safe_x"], d["fw01_safe_y"] = sha(mg.CAPTURE["safe_x"].numpy()), sha(mg.CAPTURE["safe_y"].numpy())
        d["fw01_out"], d["fw01_valid"], d["fw01_collision"] = sha(out.numpy()), sha(valid.numpy()), sha(coll.numpy())
        _, _, _, win, _ = oracle.fw_forward(obj.numpy(), flow01.numpy(), depth.numpy())
        d["fw01_winner"] = sha(win.astype(np.int32))
        d["fw01_hit_fraction"] = float(valid.mean())
        # pair 0->1 post-ops (preprocess.py:361-365): mask, fix_warped_depth
        img1 = out[0:3] * valid
        depth1 = ref_utils.fix_warped_depth(out[3:4] * valid)
        back01 = out[4:6] * valid
        d["img1"], d["depth1"], d["back_flow01"] = sha(img1.numpy()), sha(depth1.numpy()), sha(back01.numpy())
        # 6-DoF flow of the warped view (preprocess.py:372 -> 265-298): the reference's torch geometry on the CPU
        flow12, T1 = pp.Convert.depth_to_random_flow(depth1, "cpu")
        flows[f"{tag}_T1"] = T1.numpy()
        flows[f"{tag}_flow12_sample"] = np.ascontiguousarray(flow12.numpy()[:, ::stride, ::stride])
        flows[f"{tag}_stride"] = np.array(stride)
        d["flow12_max_abs"] = float(flow12.abs().max())
        # given THAT flow, the C=7 splat + hole mask (preprocess.py:373-382) is exact arithmetic: digests of what the reference makes of
        # its own flow; the GPU test feeds the product the same flow (rebuilt from the reference's T1 by the oracle restatement, which
        # make_golden.py asserts bit-identical to the reference on the CPU) ... only if that restatement reproduces it here too:
        from oracle import flow as oflow
        same = bool(torch.equal(oflow.reproject_flow(depth1, T1), flow12))
        d["oracle_flow12_bit_identical_to_reference"] = same
        assert same, tag
        obj1 = torch.cat((img1, depth1, flow12 * -1.0, valid), 0)
        o2, v2, c2 = fw(obj1, flow12, depth1)
        valid2 = v2 * o2[6:7]
        d["fw12_out_masked"] = sha((o2[0:6] * valid2).numpy())
        d["fw12_valid2"], d["fw12_collision"] = sha(valid2.numpy()), sha(c2.numpy())
        d["fw12_hit_fraction"] = float(v2.mean())
        # ConcatFlow / BackFlow (preprocess.py:301-326) on the flows at hand
        cf, bf = pp.ConcatFlow("cpu"), pp.BackFlow("cpu")
        flow02, flow02_valid = cf(flow01, back01, flow12, depth1)
        d["concat_flow02"], d["concat_flow02_valid"] = sha(flow02.numpy()), sha(flow02_valid.numpy())
        back, back_valid = bf(flow12, depth1)
        d["backflow12"], d["backflow12_valid"] = sha(back.numpy()), sha(back_valid.numpy())
        # sparse_bilateral_filtering, the reference's numpy loop, unmodified
        tb = time.time()
        filt = ref_bil.sparse_bilateral_filtering(depth[0].numpy().copy(), np.zeros((h, w, 3), np.uint8), FILTER, depth_threshold=0.04,
                                                  num_iter=len(FILTER))
        d["bilateral_5iter"] = sha(np.asarray(filt, dtype=np.float32))
        d["bilateral_changed_fraction"] = float((np.asarray(filt) != depth[0].numpy()).mean())
        d["reference_bilateral_seconds"] = round(time.time() - tb, 1)
        digests[tag] = d
        print(f"{tag}: {h}x{w} done in {time.time() - t0:.1f} s (bilateral {d['reference_bilateral_seconds']} s)", flush=True)
    (HERE / "fullsize_digests.json").write_text(json.dumps(digests, indent=1, sort_keys=True))
    np.savez_compressed(HERE / "fullsize_flows.npz", **flows)
    print(f"fullsize_flows.npz: {(HERE / 'fullsize_flows.npz').stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
