"""How does ofd_inpaint_telea scale with the batch?  The driver fills 9 augmented images per call (10 calls per frame); the kernel is bound
by its per-layer grid barrier at that size, so bigger batches should be cheaper per image.  Prints ms per call and per image for batches
of 9 / 18 / 45 / 90 frames 480x640 with the masks of the geometric augmentations (flip / rotate / shear) and of the 6-DoF pairs."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    H, W, B = 480, 640, 9
    frames = [synthetic.diml_frame(k, H, W) for k in range(B)]
    img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)
    depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))
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
    for tag, (im, v, c) in {"augment": (aug["aug_img0"], aug["valid_img0"], aug["collision_img0"]), "sixdof": (six[0], six[4], six[5])}.items():
        mask = ops.inpaint_mask(v, c)
        for rep in (1, 2, 5, 10):
            im_r, mask_r = im.repeat(rep, 1, 1, 1).contiguous(), mask.repeat(rep, 1, 1, 1).contiguous()
            for _ in range(2):
                ops.inpaint_telea(im_r, mask_r, 3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.inpaint_telea(im_r, mask_r, 3)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{tag:8s} batch {B * rep:3d}: {ms:7.3f} ms per call, {ms / (B * rep):6.3f} ms per image", flush=True)


if __name__ == "__main__":
    main()
