"""Quick tour of the B200 flow-synthesis path on synthetic frames (needs a B200 and the built library:
`python -c "import __graft_entry__ as g; g.build()"`).   python examples/quickstart.py"""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import bilateral_filter, dataloader, geometry, ops, synthesis, synthetic  # noqa: E402
from opticalflowfromdepth_b200 import preprocess as pp  # noqa: E402
from opticalflowfromdepth_b200.fw import FW  # noqa: E402

dev = torch.device("cuda:0")
H, W, B = 240, 320, 4
frames = [synthetic.diml_frame(k, H, W) for k in range(B)]
img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)                  # [B,3,H,W] float32, 0..255
depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))  # utils.normalize_depth, [1,99] u {100}

# 1. the reference's own operator: FW(device).forward(obj, flow, depth)  (alt_cuda/fw.py)
synthesis.set_seed(12345)
flow01 = synthesis.Convert.depth_to_disparity_flow(depth[0])                      # virtual-stereo flow of frame 0
warped, valid, collision = FW(dev)(torch.cat((img[0], depth[0])), flow01, depth[0])
print("FW.forward:", tuple(warped.shape), "hit rate", float(valid.mean()))

# 2. one fused kernel per batch of virtual-stereo pairs (preprocess.py:356-365)
pair = synthesis.synthesize_pairs(img, depth, torch.full((B,), 47.0, device=dev))
print("pairs:", {k: tuple(v.shape) for k, v in pair.items()})

# 3. the reference's 5-pair group of every frame (preprocess.py:356-432): 9 launches for the batch
K, inv_K = synthesis.Plausible.K((H, W))
cams = []
for k in range(B):
    synthesis.set_seed(12345 + k)
    cams.append(geometry.camera_constants(K, inv_K, synthesis.Plausible.random_motion(1 / 36, 1 / 36, 0.1, 0.1)[0]))
group = synthesis.synthesize_group(img, depth, torch.full((B,), 47.0, device=dev), torch.cat(cams).to(dev))
print("group:", len(group), "tensors, flow02", tuple(group["flow02"].shape))

# 3b. one pair of that group by hand: pair 0->2' = ConcatFlow of flow12 along the horizontal back_flow01 fused with the frame splat of
#     (img0, depth0) along the result (preprocess.py:400-411) - two launches, same tensors as the group's
f02, f02_valid, img2p, depth2p, back02p, valid2p, _ = ops.concat_frame_splat(group["flow12"], group["back_flow01"], group["depth1"],
                                                                               group["flow01"], img, depth)
assert torch.equal(f02, group["flow02"]) and torch.equal(img2p, group["img2_prime"])
print("pair 0->2':", tuple(f02.shape), "valid fraction", float(valid2p.mean()))

# 4. in-loop geometric augmentation of a batch of pairs (one native call, 13 launches)
s1, s2, (sf, bsf) = synthesis.augment_flow_batch(img, depth, pair["img1"], pair["depth1"], pair["flow"], pair["back_flow"],
                                                 kinds=[5, 6, 7, 6], reference_draws=False)
print("augmented:", tuple(s1[0].shape), tuple(s1[2].shape))

# 5. depth-edge gated-median "bilateral" smoothing, single image and ragged batch
smooth = bilateral_filter.sparse_bilateral_filtering(depth[0, 0], None, [7, 7, 5, 5, 5], depth_threshold=0.04, num_iter=5)
ragged = bilateral_filter.sparse_bilateral_filtering_batch([depth[0, 0], depth[1, 0, :100, :150].contiguous()], [7, 5], 0.04, 2)
print("bilateral:", tuple(smooth.shape), [tuple(r.shape) for r in ragged])

# 5b. mixed-resolution frames end to end: ragged normalize + bilateral, then every frame's virtual-stereo pair in ONE launch
r_depths = [depth[0, 0], depth[1, 0, :100, :150].contiguous()]
packed_depth, shapes, offsets = bilateral_filter.sparse_bilateral_filtering_batch(r_depths, [7, 5], 0.04, 2, return_packed=True)
packed_img = torch.cat([img[0].reshape(-1), img[1, :, :100, :150].reshape(-1)])  # frame i as [3,H_i,W_i] at element 3 * offsets[i]
packed = ops.disparity_pair_ragged(packed_img, packed_depth, torch.full((2,), 47.0, device=dev), shapes, offsets)
print("ragged pairs:", [tuple(v.shape) for v in ops.ragged_views(packed[0], 3, shapes, offsets)])

# 5c. utils.inpaint with the Telea fill on the device (cv2 on the host stays available: backend="cv2")
filled = synthesis.inpaint(pair["img1"], pair["valid"], pair["collision"], backend="cuda")
print("inpainted:", tuple(filled.shape), "holes filled:", int((pair["valid"] == 0).sum()))

# 5d. host frames in, the reference's 44-channel group arrays out (page-locked host memory), sharded by image index
from opticalflowfromdepth_b200 import sweep, synthetic  # noqa: E402

got = []
sink = sweep.PinnedGroupSink(lambda idx, arr, release: (got.append((idx, arr.shape)), release()))
sweep.run_sweep(range(4), lambda i: synthetic.diml_frame(i, H, W), dev, batch=2, dataset_len=4, sink=sink)
print("sweep:", got)

# 6. the reference's driver: 121 .npz files per frame, then read one sample back like the training loader does
with tempfile.TemporaryDirectory() as tmp:
    driver = pp.PreprocessPlusAugment(dev, inpaint="cuda", quiet=True, compress=1, reader_compat=True)
    synthesis.set_seed(12345)
    driver(pp.SyntheticDataset(1, H, W)[0], f"{tmp}/0", is_stereo=False)
    driver.close()
    print("driver wrote", len(list(Path(tmp, "0").glob("*.npz"))), "files")
    np.random.seed(0)
    img0, img1, flow, d0, label = dataloader.AugmentedFolder(tmp, 1, crop_size=(128, 160))[0]
    print("sample:", tuple(img0.shape), tuple(flow.shape), label.tolist())
# 7. ... or skip the files: synthesise the trainers' samples inside the training loop (BASELINE config 4)
from opticalflowfromdepth_b200 import inloop  # noqa: E402

raw = torch.from_numpy(np.stack([synthetic.diml_frame(k, H, W)[1] for k in range(B)])).to(dev)  # un-normalised depth
i1, i2, fl, bfl, d1, d2, valid, back_valid, label = inloop.InLoopSampler(dev, seed=0)(img, raw).raft_tuple()
print("in-loop batch:", tuple(i1.shape), tuple(fl.shape), label.argmax(1).tolist())
torch.cuda.synchronize()
print("ok")
