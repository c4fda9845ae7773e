"""In-loop synthesis of training samples on the GPU (BASELINE config 4: "in-loop dataloader augmentation for RAFT training:
geometric augmentation on one image of the pair, fused with flow synthesis").

The reference pre-bakes every frame into 121 files (group.npz + 5 pairs x 12 augmentations x 2 sets, preprocess.py:427-476) and the
trainers read ONE of them back per sample: `AugmentedDataset.getitem_from_npz` picks a pair group 0..2, an augmentation slot 0..11
and a set 1..2 (dataloader.py:60-157, 235-268), `RAFTAugmentedDataset.__getitem__` / `GMFlowAugmentedDataset.__getitem__` turn the
result into `(img1, img2, flow, back_flow, img1_depth, img2_depth, valid, back_valid, label)` (adjusted_RAFT/core/datasets.py:249-292,
adjusted_gmflow/data/datasets.py:323-358).  Here the same sample is SYNTHESISED per training batch from the raw RGB-D frames:

    (img[B,3,H,W], depth[B,1,H,W])  ->  5-pair group (synthesis.synthesize_group: 7 splats)
                                    ->  per sample: pair group g, augmentation slot a (type AUGMENT_TYPES[a]), augmented image `which`
                                    ->  geometric types 5-7: synthesis.augment_flow_batch (one native call, 6 splats per sample),
                                        photometric types 0-2 and the placement of every sample's 12 planes: two tables of plane
                                        operations, one launch each (ofd_plane_ops)
                                    ->  the trainers' 9-tuple, as CUDA tensors

No file is written or read, nothing crosses PCIe but the input frames, and a sample is one of 5 x 12 x 2 variants of a FRESH random
stereo baseline / camera pose every time it is drawn instead of one of 120 frozen files.

Equivalence: for the same host draws (`Plan`) sample b equals what `preprocess.PreprocessPlusAugment` writes into
`{g}_{a}_{which+1}.npz` for that frame, bit for bit (tests/test_gpu_parity.py::test_inloop_sampler_equals_the_prebaked_files).
The draws themselves follow the reference's DISTRIBUTIONS (utils.get_random ranges, preprocess.py:24-105,150-163,194-235), not its
global-RNG call sequence: an in-loop sampler has its own torch.Generator.

Two deliberate differences from the reference's reader, both reference defects: the second image of a sample is
the pair's own second image (the reader takes `img2` where pair 2 is `img0 -> img2_prime`, dataloader.py:100-104, and pairs an
augmented first image with the unaugmented group image); and the tuple has the seven members the trainers unpack
(`AugmentedDataset` returns five, adjusted_RAFT/core/datasets.py:260).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops, synthesis
from .preprocess import AUGMENT_TYPES, GROUP_PAIRS

NUM_CLASSES = 1 + 3  # dataloader.py:11
_PAIR_CHANNELS = (3, 1, 3, 1, 2, 2)  # imgA, depthA, imgB, depthB, flowAB, back_flowAB (preprocess.GROUP_PAIRS order)


@dataclass
class Plan:
    """Host-side random decisions of one batch.  Everything a sample depends on besides its frame."""
    sBf: torch.Tensor                 # [B] float32 (CPU): s * B * f of the virtual stereo pair (preprocess.py:239-246)
    cam: torch.Tensor                 # [B,21] float32 (CPU): inv_K3 | (K T)[:3] of the random camera motion (preprocess.py:265-298)
    group: List[int]                  # pair group per sample, index into preprocess.GROUP_PAIRS (the trainers use 0..2)
    slot: List[int]                   # augmentation slot 0..11 -> type AUGMENT_TYPES[slot]
    which: List[int]                  # 0: the FIRST image of the pair is augmented (file *_1.npz), 1: the second (*_2.npz)
    draws: list = field(default_factory=list)  # per sample: 10 floats / None (types 5-7, SpecialFlow.params) or photometric draws

    @property
    def types(self) -> List[int]:
        return [AUGMENT_TYPES[s] for s in self.slot]


def _photometric_draws(t: int, gen: torch.Generator):
    """photometric_draws (preprocess.py:150-163) from a private generator: type 0 brightness scale U[0,1); type 1 one channel
    shifted by +-[15,25); type 2 grayscale (no draw)."""
    if t >= 2:
        return None
    if t >= 1:
        channel = int(torch.rand(1, generator=gen).item() * 3)
        sign = int(torch.randint(0, 2, (1,), generator=gen).item()) * 2 - 1
        shift = torch.tensor(sign) * (torch.rand(1, generator=gen)[0] * 10 + torch.tensor(15))
        return channel, shift
    return torch.rand(1, generator=gen)[0] * 1 + torch.tensor(0)


class InLoopSampler:
    """frames -> training samples, per batch, on the GPU.

    sampler = InLoopSampler("cuda:0", seed=0)
    batch = sampler(img, depth)                  # img[B,3,H,W] float32 uint8-valued, depth[B,1,H,W] raw (un-normalised) depth, CUDA
    img1, img2, flow, back_flow, d1, d2, valid, back_valid, label = batch.raft_tuple()

    inpaint: None (holes stay 0, as with `--no_inpaint`), "cuda" (ofd_inpaint_telea) or any callable (img, valid, collision) -> img.
    groups: the pair groups to draw from (the reference's reader uses 0..2: 0->1, 1->2, 0->2')."""

    def __init__(self, device, seed: Optional[int] = None, inpaint=None, groups: Sequence[int] = (0, 1, 2), normalized: bool = False):
        self.device = torch.device(device)
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(int(seed))
        self.inpaint = synthesis.inpaint_cuda if inpaint == "cuda" else inpaint
        self.groups = tuple(int(g) for g in groups)
        self.normalized = normalized

    # ---- host draws ---------------------------------------------------------------------------------------------------
    def draw(self, B: int, size) -> Plan:
        h, w = size
        gen = self.gen
        # one stereo scale + one 6-DoF pose per sample (Convert.disparity_scale, Plausible.random_motion(1/36, 1/36, 0.1, 0.1))
        u = torch.rand((B, 7), generator=gen)
        sg = torch.randint(0, 2, (B, 6), generator=gen) * 2 - 1
        sBf, cam, _ = synthesis.frame_params_from_uniforms(u, sg, (h, w))
        group = [self.groups[i] for i in torch.randint(0, len(self.groups), (B,), generator=gen).tolist()]
        slot = torch.randint(0, len(AUGMENT_TYPES), (B,), generator=gen).tolist()
        which = torch.randint(0, 2, (B,), generator=gen).tolist()
        types = [AUGMENT_TYPES[s] for s in slot]
        geo = [b for b, t in enumerate(types) if t >= 5]
        special = synthesis.sample_special_params([types[b] for b in geo], (h, w), gen) if geo else []
        draws: list = [None] * B
        for b, p in zip(geo, special):
            draws[b] = p
        for b, t in enumerate(types):
            if t < 5:
                draws[b] = _photometric_draws(t, gen)
        return Plan(sBf=sBf, cam=cam, group=group, slot=slot, which=which, draws=draws)

    # ---- device work --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, img: torch.Tensor, depth: torch.Tensor, plan: Optional[Plan] = None) -> "Batch":
        dev = self.device
        img = img.to(dev).float().contiguous()
        depth = depth.to(dev).float().contiguous()
        B, _, h, w = img.shape
        hw = h * w
        if plan is None:
            plan = self.draw(B, (h, w))
        types = plan.types
        with torch.cuda.device(dev):
            depth0 = depth if self.normalized else ops.normalize_depth(depth)
            # host -> device through page-locked staging, asynchronously: nothing in a step synchronises the stream
            up = lambda t: t.pin_memory().to(dev, non_blocking=True)  # noqa: E731
            grp = synthesis.synthesize_group(img, depth0, up(plan.sBf), up(plan.cam), inpaint=self.inpaint)
            # Every sample is 12 planes (imgA 3, depthA 1, imgB 3, depthB 1, flowAB 2, back_flowAB 2) picked from the group's tensors - or,
            # for a geometric augmentation, from what ofd_augment_pairs makes of them.  Picking, the photometric functions and the final
            # placement are plane operations: two tables, two launches (ofd_plane_ops), whatever the batch draws.
            f32 = {}

            def plane(name, b, ch):  # device address of channel `ch` of sample `b` of a group tensor
                t = f32.get(name)
                if t is None:
                    t = f32[name] = grp[name].float().contiguous()
                return t.data_ptr() + 4 * hw * (b * t.shape[1] + ch)

            out = [torch.empty((B, c, h, w), dtype=torch.float32, device=dev) for c in _PAIR_CHANNELS]
            first_img, first_dep, second_img, second_dep, flow, back = out
            rows = []  # (src, dst, op, p)
            geo = [b for b, t in enumerate(types) if t >= 5]
            keep = [f32]  # everything a table points into stays referenced until the launches are issued
            if geo:
                sub = [torch.empty((len(geo), c, h, w), dtype=torch.float32, device=dev) for c in _PAIR_CHANNELS]
                for j, b in enumerate(geo):
                    for m, name in enumerate(GROUP_PAIRS[plan.group[b]]):
                        for ch in range(_PAIR_CHANNELS[m]):
                            rows.append((plane(name, b, ch), sub[m].data_ptr() + 4 * hw * (j * _PAIR_CHANNELS[m] + ch), _lib.PLANE_COPY, 0.0))
                ops.plane_ops(np.array(rows, dtype=ops.PLANE_OP_DTYPE), hw, dev)
                rows = []
                set1, set2, _ = synthesis.augment_flow_batch(*sub, kinds=[types[b] for b in geo], inpaint=self.inpaint,
                                                             params=[plan.draws[b] for b in geo])
                # set1 = (aug_imgA, aug_depthA, augA_flow, back_augA_flow, imgB, depthB); set2 = (imgA, depthA, aug1_flow, back_aug1_flow, aug_imgB, aug_depthB)
                sets = [[t.float().contiguous() for t in s] for s in (set1, set2)]
                keep += [sub, sets]
                for j, b in enumerate(geo):
                    src = sets[plan.which[b]]
                    for m, k in enumerate((0, 1, 4, 5, 2, 3)):  # out order: imgA, depthA, imgB, depthB, flow, back_flow
                        for ch in range(_PAIR_CHANNELS[m]):
                            off = 4 * hw * (j * _PAIR_CHANNELS[m] + ch)
                            rows.append((src[k].data_ptr() + off, out[m].data_ptr() + 4 * hw * (b * _PAIR_CHANNELS[m] + ch), _lib.PLANE_COPY, 0.0))
            for b, t in enumerate(types):
                if t >= 5:
                    continue
                # photometric augmentation (types 0-2, preprocess.py:150-182) of the image `which` points at: brightness scale /
                # one-channel shift / grayscale (synthesis.photometric_apply's expressions)
                for m, name in enumerate(GROUP_PAIRS[plan.group[b]]):
                    for ch in range(_PAIR_CHANNELS[m]):
                        op, p, src = _lib.PLANE_COPY, 0.0, plane(name, b, ch)
                        if m == (0, 2)[plan.which[b]]:
                            if t == 0:
                                op, p = _lib.PLANE_SCALE, float(plan.draws[b])
                            elif t == 1 and ch == int(plan.draws[b][0]):
                                op, p = _lib.PLANE_ADD, float(plan.draws[b][1])
                            elif t == 2:
                                op, src = _lib.PLANE_GRAY, plane(name, b, 0)
                        rows.append((src, out[m].data_ptr() + 4 * hw * (b * _PAIR_CHANNELS[m] + ch), op, p))
            ops.plane_ops(np.array(rows, dtype=ops.PLANE_OP_DTYPE), hw, dev)
            del keep
            label_host = np.zeros((B, NUM_CLASSES), np.float32)
            label_host[np.arange(B), [max(0, t - 4) for t in types]] = 1.0  # dataloader.py:153-156
            label = up(torch.from_numpy(label_host))
        return Batch(first_img, second_img, flow, back, first_dep, second_dep, label, plan)


@dataclass
class Batch:
    img1: torch.Tensor        # [B,3,H,W] first image of the pair (augmented when plan.which == 0)
    img2: torch.Tensor        # [B,3,H,W] second image (augmented when plan.which == 1)
    flow: torch.Tensor        # [B,2,H,W] flow img1 -> img2
    back_flow: torch.Tensor   # [B,2,H,W]
    img1_depth: torch.Tensor  # [B,1,H,W]
    img2_depth: torch.Tensor  # [B,1,H,W]
    label: torch.Tensor       # [B,4] one-hot of max(0, type - 4)
    plan: Plan

    def file_arrays(self, b: int) -> torch.Tensor:
        """The 8-channel array preprocess.py stores for this sample (`img_depth_flow` of {g}_{a}_{which+1}.npz, :459-476)."""
        if self.plan.which[b] == 0:
            return torch.cat((self.img1[b], self.img1_depth[b], self.flow[b], self.back_flow[b]), 0)
        return torch.cat((self.flow[b], self.back_flow[b], self.img2[b], self.img2_depth[b]), 0)

    def raft_tuple(self):
        """RAFTAugmentedDataset.__getitem__ / GMFlowAugmentedDataset.__getitem__ without the trainers' own FlowAugmentor
        (adjusted_RAFT/core/datasets.py:281-288): valid = |flow| < 1000 on both components and depth != 100."""
        valid = (self.flow[:, 0].abs() < 1000) & (self.flow[:, 1].abs() < 1000) & (self.img1_depth[:, 0] != 100)
        back_valid = (self.back_flow[:, 0].abs() < 1000) & (self.back_flow[:, 1].abs() < 1000) & (self.img2_depth[:, 0] != 100)
        return (self.img1, self.img2, self.flow, self.back_flow, self.img1_depth, self.img2_depth, valid.float(), back_valid.float(),
                self.label)

    def reader_tuple(self):
        """The five members dataloader.AugmentedDataset returns (dataloader.py:157): img0, img1, flow, img0_depth, label."""
        return self.img1, self.img2, self.flow, self.img1_depth, self.label

