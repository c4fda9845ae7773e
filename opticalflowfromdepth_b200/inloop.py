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
                                        photometric types 0-2: three elementwise expressions
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

import torch

from . import ops, synthesis
from .preprocess import AUGMENT_TYPES, GROUP_PAIRS

NUM_CLASSES = 1 + 3  # dataloader.py:11


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
        if plan is None:
            plan = self.draw(B, (h, w))
        with torch.cuda.device(dev):
            depth0 = depth if self.normalized else ops.normalize_depth(depth)
            grp = synthesis.synthesize_group(img, depth0, plan.sBf.to(dev), plan.cam.to(dev), inpaint=self.inpaint)
            # the pair of every sample: (imgA, depthA, imgB, depthB, flowAB, back_flowAB) of its group
            used = sorted(set(plan.group))
            rows = torch.arange(B, device=dev)
            sel = {g: torch.tensor([b for b in range(B) if plan.group[b] == g], device=dev) for g in used[1:]}
            pair = []
            for m in range(6):
                t = grp[GROUP_PAIRS[used[0]][m]].float().clone()
                for g in used[1:]:
                    t[sel[g]] = grp[GROUP_PAIRS[g][m]].float()[sel[g]]
                pair.append(t)
            types = plan.types
            first_img, first_dep, second_img, second_dep, flow, back = pair  # private copies: augmented samples are overwritten in place
            which = torch.tensor(plan.which, device=dev)
            # geometric augmentation (types 5-7) of the samples that drew one: one native call for the sub-batch
            geo = [b for b, t in enumerate(types) if t >= 5]
            if geo:
                gi = torch.tensor(geo, device=dev)
                sub = [x[gi].contiguous() for x in pair]
                set1, set2, _ = synthesis.augment_flow_batch(*sub, kinds=[types[b] for b in geo], inpaint=self.inpaint,
                                                             params=[plan.draws[b] for b in geo])
                w0 = (which[gi] == 0).view(-1, 1, 1, 1)
                # set1 = (aug_imgA, aug_depthA, augA_flow, back_augA_flow, imgB, depthB); set2 = (imgA, depthA, aug1_flow, back_aug1_flow, aug_imgB, aug_depthB)
                for dst, a, c in ((first_img, set1[0], set2[0]), (first_dep, set1[1], set2[1]), (flow, set1[2], set2[2]),
                                  (back, set1[3], set2[3]), (second_img, set1[4], set2[4]), (second_dep, set1[5], set2[5])):
                    dst[gi] = torch.where(w0, a.float(), c.float())
            # photometric augmentation (types 0-2, preprocess.py:150-182) of the image `which` points at, batched per type with
            # synthesis.photometric_apply's expressions: brightness scale / one-channel shift / grayscale
            for wsel, tgt in ((0, first_img), (1, second_img)):
                for t in (0, 1, 2):
                    idx = [b for b in range(B) if types[b] == t and plan.which[b] == wsel]
                    if not idx:
                        continue
                    ii = torch.tensor(idx, device=dev)
                    src = tgt[ii]
                    if t == 0:
                        scale = torch.stack([plan.draws[b].reshape(()) for b in idx]).to(dev)
                        tgt[ii] = src * scale.view(-1, 1, 1, 1)
                    elif t == 1:
                        ch = torch.tensor([plan.draws[b][0] for b in idx], device=dev)
                        shift = torch.stack([plan.draws[b][1].reshape(()) for b in idx]).to(dev)
                        src[torch.arange(len(idx), device=dev), ch] += shift.view(-1, 1, 1)
                        tgt[ii] = src
                    else:
                        gray = (src[:, 0] * 0.2989 + src[:, 1] * 0.5870) + src[:, 2] * 0.1140
                        tgt[ii] = gray.unsqueeze(1).expand_as(src)
        label = torch.zeros((B, NUM_CLASSES), dtype=torch.float32, device=dev)
        label[rows, torch.tensor([max(0, t - 4) for t in types], device=dev)] = 1.0  # dataloader.py:153-156
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

