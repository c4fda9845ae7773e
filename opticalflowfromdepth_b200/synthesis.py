"""Flow-synthesis operators with the names and argument meaning of the reference's preprocess.py / utils.py
(`Plausible`, `Convert`, `ConcatFlow`, `BackFlow`, `SpecialFlow`, `augment_flow`, `normalize_depth`,
`fix_warped_depth`, `get_random`, `set_seed`), each backed by one fused libofd_b200 kernel instead of a chain of
torch ops, plus batched frame-level entry points (`synthesize_pairs`, `synthesize_group`).

Random numbers are drawn on the host with torch's CPU generator in the reference's order (utils.py:96-100), so a
`set_seed` + call sequence reproduces the reference's poses and disparity scales exactly.
"""
from __future__ import annotations

import math
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from torch import nn

from . import geometry, ops
from .fw import FW

__all__ = ["Plausible", "Convert", "ConcatFlow", "BackFlow", "SpecialFlow", "augment_flow", "augment_flow_batch", "augment_pairs_f64", "sample_special_params",
           "photometric_draws", "photometric_apply", "frame_draws_batch",
           "normalize_depth", "fix_warped_depth", "get_random", "set_seed", "inpaint", "inpaint_cuda", "synthesize_pairs", "synthesize_group"]


# ---- utils.py helpers ------------------------------------------------------------------------------------------
def set_seed(seed=42, loader=None):
    """utils.set_seed (utils.py:178-188) without the cudnn switches; `loader` is accepted and seeded like the reference does."""
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    try:
        loader.sampler.generator.manual_seed(seed)
    except AttributeError:
        pass


def get_random(random_range, random_begin, random_sign=True):
    """utils.get_random (utils.py:96-100): optional sign draw (randint) THEN the magnitude draw (rand)."""
    sign = torch.randint(0, 2, (1,))[0] * 2 - 1 if random_sign else torch.tensor(1)
    value = torch.rand(1)[0] * random_range + torch.tensor(random_begin)
    return sign * value


def normalize_depth(depth):
    """utils.normalize_depth (utils.py:102-116) for depth[1,H,W] or [B,1,H,W] CUDA tensors; out of place."""
    d = depth if depth.dim() == 4 else depth.unsqueeze(0)
    with torch.cuda.device(d.device):
        out = ops.normalize_depth(d.contiguous())
    return out if depth.dim() == 4 else out.squeeze(0)


def fix_warped_depth(depth):
    """utils.fix_warped_depth (utils.py:123-126): in place, returns its argument."""
    if depth.is_contiguous() and depth.dtype == torch.float32:
        with torch.cuda.device(depth.device):
            return ops.fix_warped_depth_(depth)
    fixed = ops.fix_warped_depth_(depth.float().contiguous())
    depth.copy_(fixed)
    return depth


def inpaint(img, valid, collision, backend="cv2"):
    """utils.inpaint (utils.py:136-151).  The hole mask is computed on the GPU (ofd_inpaint_mask); the Telea fill is
      backend="cv2":  OpenCV on the host exactly as in the reference (cv2.inpaint radius 3, the image crossing as uint8, the frames
                      of a batch spread over host threads) - the reference's values, ~12 ms per 480x640 frame of host time;
      backend="cuda": ofd_inpaint_telea - Telea's fast-marching fill with OpenCV's per-pixel arithmetic, marched in layers on the
                      device, one cooperative kernel per batch, nothing crosses PCIe.  Not bit-identical to cv2 (the serial heap
                      order is replaced by layer order); see tests/test_gpu_parity.py for the measured difference.
    Accepts [3,H,W] + [1,H,W] or batched [B,3,H,W] + [B,1,H,W]; returns float32 on img's device."""
    single = img.dim() == 3
    im = img.unsqueeze(0) if single else img
    v = (valid.unsqueeze(0) if single else valid).float().contiguous()
    c = (collision.unsqueeze(0) if single else collision).float().contiguous()
    with torch.cuda.device(im.device):
        mask_dev = ops.inpaint_mask(v, c)
        if backend == "cuda":
            res = ops.inpaint_telea(im.float().contiguous(), mask_dev, 3)
            return res[0] if single else res
    if backend != "cv2":
        raise ValueError("backend must be 'cv2' or 'cuda'")
    import cv2  # only needed when this hook is used

    mask = mask_dev.cpu().numpy()
    im_u8 = im.permute(0, 2, 3, 1).to(torch.uint8).cpu().numpy()  # .astype(np.uint8): truncation, utils.py:147
    n = im_u8.shape[0]

    def fill(b):
        return cv2.inpaint(np.ascontiguousarray(im_u8[b]), mask[b, 0], 3, cv2.INPAINT_TELEA)

    if n > 1:  # OpenCV releases the GIL: the frames of a batch are filled on a pool of host threads
        with ThreadPoolExecutor(max_workers=min(n, os.cpu_count() or 1)) as pool:
            out = np.stack(list(pool.map(fill, range(n))))
    else:
        out = np.stack([fill(0)])
    res = torch.from_numpy(out.astype(np.float32)).permute(0, 3, 1, 2).contiguous().to(im.device)
    return res[0] if single else res


def inpaint_cuda(img, valid, collision):
    """synthesis.inpaint with the device fill (a hook for synthesize_group / augment_flow / PreprocessPlusAugment(inpaint="cuda"))."""
    return inpaint(img, valid, collision, backend="cuda")


# ---- preprocess.py:184-235 -------------------------------------------------------------------------------------
class Plausible:
    @staticmethod
    def f():
        return 1

    @staticmethod
    def B():
        return 50

    @staticmethod
    def K(size, another=False):
        """Pinhole intrinsics fx=.58w, fy=.58h, cx=.5w, cy=.5h as [1,4,4] and their inverse (preprocess.py:194-209)."""
        h, w = size
        K = torch.eye(4, dtype=torch.float32).unsqueeze(0)
        K[0, 0, 0] = 0.58
        K[0, 1, 1] = 0.58
        K[0, 0, 2] = 0.5
        K[0, 1, 2] = 0.5
        if another:
            K[:, :2, :2] *= 2
        K[:, 0, :] *= w
        K[:, 1, :] *= h
        return K, torch.linalg.inv(K)

    @staticmethod
    def random_motion(axisangle_range, axisangle_base, translation_range, translation_base,
                      another_axisangle=None, another_translation=None):
        """Random pose (preprocess.py:212-235); draw order ax, ay, az, cx, cy, cz."""
        ang = [get_random(math.pi * axisangle_range, math.pi * axisangle_base) for _ in range(3)]
        mot = [get_random(translation_range, translation_base) for _ in range(3)]
        axisangle = torch.tensor([[ang]], dtype=torch.float32)
        translation = torch.tensor([[mot]], dtype=torch.float32)
        if another_axisangle is not None and another_translation is not None:
            T = geometry.transformation_from_parameters(axisangle + another_axisangle, translation + another_translation)
        else:
            T = geometry.transformation_from_parameters(axisangle, translation)
        return T, axisangle, translation


def frame_draws_batch(seeds, size):
    """The first random draws of PreprocessPlusAugment.forward for a batch of independently seeded frames (preprocess.py:555,
    356, 372 -> 277): per frame, after set_seed(seed), one disparity scale (Convert.disparity_scale) and one camera pose
    (Plausible.random_motion(1/36, 1/36, 0.1, 0.1)), packed for the kernels as (sBf[n] float32, cam[n,21] float32, T1[n,4,4]).

    Bit-identical to the per-frame calls (tests/test_host_logic_cpu.py::test_frame_draws_batch_equals_per_frame_draws) at a
    tenth of their host time: the 13 draws of a frame come from a private torch.Generator seeded like the global one (same
    mt19937 stream: rand, then 6 x (randint, rand) as utils.get_random orders them, utils.py:96-100), and the float32 arithmetic
    that turns them into scale, axis-angle, translation, Rodrigues rotation and (K T)[:3] runs ONCE per batch on [n,...] tensors
    (the same elementwise torch ops, so the same roundings)."""
    h, w = size
    n = len(seeds)
    gen = torch.Generator()
    rows_u, rows_s = [], []
    for seed in seeds:
        gen.manual_seed(int(seed))
        u = [torch.rand(1, generator=gen).item()]
        sg = []
        for _ in range(6):
            sg.append(torch.randint(0, 2, (1,), generator=gen).item() * 2 - 1)
            u.append(torch.rand(1, generator=gen).item())
        rows_u.append(u)
        rows_s.append(sg)
    u = torch.tensor(rows_u, dtype=torch.float32).reshape(n, 7)
    sg = torch.tensor(rows_s, dtype=torch.int64).reshape(n, 6)
    return frame_params_from_uniforms(u, sg, size)


def frame_params_from_uniforms(u, sg, size):
    """The float32 arithmetic of frame_draws_batch on given draws: u[n,7] uniforms in [0,1) (disparity scale, 3 axis-angle and 3
    translation magnitudes) and sg[n,6] signs (+-1) -> (sBf[n], cam[n,21], T1[n,4,4]).  inloop.InLoopSampler draws u and sg in two
    generator calls per batch (the distributions of the reference's draws, not its per-frame call sequence)."""
    h, w = size
    n = u.shape[0]
    sBf = (u[:, 0] * 0.3 + torch.tensor(0.8)) * Plausible.B() * Plausible.f()
    ang = sg[:, 0:3] * (u[:, 1:4] * (math.pi * (1. / 36.)) + torch.tensor(math.pi * (1. / 36.)))
    tr = sg[:, 3:6] * (u[:, 4:7] * 0.1 + torch.tensor(0.1))
    T1 = geometry.transformation_from_parameters(ang[:, None, :].contiguous(), tr[:, None, :].contiguous())
    K, inv_K = Plausible.K((h, w))
    cam = geometry.camera_constants(K.expand(n, 4, 4), inv_K.expand(n, 4, 4), T1)
    return sBf.to(torch.float32).contiguous(), cam, T1


# ---- preprocess.py:237-298 -------------------------------------------------------------------------------------
class Convert:
    @staticmethod
    def disparity_scale():
        """s * B * f as the reference forms it (preprocess.py:240-243): a float32 0-dim tensor."""
        s = get_random(0.3, 0.8, random_sign=False)
        return s * Plausible.B() * Plausible.f()

    @staticmethod
    def depth_to_disparity(depth):
        sBf = Convert.disparity_scale()
        flow = Convert._disparity_flow(depth, sBf)
        return flow[0:1] * -1.0  # disparity = -flow.x exactly

    @staticmethod
    def _disparity_flow(depth, sBf):
        d = depth.unsqueeze(0).contiguous()
        if d.dtype not in (torch.float32, torch.float64):
            d = d.float()
        s = torch.as_tensor(sBf, dtype=torch.float32).reshape(1).to(d.device)
        with torch.cuda.device(d.device):
            return ops.disparity_flow(d, s).squeeze(0)

    @staticmethod
    def depth_to_disparity_flow(depth, device=None):
        """depth_to_disparity + disparity_to_flow(random_sign=False) in one kernel (preprocess.py:356-357)."""
        return Convert._disparity_flow(depth.to(device) if device is not None else depth, Convert.disparity_scale())

    @staticmethod
    def disparity_to_flow(disparity, device=None, random_sign=True):
        flow = torch.cat((disparity, torch.zeros_like(disparity)), axis=0) * -1.0
        if random_sign:
            flow = flow * get_random(0, 1)
        return flow.to(device)

    @staticmethod
    def disparity_to_depth(disparity):
        return Plausible.B() * Plausible.f() / (disparity + 0.005)

    @staticmethod
    def depth_to_random_flow(depth, device=None, segment=None, T1=None):
        """Random 6-DoF reprojection flow (preprocess.py:265-298): depth[1,h,w] -> (flow[2,h,w] float32, T1[1,4,4])."""
        _, h, w = depth.shape
        K, inv_K = Plausible.K((h, w))
        if T1 is None:
            T1, _, _ = Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)
        dev = torch.device(device) if device is not None else depth.device
        d = depth.to(dev).unsqueeze(0)
        if d.dtype not in (torch.float32, torch.float64):
            d = d.float()
        flow = geometry.depth_to_flow(d, K, inv_K, T1.cpu())
        return flow.squeeze(0), T1.to(dev)


# ---- preprocess.py:301-326 -------------------------------------------------------------------------------------
class ConcatFlow(nn.Module):
    """flowAC on A's grid = valid * (FW(flowBC, back_flowAB, depthB) + flowAB); add+mask fused in the gather."""

    def __init__(self, device=None):
        super().__init__()
        self.device = device
        self.fw = FW(device)

    def forward(self, flowAB, back_flowAB, flowBC, imgB_depth):
        concat, valid, _ = self.fw(flowBC, back_flowAB, imgB_depth, epilogue=ops.EPI_CONCAT, aux=flowAB)
        return concat, valid


class BackFlow(nn.Module):
    """back_flowAB on B's grid = -valid * FW(flowAB, flowAB, depthA); negate+mask fused in the gather."""

    def __init__(self, device=None):
        super().__init__()
        self.device = device
        self.fw = FW(device)

    def forward(self, flowAB, imgA_depth):
        back, valid, _ = self.fw(flowAB, flowAB, imgA_depth, epilogue=ops.EPI_BACK)
        return back, valid


# ---- preprocess.py:24-105 --------------------------------------------------------------------------------------
class SpecialFlow(nn.Module):
    """Analytic augmentation flows; returns (p1 - p0, p_prev - p0) as [2,h,w] float32 tensors.

    Like the reference, a fresh instance takes the vertical-flip branch and the [[1,s],[0,1]] shear branch first (the toggles at
    preprocess.py:49,83 flip from their initial True on first use) and alternates on later uses of the same instance; the
    reference builds a fresh instance per augment_flow call, so in its pipeline only the first branches occur."""

    def __init__(self, device=None):
        super().__init__()
        self.device = device
        self.horizontal_flip = True
        self.horizontal_shear = True

    def params(self, size, augment_flow_type):
        """The host half of forward(): draws the random parameters in the reference's order (utils.py:96-100) and returns
        (kind, params) with params = [cx, cy, M(4), Mrev(4)] or None for the flip."""
        h, w = size
        if augment_flow_type >= 7.:
            self.horizontal_shear = not self.horizontal_shear
            s = get_random(0.15, 0.2)
            if self.horizontal_shear:
                m, mrev = [1, 0, s, 1], [1, 0, -s, 1]
            else:
                m, mrev = [1, s, 0, 1], [1, -s, 0, 1]
            return 7, [0.0, 0.0] + [float(v) for v in m + mrev]
        if augment_flow_type >= 6.:
            c0 = (get_random(w / 4, w / 2) + w / 2, get_random(h / 4, h / 2) + h / 2)
            c0 = torch.tensor(c0)
            theta = torch.deg2rad(get_random(2, 8))
            rot = torch.tensor([[torch.cos(theta), -torch.sin(theta)], [torch.sin(theta), torch.cos(theta)]]).type(torch.float32)
            rev = torch.tensor([[torch.cos(-theta), -torch.sin(-theta)], [torch.sin(-theta), torch.cos(-theta)]]).type(torch.float32)
            return 6, [float(c0[0]), float(c0[1])] + [float(v) for v in rot.reshape(-1)] + [float(v) for v in rev.reshape(-1)]
        if augment_flow_type >= 5.:
            self.horizontal_flip = not self.horizontal_flip
            if self.horizontal_flip:
                # second use of one instance (preprocess.py:52): p1 = (w-1-x, y).  Expressed as the affine map about
                # c = ((w-1)/2, 0) with M = diag(-1, 1): every intermediate is a half-integer, so the flow (w-1-2x, +0) is exact
                return 6, [(w - 1) / 2.0, 0.0, -1.0, 0.0, 0.0, 1.0, -1.0, 0.0, 0.0, 1.0]
            return 5, None
        raise ValueError("augment_flow_type must be >= 5 for a special flow")

    def forward(self, size, augment_flow_type):
        h, w = size
        dev = self.device if self.device is not None else "cuda"
        kind, params = self.params(size, augment_flow_type)
        with torch.cuda.device(dev):
            return ops.special_flow(kind, params, h, w, dev)


def sample_special_params(kinds, size, generator=None):
    """Vectorised parameter draw for a batch of special flows: the same distributions as SpecialFlow.params
    (preprocess.py:62-99: theta = +-[8,10) deg, centre = +-[w/2,3w/4)+w/2, +-[h/2,3h/4)+h/2, shear = +-[0.2,0.35)) from
    TWO generator calls per batch instead of up to six per sample.  The draw ORDER therefore differs from the reference's
    per-call sequence (in-loop augmentation needs the distribution, not the sequence); augment_flow_batch(...,
    reference_draws=True) keeps the reference's sequence."""
    h, w = size
    B = len(kinds)
    f32 = np.float32
    sign = (torch.randint(0, 2, (B, 3), generator=generator).numpy() * 2 - 1).astype(f32)
    u = torch.rand((B, 3), generator=generator).numpy()
    # float32 numpy arithmetic from here on: ~20 us per batch instead of ~15 small torch ops
    cx = sign[:, 0] * (u[:, 0] * f32(w / 4) + f32(w / 2)) + f32(w / 2)
    cy = sign[:, 1] * (u[:, 1] * f32(h / 4) + f32(h / 2)) + f32(h / 2)
    theta = (sign[:, 2] * (u[:, 2] * f32(2) + f32(8))) * f32(math.pi / 180.0)
    shear = sign[:, 2] * (u[:, 2] * f32(0.15) + f32(0.2))
    cos, sin = np.cos(theta), np.sin(theta)
    cols = np.stack((cx, cy, cos, -sin, sin, cos, cos, sin, -sin, cos, shear), 1).astype(f32).tolist()
    out = []
    for b in range(B):
        k = int(kinds[b])
        c = cols[b]
        if k == 5:
            out.append(None)
        elif k == 6:
            out.append(c[:10])
        elif k == 7:
            out.append([0.0, 0.0, 1.0, c[10], 0.0, 1.0, 1.0, -c[10], 0.0, 1.0])
        else:
            raise ValueError("kinds must be 5 (flip), 6 (rotate) or 7 (shear)")
    return out


def photometric_draws(augment_flow_type: float):
    """Host random draws of the photometric branch of augment_flow in the reference's order (preprocess.py:150-163)."""
    if augment_flow_type >= 2.:
        return None
    if augment_flow_type >= 1.:
        channel = int(get_random(3, 0, False))
        shift = get_random(10, 15)
        return channel, shift
    return get_random(1, 0, False)


def photometric_apply(img: torch.Tensor, augment_flow_type: float, draws) -> torch.Tensor:
    """augment_img_func of preprocess.py:150-163 on img[...,3,H,W]: 0 brightness scale, 1 one-channel shift, 2 grayscale."""
    if augment_flow_type >= 2.:
        # (img.permute(1,2,0) @ gray).permute(2,0,1) with gray[k, :] = (0.2989, 0.5870, 0.1140)[k]: every output channel is the
        # same K=3 dot product, evaluated here as an ascending-k chain (the reference's order is a BLAS detail)
        r, g, b = img.select(-3, 0), img.select(-3, 1), img.select(-3, 2)
        gray = (r * 0.2989 + g * 0.5870) + b * 0.1140
        return gray.unsqueeze(-3).expand_as(img).contiguous()
    # the draws are float32 scalars on the host: a Python float carries the same value into the float32 kernel as a 0-dim device tensor
    # would, without the blocking host-to-device copy (and the stream synchronisation it implies) of `.to(img.device)`
    if augment_flow_type >= 1.:
        channel, shift = draws
        out = img.clone()
        out.select(-3, channel).add_(float(shift))
        return out
    return img * float(draws)


def augment_flow(img0, img0_depth, img1, img1_depth, flow01, back_flow01, device=None, augment_flow_type=None,
                 inpaint=None):
    """preprocess.augment_flow (preprocess.py:106-182): the geometric branch (types 5-7, 6 splats) and the photometric one
    (types 0-2).  `inpaint(img, valid, collision)` is the caller's hole filler (utils.inpaint in the reference — CPU OpenCV,
    outside this path); None skips it."""
    _, h, w = img0.shape
    if augment_flow_type is None:
        augment_flow_type = get_random(8, 0, False)
    if augment_flow_type < 3.:
        # photometric branch (preprocess.py:150-182): brightness scale / one-channel shift / grayscale, no warping
        draws = photometric_draws(float(augment_flow_type))
        aug0 = photometric_apply(img0, float(augment_flow_type), draws)
        aug1 = photometric_apply(img1, float(augment_flow_type), draws)
        return ((aug0, img0_depth, flow01, back_flow01, img1, img1_depth),
                (img0, img0_depth, flow01, back_flow01, aug1, img1_depth), int(augment_flow_type), None)
    if augment_flow_type < 5.:
        return None  # the reference's `elif augment_flow_type >= 3.: pass` (preprocess.py:148-149) falls off the end
    fw, cf, bf, sf = FW(device), ConcatFlow(device), BackFlow(device), SpecialFlow(device)
    special_flow, back_special_flow = sf((h, w), augment_flow_type)
    aug0_flow, _ = cf(back_special_flow, special_flow, flow01, img0_depth)
    aug1_flow, _ = cf(flow01, back_flow01, special_flow, img1_depth)

    def warp(img, depth):
        allc, valid, collision = fw(torch.cat((img, depth), axis=0), special_flow, depth)
        a_img, a_depth = allc[0:3], fix_warped_depth(allc[3:4].contiguous())
        if inpaint is not None:
            a_img = inpaint(a_img, valid, collision)
        return a_img, a_depth

    aug_img0, aug_img0_depth = warp(img0, img0_depth)
    aug_img1, aug_img1_depth = warp(img1, img1_depth)
    back_aug0_flow, _ = bf(aug0_flow, aug_img0_depth)
    back_aug1_flow, _ = bf(aug1_flow, img0_depth)
    return ((aug_img0, aug_img0_depth, aug0_flow, back_aug0_flow, img1, img1_depth),
            (img0, img0_depth, aug1_flow, back_aug1_flow, aug_img1, aug_img1_depth),
            int(augment_flow_type), (special_flow, back_special_flow))


@torch.no_grad()
def augment_flow_batch(img0, depth0, img1, depth1, flow01, back_flow01, kinds, inpaint=None, params=None,
                       reference_draws=True, generator=None):
    """Batched in-loop geometric augmentation (BASELINE config 4): the geometric branch of augment_flow
    (preprocess.py:116-147) for B pairs at once, one special flow per sample (kinds[b] in {5 flip, 6 rotate, 7 shear}).

    All tensors are [B,C,H,W] float32 CUDA.  One native call (ofd_augment_pairs) issues the special-flow kernel and the six
    batched splats back to back.  Random parameters: `params` (list of 10 floats / None per sample) if given; else drawn
    per sample in batch order with the reference's draw order (reference_draws=True, reproduces augment_flow sample by
    sample) or with sample_special_params (two generator calls per batch).
    Returns (set1, set2, (special_flow, back_special_flow)) with the same members as augment_flow, batched."""
    B, _, h, w = img0.shape
    if params is None:
        if reference_draws:
            params = []
            for b in range(B):
                gen = SpecialFlow(None)  # a fresh instance per call in the reference (preprocess.py:114)
                params.append(gen.params((h, w), float(kinds[b]))[1])
        else:
            params = sample_special_params(kinds, (h, w), generator)
    with torch.cuda.device(img0.device):
        r = ops.augment_pairs(img0, depth0, img1, depth1, flow01, back_flow01, [int(k) for k in kinds], params)
        aug_img0, aug_img1 = r["aug_img0"], r["aug_img1"]
        if inpaint is not None:
            aug_img0 = inpaint(aug_img0, r["valid_img0"], r["collision_img0"])
            aug_img1 = inpaint(aug_img1, r["valid_img1"], r["collision_img1"])
    return ((aug_img0, r["aug_depth0"], r["aug0_flow"], r["back_aug0_flow"], img1, depth1),
            (img0, depth0, r["aug1_flow"], r["back_aug1_flow"], aug_img1, r["aug_depth1"]),
            (r["special_flow"], r["back_special_flow"]))


def augment_pairs_f64(img0, depth0, img1, depth1, flow01, back_flow01, kinds, params):
    """ops.augment_pairs when flow01 is float64 (pairs 0->1 and 0->2' of a float64-depth frame): the reference's type promotion
    makes aug1_flow = (FW(special, back_flow01) + flow01) * valid float64 (preprocess.py:122,312) and evaluates the targets of its
    BackFlow splat in float64 (:138, fw.py:31).  Composed from the general splat entry points (6 splats, 12 launches + the
    special-flow kernel); everything else is float32 exactly as in ofd_augment_pairs.  Returns the same dict."""
    B, _, H, W = img0.shape
    dev = img0.device
    special, back_special = ops.special_flow_batch([int(k) for k in kinds], params, H, W, dev)
    aug0_flow, _, _ = ops.splat_flow(flow01.float(), special, depth0, epilogue=ops.EPI_CONCAT, aux=back_special)   # :121
    w1, v1, _ = ops.splat_flow(special, back_flow01, depth1)                                                       # :122
    aug1_flow = (w1 + flow01) * v1
    r = dict(special_flow=special, back_special_flow=back_special, aug0_flow=aug0_flow, aug1_flow=aug1_flow)
    for tag, img, dep in (("0", img0, depth0), ("1", img1, depth1)):                                                # :124-135
        allc, valid, coll = ops.splat_flow(torch.cat((img, dep), 1), special, dep)
        r["aug_img" + tag] = allc[:, 0:3].contiguous()
        r["aug_depth" + tag] = ops.fix_warped_depth_(allc[:, 3:4].contiguous())
        r["valid_img" + tag], r["collision_img" + tag] = valid, coll
    r["back_aug0_flow"], _, _ = ops.splat_flow(aug0_flow, aug0_flow, r["aug_depth0"], epilogue=ops.EPI_BACK)       # :137
    r["back_aug1_flow"], _, _ = ops.splat_flow(aug1_flow.float(), aug1_flow, depth0, epilogue=ops.EPI_BACK)        # :138
    return r


# ---- batched frame-level entry points ----------------------------------------------------------------------------
@torch.no_grad()
def synthesize_pairs(img0, depth0, sBf, want_flow=True, want_collision=True, counters=None):
    """Batched virtual-stereo pair synthesis (preprocess.py:356-365 minus inpaint), ONE kernel for the batch.

    img0[B,3,H,W] f32, depth0[B,1,H,W] f32|f64 (already normalised), sBf[B] f32 (= s*B*f per frame), all CUDA.
    Returns dict(img1, depth1, back_flow, flow, valid, collision)."""
    with torch.cuda.device(img0.device):
        img1, depth1, back, flow, valid, coll = ops.disparity_pair(img0, depth0, sBf, want_flow, want_collision, counters)
    return dict(img1=img1, depth1=depth1, back_flow=back, flow=flow, valid=valid, collision=coll)


@torch.no_grad()
def _fill_leaves(inpaint, leaves):
    """The inpainted images 2, 3, 2' and 3' of a group are results only - nothing downstream reads them (preprocess.py:381,393,410,423) -
    so their four fills go out as ONE batched call: the device fill is bound by its per-layer grid barrier at small batches (a batch of
    4 x B images costs little more than one of B, tools/probe_telea_batch.py), and a fill never looks across images, so the values are
    those of four separate calls.  leaves: [(img, valid, collision)] x 4 -> [img] x 4."""
    if inpaint is None:
        return [im for im, _, _ in leaves]
    n = leaves[0][0].shape[0]
    out = inpaint(torch.cat([im for im, _, _ in leaves]), torch.cat([v for _, v, _ in leaves]), torch.cat([c for _, _, c in leaves]))
    return list(out.split(n))


def synthesize_group(img0, depth0, sBf, cam, inpaint=None, counters=None):
    """The reference's 5-pair group of one frame batch (preprocess.py:356-432), inpaint optional, all on the GPU:
    7 splats with the flow producers and consumers fused into them = 9 kernel launches per batch (1 fused stereo pair,
    2 x (6-DoF z-test + gather), 2 x (row-local ConcatFlow that also runs the next frame splat's z-test + gather); 11 when the width
    is not a multiple of 4).

    img0[B,3,H,W], depth0[B,1,H,W] f32 (normalised), sBf[B], cam1/cam0 built from the same pose: cam[B,21] float32.
    depth0 may be float64 (dataset path, utils.py:44-72): the group then follows the reference's dtype rules end to end
    (_synthesize_group_f64: float64 disparity flow, float64 depth x ray product, float64 flow composition and targets).
    Returns a dict of tensors named as in preprocess.py (float32; on the float64 path depth0, flow01 and flow02 are float64
    like the reference's)."""
    if depth0.dtype == torch.float64:
        return _synthesize_group_f64(img0, depth0, sBf, cam, inpaint, counters)
    dev = img0.device
    fill = (lambda im, v, c: im) if inpaint is None else inpaint
    wc = inpaint is not None  # the collision planes only feed utils.inpaint's mask: without a fill they are not produced (4 B/px per splat)
    with torch.cuda.device(dev):
        # pair 0->1: virtual stereo (preprocess.py:356-366)
        img1, depth1, back01, flow01, valid1, coll1 = ops.disparity_pair(img0, depth0, sBf, True, wc, counters)
        img1 = fill(img1, valid1, coll1)
        # pair 1->2: random camera motion from view 1 (preprocess.py:372-382); flow computed inside the z-test
        img2, depth2, back12, flow12, valid2, coll2, _ = ops.reproject_pair(img1, depth1, cam, valid1, want_collision=wc, counters=counters)
        # pair 0->3: the same motion from view 0 (preprocess.py:385-394)
        img3, depth3, back03, flow03, valid3, coll3, _ = ops.reproject_pair(img0, depth0, cam, None, want_collision=wc, counters=counters)
        B, _, h, w = img0.shape
        if ops.concat_frame_splat_applies(flow12, w, h, B):
            # pairs 0->2' (preprocess.py:400-411) and 1->3' (:414-424): the ConcatFlow along the horizontal flow (back01.y == +0, flow01.y == -0)
            # also runs the z-test of the frame splat along its result - two launches per pair instead of three
            flow02, _, img2p, depth2p, back02p, valid2p, coll2p = ops.concat_frame_splat(flow12, back01, depth1, flow01, img0, depth0,
                                                                                            want_collision=wc, counters=counters)
            flow13, _, img3p, depth3p, back13p, valid3p, coll3p = ops.concat_frame_splat(flow03, flow01, depth1, back01, img1, depth1,
                                                                                            valid_mul=valid1, want_collision=wc, counters=counters)
        else:
            # pair 0->2': concatenated flow (preprocess.py:400-411)
            flow02, flow02_valid, _ = ops.splat_flow(flow12, back01, depth1, epilogue=ops.EPI_CONCAT, aux=flow01, want_collision=False,
                                                        horizontal=True)  # back01.y == +0: row-local kernel
            img2p, depth2p, back02p, valid2p, coll2p, _ = ops.frame_splat(img0, depth0, flow02, flow02_valid, want_collision=wc, counters=counters)
            # pair 1->3': (preprocess.py:414-424)
            flow13, flow13_valid, _ = ops.splat_flow(flow03, flow01, depth1, epilogue=ops.EPI_CONCAT, aux=back01, want_collision=False,
                                                        horizontal=True, valid_mul=valid1)  # flow01.y == -0; flow13_valid * img1_valid (:415) fused
            img3p, depth3p, back13p, valid3p, coll3p, _ = ops.frame_splat(img1, depth1, flow13, flow13_valid, want_collision=wc, counters=counters)
        img2, img3, img2p, img3p = _fill_leaves(inpaint, [(img2, valid2, coll2), (img3, valid3, coll3), (img2p, valid2p, coll2p),
                                                          (img3p, valid3p, coll3p)])
    return dict(img0=img0, depth0=depth0, img1=img1, depth1=depth1, img2=img2, depth2=depth2, img3=img3, depth3=depth3,
                img2_prime=img2p, depth2_prime=depth2p, img3_prime=img3p, depth3_prime=depth3p,
                flow01=flow01, back_flow01=back01, flow12=flow12, back_flow12=back12, flow02=flow02,
                back_flow02_prime=back02p, flow03=flow03, back_flow03=back03, flow13=flow13, back_flow13_prime=back13p,
                valid1=valid1, valid2=valid2, valid3=valid3, valid2_prime=valid2p, valid3_prime=valid3p)


def _synthesize_group_f64(img0, depth0, sBf, cam, inpaint, counters):
    """synthesize_group for float64 depth0 - what the reference's loaders deliver (cv2.imread(...).astype(float), utils.py:48,62).
    The reference then keeps float64 wherever torch's type promotion does (preprocess.py:355-425):
      flow01 = -(sBf / depth0) is float64, and its FW targets are evaluated in float64 (fw.py:31);
      flow03: depth0 * ray is a float64 product rounded once to float32 (geometry.py:39-40);
      flow02 = (FW(flow12) + flow01) * valid is float64 (:312) and drives the 0->2' splat with float64 targets;
      flow13's splat uses flow01 (float64) as the warp flow (:414); everything that comes OUT of FW is float32 (fw.py:45-58).
    Same kernels, more launches than the float32 path (the 0->3 pair is flow + splat instead of the fused pair)."""
    dev = img0.device
    fill = (lambda im, v, c: im) if inpaint is None else inpaint
    with torch.cuda.device(dev):
        img1, depth1, back01, _, valid1, coll1 = ops.disparity_pair(img0, depth0, sBf, False, True, counters)
        flow01 = ops.disparity_flow(depth0, sBf)                      # float64 (preprocess.py:356-357)
        depth0_f = depth0.float()                                     # what FW sees (fw.py:43,45)
        img1 = fill(img1, valid1, coll1)
        img2, depth2, back12, flow12, valid2, coll2, _ = ops.reproject_pair(img1, depth1, cam, valid1, counters=counters)
        flow03 = ops.reproject_flow(depth0, cam)                      # float64 depth x ray, float32 flow (geometry.py:39-40)
        img3, depth3, back03, valid3, coll3, _ = ops.frame_splat(img0, depth0_f, flow03, None, counters=counters)
        warp12, flow02_valid, _ = ops.splat_flow(flow12, back01, depth1, want_collision=False, horizontal=True)  # back01 is float32 with y == +0: row-local kernel
        flow02 = (warp12 + flow01) * flow02_valid                     # float32 + float64 -> float64 (preprocess.py:312)
        img2p, depth2p, back02p, valid2p, coll2p, _ = ops.frame_splat(img0, depth0_f, flow02, flow02_valid, counters=counters)
        flow13, flow13_valid, _ = ops.splat_flow(flow03, flow01, depth1, epilogue=ops.EPI_CONCAT, aux=back01)
        flow13_valid = flow13_valid * valid1
        img3p, depth3p, back13p, valid3p, coll3p, _ = ops.frame_splat(img1, depth1, flow13, flow13_valid, counters=counters)
        img2, img3, img2p, img3p = _fill_leaves(inpaint, [(img2, valid2, coll2), (img3, valid3, coll3), (img2p, valid2p, coll2p),
                                                          (img3p, valid3p, coll3p)])
    return dict(img0=img0, depth0=depth0, img1=img1, depth1=depth1, img2=img2, depth2=depth2, img3=img3, depth3=depth3,
                img2_prime=img2p, depth2_prime=depth2p, img3_prime=img3p, depth3_prime=depth3p,
                flow01=flow01, back_flow01=back01, flow12=flow12, back_flow12=back12, flow02=flow02,
                back_flow02_prime=back02p, flow03=flow03, back_flow03=back03, flow13=flow13, back_flow13_prime=back13p,
                valid1=valid1, valid2=valid2, valid3=valid3, valid2_prime=valid2p, valid3_prime=valid3p)
