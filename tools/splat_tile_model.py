"""VERDICT r1 next #5 - would a tile-binned splat (z-buffer of a TARGET tile in shared memory, global keys only for sources that leave
the tile's source window) cut the key-plane traffic of the general splat?  This tool measures, on the flows the pipeline really
produces, the two quantities that decide it, for 64x64 target tiles and source windows = the tile shifted by its median displacement
and grown by a margin m:

    local fraction   sources whose target tile's window contains them (served by shared-memory atomics)
    stragglers       the rest: they still need the global 64-bit key (atomicMin + read + re-arm = 32 B per straggler TARGET, and the
                     tile's gather must wait for them: a grid-wide barrier or a second launch)
    re-read factor   (64 + 2m)^2 / 64^2 : every window pixel is loaded (flow 8 B + depth 4 B) by every tile whose window holds it

Projected DRAM bytes per pixel of a C=6 splat: 12 * rho + 56 + 32 * stragglers, with rho between 1 (L2 absorbs every window overlap)
and the raw re-read factor (none absorbed); today's two-launch path moves 97 B/px for 68 algorithmic (profiles/traffic.json).
Also reported: how many sources a straggler-free design would have to scan for the worst tile (border pile-ups).
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import geometry, ops, synthesis, synthetic  # noqa: E402

dev = torch.device("cuda:0")
TILE = 64


def analyse(name, flow):
    """flow [B,2,H,W] float32"""
    B, _, H, W = flow.shape
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    tx = (xs + flow[:, 0]).clamp(0, W - 1).long()
    ty = (ys + flow[:, 1]).clamp(0, H - 1).long()
    ntx = (W + TILE - 1) // TILE
    nty = (H + TILE - 1) // TILE
    tile = (torch.arange(B, device=dev)[:, None, None] * nty + ty // TILE) * ntx + tx // TILE
    dx = (xs[None] - tx).reshape(-1)
    dy = (ys[None] - ty).reshape(-1)
    tile = tile.reshape(-1)
    ntiles = B * nty * ntx
    # per-tile median displacement (the window's shift): sort by (tile, d) and pick the middle of every group
    counts = torch.bincount(tile, minlength=ntiles)
    start = torch.cumsum(counts, 0) - counts
    med = []
    for d in (dx, dy):
        key = tile * 8192 + (d + 4096)
        srt, _ = torch.sort(key)
        mid = (start + counts // 2).clamp(max=srt.numel() - 1)
        m = srt[mid] % 8192 - 4096
        med.append(torch.where(counts > 0, m, torch.zeros_like(m)))
    ex = (dx - med[0][tile]).abs()
    ey = (dy - med[1][tile]).abs()
    line = f"{name:34s} {B}x{H}x{W}  fan-in max {int(counts.max()):7d}/tile"
    for m in (8, 16, 32):
        local = ((ex <= m) & (ey <= m)).float().mean().item()
        rho = (TILE + 2 * m) ** 2 / TILE ** 2
        lo = 12 * 1.0 + 56 + 32 * (1 - local)
        hi = 12 * rho + 56 + 32 * (1 - local)
        line += f" | m={m:2d}: local {local:.4f} re-read x{rho:.2f} -> {lo:.0f}..{hi:.0f} B/px"
    print(line, flush=True)


def main():
    print(__doc__)
    for (H, W, B) in ((480, 640, 8), (1080, 1920, 4)):
        frames = [synthetic.diml_frame(k, H, W) for k in range(B)]
        img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev)
        depth = ops.normalize_depth(torch.from_numpy(np.stack([f[1] for f in frames])).to(dev))
        sBf = torch.full((B,), 47.0, device=dev)
        pair = synthesis.synthesize_pairs(img, depth, sBf)
        Kc, invK = synthesis.Plausible.K((H, W))
        cams = []
        for k in range(B):
            torch.manual_seed(12345 + k)
            cams.append(geometry.camera_constants(Kc, invK, synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)[0]))
        cam = torch.cat(cams).to(dev)
        flow03 = ops.reproject_flow(depth, cam)
        flow12 = ops.reproject_flow(pair["depth1"], cam)
        flow02, _, _ = ops.splat_flow(flow12, pair["back_flow"], pair["depth1"], epilogue=ops.EPI_CONCAT, aux=pair["flow"])
        analyse("disparity flow 0->1", pair["flow"])
        analyse("6-DoF flow 0->3 (source depth0)", flow03)
        analyse("6-DoF flow 1->2 (warped depth1)", flow12)
        analyse("concatenated flow 0->2'", flow02)
        kinds = [5 + k % 3 for k in range(B)]
        sp, _ = ops.special_flow_batch(kinds, synthesis.sample_special_params(kinds, (H, W), torch.Generator().manual_seed(1)), H, W, dev)
        analyse("special flows (flip/rotate/shear)", sp)


if __name__ == "__main__":
    main()
