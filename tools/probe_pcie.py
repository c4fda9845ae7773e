"""PCIe ceiling probe for the e2e leg: pinned D2H alone, H2D alone, both at once (torch copies on two streams), then the
pair pipeline at several frames-per-call / chunk sizes."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, W = 480, 640


def bw(nbytes, fn, n=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return nbytes * n / (time.perf_counter() - t0) / 1e9


def main():
    n = 512 << 20
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def d2h():
        with torch.cuda.stream(s1):
            h_a.copy_(d_a, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s2):
            d_b.copy_(h_b, non_blocking=True)

    def both():
        d2h()
        h2d()

    print(f"D2H alone {bw(n, d2h):.1f} GB/s | H2D alone {bw(n, h2d):.1f} GB/s | both: {bw(n, both):.1f} GB/s each direction", flush=True)
    del h_a, h_b, d_a, d_b
    rng = np.random.default_rng(0)
    for Fe, chunk in ((64, 8), (128, 8), (256, 8), (256, 4), (256, 16), (256, 32)):
        h_img = torch.from_numpy(rng.integers(0, 256, (Fe, 3, H, W)).astype(np.float32)).pin_memory()
        h_dep = torch.from_numpy((rng.random((Fe, 1, H, W)) * 98 + 1).astype(np.float32)).pin_memory()
        h_s = torch.full((Fe,), 47.0)
        h_out = [torch.empty((Fe, c, H, W), dtype=torch.float32).pin_memory() for c in (3, 1, 2, 2, 1, 1)]
        pipe = ops.PairPipeline(0, H, W, chunk_frames=chunk)
        pipe.run(h_img, h_dep, h_s, *h_out)
        t0 = time.perf_counter()
        K = 5
        for _ in range(K):
            pipe.run(h_img, h_dep, h_s, *h_out)
        dt = (time.perf_counter() - t0) / K
        pipe.close()
        print(f"pipeline F={Fe:3d} chunk={chunk:2d}: {Fe/dt:7.0f} pairs/s  D2H {Fe*40*H*W/dt/1e9:.1f} GB/s  H2D {Fe*16*H*W/dt/1e9:.1f} GB/s", flush=True)
        del h_img, h_dep, h_out


if __name__ == "__main__":
    main()
