"""Throughput of the 8f-2 driver (PreprocessPlusAugment.forward: group + 60 augmentations = 121 arrays per frame) on
480x640 synthetic frames: device + D2H side (discarding writer), and with the real compressed npz writer."""
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import preprocess as pp, synthesis  # noqa: E402


class Discard:
    files = 0

    def submit(self, path, **arrays):
        self.files += 1

    def drain(self):
        pass

    def close(self):
        pass


def main():
    ds = pp.SyntheticDataset(64)
    n = 8
    for label, writer, frames, inpaint in (("discarding writer (device + D2H + host staging)", Discard(), n, None),
                                           ("same + device Telea fill (ofd_inpaint_telea, 95 fills per frame in 3 calls: 1 + 4 + 90 images)", Discard(), n, "cuda"),
                                           ("same + utils.inpaint (OpenCV Telea, 95 calls/frame, host thread pool)", Discard(), 3, "reference"),
                                           ("npz uncompressed, 16 threads, tmpfs", pp.NpzWriter(16, compress=False), n, None),
                                           ("npz compressed (reference format), 16 threads, tmpfs", pp.NpzWriter(16, compress=True), 3, None)):
        with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
            ppa = pp.PreprocessPlusAugment("cuda:0", inpaint=inpaint, writer=writer, quiet=True)
            synthesis.set_seed(1)
            ppa(ds[0], f"{tmp}/w", False)
            writer.drain()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(frames):
                synthesis.set_seed(12345 + k)
                ppa(ds[k], f"{tmp}/{k}", False)
            writer.drain()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / frames
            print(f"{label:75s}: {dt*1e3:8.1f} ms/frame = {1/dt:6.2f} frames/s ({121/dt:7.0f} arrays/s, {300/dt:7.0f} flow pairs/s incl. augmented)", flush=True)
            writer.close()


if __name__ == "__main__":
    main()
