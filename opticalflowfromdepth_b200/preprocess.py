"""Drop-in for the driver side of the reference's preprocess.py (the caller of the hot path and the
`.npz` format on its other side): `PreprocessPlusAugment.forward(datas, output_dir, is_stereo)` writes the same 121 files
per frame — `group.npz` (44 channels, preprocess.py:437-447) and `{group}_{k}_{1,2}.npz` for the 12 augmentations of each
of the 5 pairs (preprocess.py:453-476) — with the same keys, shapes and values.

What differs from the reference is how it gets there:
  * the group's 7 splats run as the fused kernels of `synthesis.synthesize_group` (9 launches);
  * the 45 geometric augmentations of a frame run as FIVE `ofd_augment_pairs` calls (one per pair, batch of 9), their
    random parameters pre-drawn on the host in the reference's exact order (utils.py:96-100), so the files are the same;
  * results cross PCIe once per batch into host memory and are compressed / written by a pool of writer threads
    (zlib releases the GIL) while the GPU goes on with the next frame;
  * `utils.inpaint` (OpenCV Telea on the CPU) is a pluggable hook: "reference" reproduces it, None skips it.
Float64 dataset inputs (cv2.imread(...).astype(float), utils.py:44-72) follow the reference's type promotion end to end
(synthesis._synthesize_group_f64 / augment_pairs_f64): float64 disparity flow and FW targets, float64 depth x ray product in
the 0->3 reprojection, float64 flow composition (flow02) and float64 targets wherever the reference's warp flow is float64.
Saved arrays are float32 unless save_dtype says otherwise (the reference saves the float64 promotion of the same values;
flow01 / flow02 / depth0, which ARE float64 there, are rounded once on save).
"""
from __future__ import annotations

import os
import queue
import threading
import time
from argparse import ArgumentParser
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib, geometry, ops, sweep, synthesis

AUGMENT_TYPES = [0, 5, 6, 7, 1, 5, 6, 7, 2, 5, 6, 7]  # preprocess.py:454
GROUP_CHANNELS = ("img0", "depth0", "img1", "depth1", "img2", "depth2", "img3", "depth3", "img2_prime", "depth2_prime",
                  "img3_prime", "depth3_prime", "flow01", "back_flow01", "flow12", "back_flow12", "flow02",
                  "back_flow02_prime", "flow03", "back_flow03", "flow13", "back_flow13_prime")  # preprocess.py:437-441
# the 5 pairs of a group (preprocess.py:427-432): (imgA, depthA, imgB, depthB, flowAB, back_flowAB)
GROUP_PAIRS = (("img0", "depth0", "img1", "depth1", "flow01", "back_flow01"),
               ("img1", "depth1", "img2", "depth2", "flow12", "back_flow12"),
               ("img0", "depth0", "img2_prime", "depth2_prime", "flow02", "back_flow02_prime"),
               ("img0", "depth0", "img3", "depth3", "flow03", "back_flow03"),
               ("img1", "depth1", "img3_prime", "depth3_prime", "flow13", "back_flow13_prime"))


def save_npz(path, level, **arrays):
    """np.savez / np.savez_compressed with a choice of deflate level: the same `.npz` container (a zip of `.npy` members, read back
    by np.load), level None = stored, 1..9 = zlib level (np.savez_compressed is level 6; on float data level 1 is ~3x faster for a
    few percent more bytes)."""
    import zipfile

    path = os.fspath(path)
    if not path.endswith(".npz"):
        path += ".npz"
    kind = zipfile.ZIP_STORED if level is None else zipfile.ZIP_DEFLATED
    with zipfile.ZipFile(path, "w", compression=kind, compresslevel=level, allowZip64=True) as zf:
        for name, value in arrays.items():
            with zf.open(name + ".npy", "w", force_zip64=True) as f:
                np.lib.format.write_array(f, np.asanyarray(value), allow_pickle=False)


class NpzWriter:
    """Asynchronous `.npz` writer: `submit(path, **arrays)` returns at once, a pool of threads compresses and writes.
    `close()` waits for everything and re-raises the first error.  compress: True = np.savez_compressed's level 6 (the
    reference's files), False = stored, or an int deflate level 1..9."""

    def __init__(self, threads: int = 4, compress=True, max_pending: int = 256):
        self._q: "queue.Queue" = queue.Queue(max_pending)
        level = None if compress is False else (6 if compress is True else int(compress))
        self._save = lambda path, **arrays: save_npz(path, level, **arrays)
        self._err: List[BaseException] = []
        self.files = 0
        self.bytes = 0
        self._lock = threading.Lock()
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(1, threads))]
        for t in self._threads:
            t.start()

    def _run(self):
        while True:
            job = self._q.get()
            if job is None:
                self._q.task_done()
                return
            path, arrays = job
            try:
                self._save(path, **arrays)
                with self._lock:
                    self.files += 1
                    self.bytes += sum(int(np.asarray(a).nbytes) for a in arrays.values())
            except BaseException as e:  # noqa: BLE001 - reported by close()
                self._err.append(e)
            finally:
                self._q.task_done()

    def submit(self, path: str, **arrays):
        if self._err:
            raise self._err[0]
        self._q.put((path, arrays))

    def drain(self):
        self._q.join()
        if self._err:
            raise self._err[0]

    def close(self):
        self.drain()
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()


photometric_draws, photometric_apply = synthesis.photometric_draws, synthesis.photometric_apply


class PreprocessPlusAugment(nn.Module):
    """preprocess.PreprocessPlusAugment (preprocess.py:328-505): same constructor / forward signature and output files."""

    def __init__(self, device, inpaint="reference", writer: Optional[NpzWriter] = None, compress: bool = True,
                 save_dtype=np.float32, quiet: bool = False, reader_compat: bool = False):
        super().__init__()
        self.device = torch.device(device)
        # "reference": OpenCV's Telea fill on the host (the reference's values); "cuda": the device fill (ofd_inpaint_telea);
        # None: skipped; or any callable (img, valid, collision) -> img
        self.inpaint: Optional[Callable] = {"reference": synthesis.inpaint, "cuda": synthesis.inpaint_cuda}.get(inpaint, inpaint) \
            if isinstance(inpaint, str) else inpaint
        self._own_writer = writer is None
        self.writer = writer if writer is not None else NpzWriter(compress=compress)
        self.save_dtype = save_dtype
        self.quiet = quiet
        # The reference's reader wants a key `augment_img` (dataloader.py:83: 0 = the file holds the augmented FIRST image,
        # layout img|depth|flow|back_flow; else the augmented second image, layout flow|back_flow|img|depth) that its own
        # writer never stores.  reader_compat=True stores it, so dataloader.py reads these files as is.
        self.reader_compat = reader_compat
        self.counters = None

    # ---- the group (preprocess.py:341-447) -------------------------------------------------------------------------
    def synthesize(self, datas, is_stereo=False) -> Dict[str, torch.Tensor]:
        dev = self.device
        if not is_stereo:
            img0, depth = datas
        else:
            img0, _img1_unused, depth = datas  # the real right view is ignored and overwritten (preprocess.py:352,361)
        img0 = img0.to(dev).float().contiguous()[None]
        if depth.dtype in (torch.uint8, torch.uint16):
            # raw PNG payload: the loaders' arithmetic runs on the device, 1-2 B/px over PCIe instead of 8
            with torch.cuda.device(dev):
                depth = ops.depth_from_png(depth.to(dev).contiguous(), "disparity" if is_stereo else "reldepth")
        elif is_stereo:
            depth = synthesis.Convert.disparity_to_depth(depth)
        depth = depth.to(dev)
        if depth.dtype not in (torch.float32, torch.float64):
            depth = depth.float()
        with torch.cuda.device(dev):
            depth0 = ops.normalize_depth(depth.contiguous()[None])
        h, w = img0.shape[-2:]
        sBf = torch.as_tensor(synthesis.Convert.disparity_scale(), dtype=torch.float32).reshape(1).to(dev)  # :356
        T1, _, _ = synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)                      # :372 -> :277
        K, inv_K = synthesis.Plausible.K((h, w))
        cam = geometry.camera_constants(K, inv_K, T1).to(dev)
        if self.counters is None:
            self.counters = ops.new_counters(dev)
        res = synthesis.synthesize_group(img0, depth0, sBf, cam, inpaint=self.inpaint, counters=self.counters)
        self.counters[_lib.CNT_FRAMES] += 1
        self.counters[_lib.CNT_PAIRS] += len(GROUP_PAIRS)
        return res

    # ---- the 60 augmentations (preprocess.py:453-476) ------------------------------------------------------------------
    def augment_pair_block(self, group: Dict[str, torch.Tensor], gi: int, draws, defer_fill: bool = False):
        """All 12 augmentations of group pair `gi` as ONE device tensor [12, 2, 8, H, W]: [k, 0] is file {gi}_{k}_1
        (set1[0:4] = aug_img0, aug_depth0, aug0_flow, back_aug0_flow) and [k, 1] is file {gi}_{k}_2 (set2[2:6] = aug1_flow,
        back_aug1_flow, aug_img1, aug_depth1), preprocess.py:459-476.  `draws[k]` are the pre-drawn host parameters.
        defer_fill: the block comes back with the un-inpainted warped images in place together with what `fill_blocks` needs to
        inpaint the images of several blocks in one batched call: returns (block, pending)."""
        fAB64 = group[GROUP_PAIRS[gi][4]] if group[GROUP_PAIRS[gi][4]].dtype == torch.float64 else None
        imgA, depA, imgB, depB, fAB, bAB = (group[n].float() for n in GROUP_PAIRS[gi])
        h, w = imgA.shape[-2:]
        geo = [k for k, t in enumerate(AUGMENT_TYPES) if t >= 5]
        n = len(geo)
        rep = lambda x: x.expand(n, -1, -1, -1).contiguous()  # noqa: E731
        block = torch.empty((len(AUGMENT_TYPES), 2, 8, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if fAB64 is not None:  # float64 dataset depth: flowAB is float64 in the reference too (pairs 0->1, 0->2')
                r = synthesis.augment_pairs_f64(rep(imgA), rep(depA), rep(imgB), rep(depB), rep(fAB64), rep(bAB),
                                                [AUGMENT_TYPES[k] for k in geo], [draws[k] for k in geo])
            else:
                r = ops.augment_pairs(rep(imgA), rep(depA), rep(imgB), rep(depB), rep(fAB), rep(bAB),
                                      [AUGMENT_TYPES[k] for k in geo], [draws[k] for k in geo])
            a_img0, a_img1 = r["aug_img0"], r["aug_img1"]
            if self.inpaint is not None and not defer_fill:
                a_img0 = self.inpaint(a_img0, r["valid_img0"], r["collision_img0"])
                a_img1 = self.inpaint(a_img1, r["valid_img1"], r["collision_img1"])
            # the geometric slots are 1-3, 5-7, 9-11 (AUGMENT_TYPES): strided views instead of an index tensor, whose upload would
            # synchronise the stream in the middle of a frame
            assert geo == [1, 2, 3, 5, 6, 7, 9, 10, 11]
            slots = block.view(3, 4, 2, 8, h, w)[:, 1:4]
            for ch0, t in ((0, a_img0), (3, r["aug_depth0"]), (4, r["aug0_flow"]), (6, r["back_aug0_flow"])):
                slots[:, :, 0, ch0:ch0 + t.shape[1]] = t.view(3, 3, t.shape[1], h, w)
            for ch0, t in ((0, r["aug1_flow"].float()), (2, r["back_aug1_flow"]), (4, a_img1), (7, r["aug_depth1"])):
                slots[:, :, 1, ch0:ch0 + t.shape[1]] = t.view(3, 3, t.shape[1], h, w)
            for k, t in enumerate(AUGMENT_TYPES):
                if t < 5:
                    pa = photometric_apply(imgA[0], float(t), draws[k])
                    pb = photometric_apply(imgB[0], float(t), draws[k])
                    block[k, 0] = torch.cat((pa, depA[0], fAB[0], bAB[0]), 0)
                    block[k, 1] = torch.cat((fAB[0], bAB[0], pb, depB[0]), 0)
        if defer_fill:
            return block, (geo, r["valid_img0"], r["collision_img0"], r["valid_img1"], r["collision_img1"])
        return block

    def fill_blocks(self, blocks, pendings) -> None:
        """utils.inpaint of every warped image of the given augmentation blocks (preprocess.py:127,133: 2 x 9 per group pair, 90 per
        frame) in ONE call of the hook instead of ten: the device fill is bound by its per-layer grid barrier at 9 images per call
        and costs half as much per image at 90 (profiles/r2/probe_telea_batch.txt); a fill never looks across images, so the files are
        the same.  In place: block[geo, 0, 0:3] (aug_img0) and block[geo, 1, 4:7] (aug_img1)."""
        if self.inpaint is None or not blocks:
            return
        with torch.cuda.device(self.device):
            imgs, valids, colls, where = [], [], [], []
            for block, (geo, v0, c0, v1, c1) in zip(blocks, pendings):
                h, w = block.shape[-2:]
                slots = block.view(3, 4, 2, 8, h, w)[:, 1:4]  # the geometric slots 1-3, 5-7, 9-11 without an index tensor
                for which, ch, v, c in ((0, slice(0, 3), v0, c0), (1, slice(4, 7), v1, c1)):
                    dst = slots[:, :, which, ch]
                    imgs.append(dst.reshape(len(geo), 3, h, w))
                    valids.append(v)
                    colls.append(c)
                    where.append(dst)
            filled = self.inpaint(torch.cat(imgs).contiguous(), torch.cat(valids), torch.cat(colls))
            o = 0
            for dst in where:
                n = dst.shape[0] * dst.shape[1]
                dst.copy_(filled[o:o + n].view(dst.shape))
                o += n

    @staticmethod
    def draw_augmentations(h: int, w: int):
        """Host draws of a frame's 60 augmentations in the reference's order: pair by pair, type by type (:454-455)."""
        plan = []
        for _ in GROUP_PAIRS:
            row = []
            for t in AUGMENT_TYPES:
                if t >= 5:
                    row.append(synthesis.SpecialFlow(None).params((h, w), float(t))[1])  # fresh instance per call (:114)
                else:
                    row.append(photometric_draws(float(t)))
            plan.append(row)
        return plan

    def _copy_stream(self):
        """The D2H copies of a frame run on their own stream: they overlap the kernels of the next augmentation blocks and the fills."""
        if getattr(self, "_cs", None) is None:
            self._cs = torch.cuda.Stream(self.device)
        return self._cs

    def _after_compute(self):
        """The copy stream waits for everything issued so far on the compute stream."""
        cs = self._copy_stream()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        cs.wait_event(ev)
        return cs

    def _group_to_host(self, group: Dict[str, torch.Tensor]) -> torch.Tensor:
        """The 44-channel group array of one frame (preprocess.py:437-447) assembled in page-locked host memory: each result
        tensor goes by one strided DMA into its channel slice, no torch.cat on the device.  Asynchronous (copy stream): the caller
        synchronises that stream before reading the returned page-locked tensor [1,44,H,W]."""
        h, w = group["img0"].shape[-2:]
        parts = []
        for n in GROUP_CHANNELS:
            t = group[n][0:1]
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            parts.append(t)
        host = torch.empty((1, sum(t.shape[1] for t in parts), h, w), dtype=torch.float32, pin_memory=True)
        cs = self._after_compute()
        c0 = 0
        for t in parts:
            ops.scatter_channels_to_host(t, host, c0, stream=cs)
            c0 += t.shape[1]
        self._keep = getattr(self, "_keep", []) + parts  # referenced until the copy stream has been synchronised
        return host

    def forward(self, datas, output_dir, is_stereo=False, n_continuous=4):
        t0 = time.time()
        group = self.synthesize(datas, is_stereo)
        os.makedirs(output_dir, exist_ok=True)
        geo = [k for k, t in enumerate(AUGMENT_TYPES) if t >= 5]
        with torch.cuda.device(self.device):
            self._keep = []
            host_group = self._group_to_host(group)
            t1 = time.time()
            h, w = group["img0"].shape[-2:]
            plan = self.draw_augmentations(h, w)
            # Every augmentation block starts its way to the host as soon as its kernels are issued (copy stream), so the transfers of a
            # frame - 1.2 GB at 480x640, half of its time - overlap the kernels of the following blocks and the fills.  The warped images
            # of the geometric augmentations are inpainted afterwards in ONE batched call (fill_blocks) and only those planes cross again.
            built, hosts = [], []
            for gi in range(len(GROUP_PAIRS)):
                block, pending = self.augment_pair_block(group, gi, plan[gi], defer_fill=True)
                host = torch.empty(block.shape, dtype=block.dtype, pin_memory=True)
                cs = self._after_compute()
                with torch.cuda.stream(cs):
                    host.copy_(block, non_blocking=True)
                built.append((block, pending))
                hosts.append(host)
            if self.inpaint is not None:
                self.fill_blocks([b for b, _ in built], [p for _, p in built])
                cs = self._after_compute()
                with torch.cuda.stream(cs):
                    for (block, _), host in zip(built, hosts):
                        for k in geo:
                            host[k, 0, 0:3].copy_(block[k, 0, 0:3], non_blocking=True)  # aug_img0 of file {gi}_{k}_1
                            host[k, 1, 4:7].copy_(block[k, 1, 4:7], non_blocking=True)  # aug_img1 of file {gi}_{k}_2
            self._copy_stream().synchronize()
            built.clear()
            self._keep = []
            self.writer.submit(f"{output_dir}/group.npz", img_depth_flow=host_group[0].numpy().astype(self.save_dtype, copy=False))
            for gi, host in enumerate(hosts):
                block = host.numpy().astype(self.save_dtype, copy=False)
                for k, t in enumerate(AUGMENT_TYPES):
                    for which in (0, 1):
                        extra = {"augment_img": which} if self.reader_compat else {}
                        self.writer.submit(f"{output_dir}/{gi}_{k}_{which + 1}.npz", img_depth_flow=block[k, which],
                                           augment_flow_type=t, **extra)
        if self._own_writer:
            self.writer.drain()
        if not self.quiet:
            print(f"{output_dir = }: preprocessing time = {t1 - t0:.3f}, augmenting time = {time.time() - t1:.3f}")

    def close(self):
        if self._own_writer:
            self.writer.close()


def read_args(argv=None):
    """preprocess.read_args (preprocess.py:507-517) plus the knobs this driver adds."""
    parser = ArgumentParser()
    parser.add_argument('--dataset')
    parser.add_argument('--gpu', default=0, type=int)
    parser.add_argument('--split', default=1, type=int)
    parser.add_argument('--split_id', default=0, type=int)
    parser.add_argument('--specific_epoch_idx', default=-1, type=int)
    parser.add_argument('--no_inpaint', action='store_true', help='skip utils.inpaint (OpenCV Telea on the CPU)')
    parser.add_argument('--inpaint', default='reference', choices=['reference', 'cuda'],
                        help="Telea fill: 'reference' = cv2.inpaint on host threads (the reference's values), 'cuda' = ofd_inpaint_telea on the device")
    parser.add_argument('--writer_threads', default=8, type=int)
    parser.add_argument('--compress_level', default=6, type=int,
                        help='deflate level of the .npz files: 6 = np.savez_compressed (the reference), 1 = ~3x faster, 0 = stored')
    parser.add_argument('--output_root', default='datasets/AugmentedDatasets')
    parser.add_argument('--epochs', default=2, type=int, help='the reference always runs 2 (preprocess.py:552)')
    parser.add_argument('--reader_compat', action='store_true',
                        help="also store the `augment_img` key the reference's dataloader.py reads but its writer omits")
    return parser.parse_args(argv)


def run(dataset, output_dir: str, is_stereo: bool, args, epochs=2) -> Dict[str, int]:
    """The reference's driver loop (preprocess.py:540-561): contiguous shard [start, end) of the dataset, two epochs,
    per-image reseeding (12345 + img_idx + epoch_idx * len(dataset))."""
    device = f"cuda:{args.gpu}"
    lvl = getattr(args, "compress_level", 6)
    writer = NpzWriter(threads=args.writer_threads, compress=False if lvl == 0 else lvl)
    ppa = PreprocessPlusAugment(device=device, inpaint=None if args.no_inpaint else getattr(args, "inpaint", "reference"), writer=writer,
                                reader_compat=getattr(args, "reader_compat", False))
    rng = sweep.shard_range(len(dataset), args.split, args.split_id)
    epochs = getattr(args, "epochs", epochs)
    for epoch_idx in range(epochs):
        for img_idx in rng:
            synthesis.set_seed(sweep.frame_seed(img_idx, epoch_idx, len(dataset)))
            datas = dataset[img_idx]
            ppa(datas, f"{output_dir}/{img_idx + epoch_idx * len(dataset)}", is_stereo)
    writer.close()
    return sweep.reduce_counters(ppa.counters) if ppa.counters is not None else {}


class SyntheticDataset:
    """`--dataset synthetic:N[:HxW]`: N DIML-shaped synthetic RGB-D frames (BASELINE config 5), float64 depth like the
    reference's loaders deliver."""

    def __init__(self, n: int, h: int = 480, w: int = 640):
        self.n, self.h, self.w = n, h, w

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        from . import synthetic

        img, raw = synthetic.diml_frame(idx, self.h, self.w)
        return torch.from_numpy(img), torch.from_numpy(raw.astype(np.float64))


def main(argv=None):
    """`python -m opticalflowfromdepth_b200.preprocess --dataset DIML|ReDWeb|synthetic:N --gpu 0 --split S --split_id K`
    (README.md:53 of the reference).  DIML / ReDWeb use the reference's own dataset classes (its `dataloader.py` must be
    importable, i.e. run from / with PYTHONPATH at the reference checkout)."""
    args = read_args(argv)
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        # torchrun: one process per GPU, the dataset sharded by rank exactly like the reference's manual --split / --split_id;
        # the only collective is the end-of-run counter sum
        import torch.distributed as dist

        args.split, args.split_id = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
        args.gpu = int(os.environ.get("LOCAL_RANK", args.split_id))
        torch.cuda.set_device(args.gpu)
        if args.split > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", args.gpu))
    if args.dataset and args.dataset.startswith("synthetic"):
        parts = args.dataset.split(":")
        n = int(parts[1]) if len(parts) > 1 else 8
        h, w = (int(v) for v in parts[2].split("x")) if len(parts) > 2 else (480, 640)
        dataset, is_stereo, name = SyntheticDataset(n, h, w), False, "synthetic"
    elif args.dataset in ("DIML", "ReDWeb"):
        import dataloader  # the reference's module (datasets + file lists)

        dataset = dataloader.DIML() if args.dataset == "DIML" else dataloader.ReDWeb()
        is_stereo, name = args.dataset == "DIML", args.dataset
    else:
        raise SystemExit(f"unknown --dataset {args.dataset!r}")
    t0 = time.time()
    totals = run(dataset, f"{args.output_root}/{name}", is_stereo, args)
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"done in {time.time() - t0:.1f} s: {totals}")


if __name__ == "__main__":
    main()
