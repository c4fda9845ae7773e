"""Multi-GPU sweep of independent frames (preprocess.py:511-515,540-561).

Frames are independent, so the path shards by image index with NO collective on the hot path: one process per GPU,
each runs its index range through the fused kernels; at the end of the run the ranks sum their counter blocks
(frames, pairs, hit/hole/collision/dropped pixel counts) with one all-reduce (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

COUNTER_NAMES = {"hit": _lib.CNT_HIT, "hole": _lib.CNT_HOLE, "collision": _lib.CNT_COLLISION, "dropped": _lib.CNT_DROPPED,
                 "tie_src": _lib.CNT_TIE_SRC, "frames": _lib.CNT_FRAMES, "pairs": _lib.CNT_PAIRS}


def shard_range(n: int, split: int, split_id: int) -> range:
    """The reference's contiguous shards (preprocess.py:543-547): ceil(n/split) per shard, last takes the rest."""
    if split < 1 or not (0 <= split_id < split):
        raise ValueError("need split >= 1 and 0 <= split_id < split")
    split_len = (n + split - 1) // split
    start = split_id * split_len
    end = (split_id + 1) * split_len
    if split_id == split - 1:
        end = n
    return range(min(start, n), min(end, n))


def shard_strided(n: int, world: int, rank: int) -> range:
    """idx % world == rank: balanced to within one frame for any n."""
    return range(rank, n, world)


def rank_cores(local_rank: int, local_world: int, cores: Optional[Sequence[int]] = None) -> List[int]:
    """The disjoint core set of one rank of a node: the cores this process may run on, split into `local_world` contiguous
    slices (the remainder goes to the first ranks).  With more ranks than cores every rank keeps the whole set."""
    import os

    cores = sorted(os.sched_getaffinity(0)) if cores is None else sorted(cores)
    if local_world < 1 or not (0 <= local_rank < local_world):
        raise ValueError("need local_world >= 1 and 0 <= local_rank < local_world")
    if len(cores) < local_world:
        return list(cores)
    base, rem = divmod(len(cores), local_world)
    start = local_rank * base + min(local_rank, rem)
    return list(cores[start:start + base + (1 if local_rank < rem else 0)])


def bind_rank_cores(local_rank: int, local_world: int) -> List[int]:
    """Pin this process (and every thread it starts later: the pipeline's host workers, torch's pools) to rank_cores(...),
    so that the ranks of a node do not migrate over each other's cores.  Returns the core list.  OFD_NO_AFFINITY=1 disables it."""
    import os

    if os.environ.get("OFD_NO_AFFINITY"):
        return sorted(os.sched_getaffinity(0))
    mine = rank_cores(local_rank, local_world)
    os.sched_setaffinity(0, mine)
    return mine


def frame_seed(img_idx: int, epoch_idx: int, dataset_len: int) -> int:
    """Per-image reseeding (preprocess.py:555) — makes results independent of the partition."""
    return 12345 + img_idx + epoch_idx * dataset_len


def batches(indices: Sequence[int], batch: int) -> Iterable[List[int]]:
    idx = list(indices)
    for k in range(0, len(idx), batch):
        yield idx[k:k + batch]


def reduce_counters(counters: torch.Tensor, group=None) -> Dict[str, int]:
    """Sum the per-rank counter blocks (int64 view of the uint64 slots) over all ranks; returns a dict on every rank."""
    total = counters.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    host = total.cpu().tolist()
    return {name: int(host[slot]) for name, slot in COUNTER_NAMES.items()}


class PinnedGroupSink:
    """A `sink` for run_sweep that brings every batch's 44-channel group tensor (preprocess.py:437-447 channel order) to the
    host.  The concatenation the reference builds with torch.cat never exists on the device: each of the 22 result tensors is
    copied by ONE strided DMA (ops.scatter_channels_to_host) straight into its channel slice of a page-locked [B,44,H,W] array,
    on a side stream, so the copies of batch k overlap the host-side preparation and the kernels of batch k+1.

    byte_images: the 18 image channels of a group are uint8-valued whenever the source frames are (warps select pixels, the
    fills produce bytes), so they can cross PCIe at 1 B instead of 4 (122 instead of 176 B/px per frame): the device narrows them
    with a verifying kernel (ofd_pack_u8), the host widens them into the float array on a few threads when the batch is delivered
    (ofd_host_widen_u8); a batch whose check fails is delivered from its float planes instead.  Default (None): on while PCIe is
    the limit, i.e. unless more than two ranks share the node (LOCAL_WORLD_SIZE > 2: there the node's host memory is, and the extra
    host stores cost more than the PCIe bytes they save - DESIGN.md section 6).

    const_planes: flow01.y and back_flow01.y of a group are the constants -0.0 and +0.0 (the virtual stereo pair moves pixels along
    their row, preprocess.py:253,361-363): they do not cross PCIe, the delivering threads write them into the array (114 B/px per
    frame).  Same default rule as byte_images.

    `on_batch(idx_list, array[B,44,H,W], release)` receives the host array once its copies have landed (when the next-but-one
    batch arrives, or at flush()).  The consumer OWNS the array until it calls release() - it may hand it to asynchronous
    writers (preprocess.NpzWriter) and release it when they are done; the sink takes another page-locked buffer from its pool
    (allocating one if none is free) instead of overwriting a buffer that is still being read.  Without on_batch the sink
    only counts and recycles two buffers."""

    CONST_PLANES = {"flow01": -0.0, "back_flow01": 0.0}  # channel 1 (y) of these tensors

    def __init__(self, on_batch: Optional[Callable] = None, byte_images: Optional[bool] = None, widen_threads: int = 4,
                 const_planes: Optional[bool] = None):
        import os
        import threading

        self.on_batch = on_batch
        pcie_bound = int(os.environ.get("LOCAL_WORLD_SIZE", "1")) <= 2
        if byte_images is None:
            byte_images = pcie_bound
        self.byte_images = bool(byte_images)
        self.const_planes = pcie_bound if const_planes is None else bool(const_planes)
        self.frames = 0
        self.bytes = 0               # bytes that crossed PCIe
        self.buffers_allocated = 0
        self.fallback_batches = 0    # batches whose image check failed (delivered from the float planes)
        self._inflight = []          # oldest first, at most 2
        self._free = []              # released page-locked float buffers
        self._free8 = []             # page-locked byte staging buffers
        self._lock = threading.Lock()
        self._stream = None
        self._pool = None
        self._widen_threads = max(1, min(int(widen_threads), len(os.sched_getaffinity(0))))

    def _take(self, shape):
        with self._lock:
            for k, h in enumerate(self._free):
                if h.shape[0] >= shape[0] and tuple(h.shape[1:]) == tuple(shape[1:]):
                    return self._free.pop(k)
            self._free.clear()  # another frame size: drop the old buffers
        self.buffers_allocated += 1
        return torch.empty(shape, dtype=torch.float32, pin_memory=True)

    def _take8(self, shape):
        for k, h in enumerate(self._free8):
            if h.shape[0] >= shape[0] and tuple(h.shape[1:]) == tuple(shape[1:]):
                return self._free8.pop(k)
        self._free8.clear()
        return torch.empty(shape, dtype=torch.uint8, pin_memory=True)

    def _release(self, host):
        with self._lock:
            self._free.append(host)

    def _deliver_oldest(self):
        from . import ops

        entry = self._inflight.pop(0)
        host, event, idx_list, n = entry["host"], entry["event"], entry["idx"], entry["n"]
        event.synchronize()
        if entry["const"]:
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor

                self._pool = ThreadPoolExecutor(max_workers=self._widen_threads)
            const_jobs = [self._pool.submit(ops.host_stream_fill, host[b, c], v) for b in range(n) for (c, v) in entry["const"]]
        else:
            const_jobs = []
        if entry["host8"] is not None:
            host8, spans = entry["host8"], entry["spans"]
            if int(entry["flag_host"][0]) == 0:
                if self._pool is None:
                    from concurrent.futures import ThreadPoolExecutor

                    self._pool = ThreadPoolExecutor(max_workers=self._widen_threads)
                jobs = [(b, c8, c0, c) for b in range(n) for (c8, c0, c) in spans]
                list(self._pool.map(lambda j: ops.host_widen_u8(host8[j[0], j[1]:j[1] + j[3]], host[j[0], j[2]:j[2] + j[3]]), jobs))
            else:  # some image value is not a uint8: deliver the float planes (kept alive for this case)
                self.fallback_batches += 1
                self.byte_images = False
                for t, c0 in entry["float_images"]:
                    ops.scatter_channels_to_host(t, host, c0)
                torch.cuda.current_stream(entry["float_images"][0][0].device).synchronize()
            self._free8.append(host8)
        for f in const_jobs:
            f.result()
        entry["float_images"] = None
        if self.on_batch is None:
            self._release(host)
            return
        done = []

        def release():
            if not done:
                done.append(True)
                self._release(host)

        self.on_batch(idx_list, host[:n].numpy(), release)

    def __call__(self, idx_list, res):
        from . import ops
        from .preprocess import GROUP_CHANNELS

        first = res[GROUP_CHANNELS[0]]
        dev, B = first.device, first.shape[0]
        H, W = first.shape[-2:]
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        while len(self._inflight) >= 2:  # at most two batches in flight: the older one has landed by now
            self._deliver_oldest()
        ctot = sum(res[n].shape[1] for n in GROUP_CHANNELS)
        host = self._take((B, ctot, H, W))
        use_bytes = self.byte_images
        images = [(n, res[n]) for n in GROUP_CHANNELS if n.startswith("img")] if use_bytes else []
        host8 = flag = flag_host = None
        packed = {}
        if images:
            # narrow the image tensors on the compute stream (they are ready there), verified by one device flag for the batch
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            for name, t in images:
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                packed[name] = ops.pack_u8(t, flag)
            host8 = self._take8((B, sum(t.shape[1] for _, t in images), H, W))
            flag_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        c0 = c8 = 0
        spans, float_images, crossed, const = [], [], 0, []
        with torch.cuda.stream(self._stream):
            for name in GROUP_CHANNELS:
                t = res[name]
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                if name in packed:
                    u = packed[name]
                    ops.scatter_channels_to_host(u, host8, c8, stream=self._stream)
                    u.record_stream(self._stream)
                    spans.append((c8, c0, t.shape[1]))
                    float_images.append((t, c0))
                    c8 += t.shape[1]
                    crossed += u.numel()
                elif self.const_planes and name in self.CONST_PLANES:
                    ops.scatter_channels_to_host(t, host, c0, stream=self._stream, n_channels=1)  # x only; y is written at delivery
                    t.record_stream(self._stream)
                    const.append((c0 + 1, self.CONST_PLANES[name]))
                    crossed += t.numel() * 2
                else:
                    ops.scatter_channels_to_host(t, host, c0, stream=self._stream)
                    t.record_stream(self._stream)
                    crossed += t.numel() * 4
                c0 += t.shape[1]
            if flag is not None:
                flag_host.copy_(flag, non_blocking=True)
                flag.record_stream(self._stream)
        event = torch.cuda.Event()
        event.record(self._stream)
        self._inflight.append(dict(host=host, event=event, idx=list(idx_list), n=B, host8=host8, spans=spans, flag_host=flag_host,
                                   float_images=float_images if images else None, const=const))
        self.frames += len(idx_list)
        self.bytes += crossed

    def flush(self):
        """Wait for the outstanding copies and hand their batches over (oldest first)."""
        while self._inflight:
            self._deliver_oldest()
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None


def run_sweep(indices: Sequence[int], load_frame: Callable[[int], tuple], device, batch: int = 32, epoch: int = 0,
              dataset_len: int = 0, inpaint: Optional[Callable] = None, sink: Optional[Callable] = None,
              counters: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's per-frame driver loop (preprocess.py:551-561 + the group part of PreprocessPlusAugment.forward,
    :341-447) over this rank's `indices`, `batch` frames at a time on `device`.

    load_frame(idx) -> (img[3,H,W] float32, raw_depth[1,H,W] float32) numpy arrays (all frames of a sweep share H x W).
    For every frame the generator is reseeded with frame_seed(idx, epoch, dataset_len) and the disparity scale and the
    camera pose are drawn in the reference's order, so a frame's result does not depend on batch size, shard or rank.
    sink(idx_list, results_dict) receives each batch's device tensors (e.g. to stage them to pinned host memory).
    Returns the counter block (frames, pairs, hit/hole/collision/dropped/tie pixel counts) of this rank."""
    from . import ops, synthesis

    dev = torch.device(device)
    if counters is None:
        counters = ops.new_counters(dev)
    import os
    from concurrent.futures import ThreadPoolExecutor

    stage = {}  # page-locked staging of the input batch: frames are copied in once, the H2D copy is asynchronous
    # the staging copies (5 MB per frame) run on a few threads (torch's copy_ releases the GIL); under torchrun every rank is
    # confined to its own core slice (bind_rank_cores), so the pool never exceeds it
    copiers = ThreadPoolExecutor(max_workers=max(1, min(4, len(os.sched_getaffinity(0)))))
    try:
        for idx_list in batches(indices, batch):
            frames = [load_frame(i) for i in idx_list]
            n = len(frames)
            key = (frames[0][0].shape, frames[0][0].dtype, frames[0][1].shape, frames[0][1].dtype)
            if stage.get("key") != key or stage["img"][0].shape[0] < n:
                stage = {"key": key, "flip": 0,
                         "img": [torch.empty((max(batch, n),) + frames[0][0].shape, dtype=torch.from_numpy(frames[0][0]).dtype, pin_memory=True) for _ in range(2)],
                         "raw": [torch.empty((max(batch, n),) + frames[0][1].shape, dtype=torch.from_numpy(frames[0][1]).dtype, pin_memory=True) for _ in range(2)],
                         "ev": [None, None]}
            slot = stage["flip"]
            stage["flip"] ^= 1
            if stage["ev"][slot] is not None:
                stage["ev"][slot].synchronize()  # the previous H2D copy out of this staging slot has finished
            h_img, h_raw = stage["img"][slot], stage["raw"][slot]

            def put(k, h_img=h_img, h_raw=h_raw, frames=frames):
                h_img[k].copy_(torch.from_numpy(frames[k][0]))
                h_raw[k].copy_(torch.from_numpy(frames[k][1]))

            pending = [copiers.submit(put, k) for k in range(n)]
            # while the copies run: this batch's random draws, in the reference's per-frame order (preprocess.py:555,356,372)
            h, w = frames[0][0].shape[-2:]
            s_host, cam_host, _ = synthesis.frame_draws_batch([frame_seed(i, epoch, dataset_len) for i in idx_list], (h, w))
            for f in pending:
                f.result()
            img = h_img[:n].to(dev, non_blocking=True)
            raw = h_raw[:n].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            stage["ev"][slot] = ev
            if os.environ.get("OFD_SWEEP_PAGEABLE_UPLOADS") == "1":  # A/B knob: the synchronising uploads of before
                sBf, cam = s_host.to(dev), cam_host.to(dev)
            else:
                sBf = s_host.pin_memory().to(dev, non_blocking=True)   # page-locked staging: no stream synchronisation per batch
                cam = cam_host.pin_memory().to(dev, non_blocking=True)
            with torch.cuda.device(dev):
                depth = ops.normalize_depth(raw)                                    # :355
            res = synthesis.synthesize_group(img, depth, sBf, cam, inpaint=inpaint, counters=counters)
            counters[_lib.CNT_FRAMES] += len(idx_list)
            counters[_lib.CNT_PAIRS] += 5 * len(idx_list)
            if sink is not None:
                sink(idx_list, res)
    finally:
        copiers.shutdown(wait=True)
    if sink is not None and hasattr(sink, "flush"):
        sink.flush()
    return counters
