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
    host: one torch.cat on the device, one D2H copy into page-locked memory per batch, issued on a side stream into one of two
    pinned slots, so the copy of batch k overlaps the host-side preparation and the kernels of batch k+1.
    `on_batch(idx_list, array[B,44,H,W])` receives the host array once its copy has landed (when its slot is reused, or at
    flush()) - e.g. to hand it to preprocess.NpzWriter; without it the sink only counts."""

    def __init__(self, on_batch: Optional[Callable] = None):
        self.on_batch = on_batch
        self.frames = 0
        self.bytes = 0
        self._slots = [None, None]   # (pinned tensor, event, idx_list, n) per slot
        self._k = 0
        self._stream = None

    def _deliver(self, slot):
        entry = self._slots[slot]
        if entry is None:
            return
        host, event, idx_list, n = entry
        event.synchronize()
        self._slots[slot] = (host, event, None, 0)
        if idx_list is not None and self.on_batch is not None:
            self.on_batch(idx_list, host[:n].numpy())

    def __call__(self, idx_list, res):
        from .preprocess import GROUP_CHANNELS

        stack = torch.cat([res[n].float() for n in GROUP_CHANNELS], 1)
        dev = stack.device
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        slot = self._k & 1
        self._k += 1
        self._deliver(slot)  # the slot's previous copy has landed and is handed over before the buffer is reused
        entry = self._slots[slot]
        host = entry[0] if entry is not None and entry[0].shape[1:] == stack.shape[1:] and entry[0].shape[0] >= stack.shape[0] else None
        if host is None:
            host = torch.empty(stack.shape, dtype=stack.dtype, pin_memory=True)
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            host[:stack.shape[0]].copy_(stack, non_blocking=True)
            event = torch.cuda.Event()
            event.record(self._stream)
        stack.record_stream(self._stream)
        self._slots[slot] = (host, event, list(idx_list), stack.shape[0])
        self.frames += len(idx_list)
        self.bytes += stack.numel() * stack.element_size()

    def flush(self):
        """Wait for the outstanding copies and hand their batches over (oldest first)."""
        for slot in (self._k & 1, (self._k + 1) & 1):
            self._deliver(slot)


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
    from . import geometry, ops, synthesis

    dev = torch.device(device)
    if counters is None:
        counters = ops.new_counters(dev)
    stage = {}  # page-locked staging of the input batch: frames are copied in once, the H2D copy is asynchronous
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
        for k, f in enumerate(frames):
            h_img[k].copy_(torch.from_numpy(f[0]))
            h_raw[k].copy_(torch.from_numpy(f[1]))
        img = h_img[:n].to(dev, non_blocking=True)
        raw = h_raw[:n].to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        stage["ev"][slot] = ev
        h, w = img.shape[-2:]
        K, inv_K = synthesis.Plausible.K((h, w))
        s_vals, cams = [], []
        for i in idx_list:
            synthesis.set_seed(frame_seed(i, epoch, dataset_len))
            s_vals.append(float(synthesis.Convert.disparity_scale()))          # preprocess.py:356 (first draw)
            T1, _, _ = synthesis.Plausible.random_motion(1. / 36., 1. / 36., 0.1, 0.1)  # :372 -> :277
            cams.append(geometry.camera_constants(K, inv_K, T1))
        sBf = torch.tensor(s_vals, dtype=torch.float32, device=dev)
        cam = torch.cat(cams).to(dev)
        with torch.cuda.device(dev):
            depth = ops.normalize_depth(raw)                                    # :355
        res = synthesis.synthesize_group(img, depth, sBf, cam, inpaint=inpaint, counters=counters)
        counters[_lib.CNT_FRAMES] += len(idx_list)
        counters[_lib.CNT_PAIRS] += 5 * len(idx_list)
        if sink is not None:
            sink(idx_list, res)
    if sink is not None and hasattr(sink, "flush"):
        sink.flush()
    return counters
