"""Multi-GPU sweep of independent frames (preprocess.py:511-515,540-561).

Frames are independent, so the path shards by image index with NO collective on the hot path: one process per GPU,
each runs its index range through the fused kernels; at the end of the run the ranks sum their counter blocks
(frames, pairs, hit/hole/collision/dropped pixel counts) with one all-reduce (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

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
