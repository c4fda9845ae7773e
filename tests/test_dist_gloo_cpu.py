"""N>1 path on CPU: world_size-2 gloo run of the sweep plumbing (sharding + end-of-run counter reduction)."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opticalflowfromdepth_b200 import _lib, sweep

    mine = sweep.shard_range(n_frames, world, rank)
    counters = torch.zeros(_lib.CNT_SLOTS, dtype=torch.int64)
    for batch in sweep.batches(mine, 4):
        counters[_lib.CNT_FRAMES] += len(batch)
        counters[_lib.CNT_PAIRS] += 5 * len(batch)
        counters[_lib.CNT_HIT] += sum(batch)  # any rank-dependent payload
    total = sweep.reduce_counters(counters)
    q.put((rank, len(mine), total))
    dist.destroy_process_group()


def test_two_rank_sweep_counters():
    world, n = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[1] for r in res) == n
    for _, _, total in res:
        assert total["frames"] == n and total["pairs"] == 5 * n and total["hit"] == n * (n - 1) // 2
