"""Host-link probe for the e2e leg: does the box's PCIe + host memory scale with the number of GPUs?

    python tools/probe_pcie.py                                   one GPU
    torchrun --nproc-per-node N tools/probe_pcie.py              one rank per GPU, all ranks measure AT THE SAME TIME

No repo kernel is in the loop for the first block: plain pinned cudaMemcpy through torch on two streams.
Per mode the table shows GB/s per rank (min / max over ranks) and the node total:
    d2h            device -> pinned host only
    h2d            pinned host -> device only
    both           both directions at once (GB/s per direction)
    both+fill T    both directions + T host threads per rank streaming 0.0f into pinned memory with non-temporal stores
                   (ofd_host_stream_fill: the store pattern of the pipeline's host threads); fill GB/s listed separately
    fill T         the host threads alone (host memory write bandwidth, no DMA)
Then the pair pipeline itself (ofd_pair_pipeline_run, float32 contract) per rank at the bench's e2e size, with and without
OFD_PIPE_KEEP_CONST_PLANES, with img1 as device-verified bytes or as float planes, and with the masks as float planes (no host threads at all).
`--affinity` pins every rank to its own core slice first (sweep.bind_rank_cores)."""
import argparse
import ctypes as C
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from opticalflowfromdepth_b200 import _lib, ops, sweep  # noqa: E402

H, W = 480, 640


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--affinity", action="store_true")
    ap.add_argument("--mb", type=int, default=256, help="MiB per copy")
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--no-pipeline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = sweep.bind_rank_cores(local, world) if args.affinity else sorted(os.sched_getaffinity(0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [float(o.item()) for o in out]
        return [x]

    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    say(f"# ranks {world}, host cores visible to rank 0: {len(cores)} of {os.cpu_count()} ({'pinned per rank' if args.affinity else 'no affinity'}), "
        f"{args.mb} MiB per copy x {args.reps} reps, GPU {torch.cuda.get_device_name(dev)}")
    n = args.mb << 20
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_f = [torch.empty(n // 16, dtype=torch.float32).pin_memory() for _ in range(8)]
    h_a.zero_(), h_b.zero_()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    lib = _lib.load()

    def run_mode(d2h, h2d, fill_threads):
        stop = threading.Event()
        filled = [0] * max(fill_threads, 1)

        def filler(t):
            buf = h_f[t % len(h_f)]
            ptr, cnt = C.c_void_p(buf.data_ptr()), C.c_size_t(buf.numel())
            while not stop.is_set():
                lib.ofd_host_stream_fill(ptr, cnt, C.c_float(0.0))
                filled[t] += buf.numel() * 4

        def copies():
            if d2h:
                with torch.cuda.stream(s1):
                    h_a.copy_(d_a, non_blocking=True)
            if h2d:
                with torch.cuda.stream(s2):
                    d_b.copy_(h_b, non_blocking=True)

        copies()
        torch.cuda.synchronize(dev)
        barrier()
        thr = [threading.Thread(target=filler, args=(t,)) for t in range(fill_threads)]
        for t in thr:
            t.start()
        t0 = time.perf_counter()
        if d2h or h2d:
            for _ in range(args.reps):
                copies()
            torch.cuda.synchronize(dev)
        else:
            time.sleep(1.0)
        dt = time.perf_counter() - t0
        stop.set()
        for t in thr:
            t.join()
        barrier()
        return n * args.reps / dt / 1e9 if (d2h or h2d) else 0.0, sum(filled) / dt / 1e9

    def report(name, d2h, h2d, T=0):
        dma, fill = run_mode(d2h, h2d, T)
        g, f = gather(dma), gather(fill)
        ndir = (1 if d2h else 0) + (1 if h2d else 0)
        msg = f"{name:14s}"
        if ndir:
            msg += f" DMA per rank per direction {min(g):6.1f} .. {max(g):6.1f} GB/s   node total {sum(g) * ndir:7.1f} GB/s"
        if T:
            msg += f"   host fill per rank {min(f):5.1f} .. {max(f):5.1f}  total {sum(f):6.1f} GB/s"
        say(msg)

    report("d2h", True, False)
    report("h2d", False, True)
    report("both", True, True)
    for T in (1, 2, 4):
        report(f"both+fill {T}", True, True, T)
    for T in (1, 2, 4):
        report(f"fill {T}", False, False, T)
    del h_a, h_b, d_a, d_b, h_f

    if args.no_pipeline:
        return
    rng = np.random.default_rng(rank)
    Fe = args.frames
    h_img = torch.from_numpy(rng.integers(0, 256, (Fe, 3, H, W)).astype(np.float32)).pin_memory()  # uint8-valued, like the reference's loader
    h_dep = torch.from_numpy((rng.random((Fe, 1, H, W)) * 98 + 1).astype(np.float32)).pin_memory()
    h_s = torch.full((Fe,), 47.0)
    h_out = [torch.empty((Fe, c, H, W), dtype=torch.float32).pin_memory() for c in (3, 1, 2, 2, 1, 1)]
    for label, env, keep in (("pipeline default", {}, False),
                             ("pipeline img1 as verified bytes", {"OFD_HOST_IMG_BYTES": "1"}, False),
                             ("pipeline img1 as float planes", {"OFD_HOST_IMG_BYTES": "0"}, False),
                             ("pipeline img1 bytes + keep_const_planes", {"OFD_HOST_IMG_BYTES": "1"}, True),
                             ("pipeline img1 floats + keep_const_planes", {"OFD_HOST_IMG_BYTES": "0"}, True),
                             ("pipeline img1 floats, 2 workers", {"OFD_HOST_IMG_BYTES": "0", "OFD_HOST_WORKERS": "2"}, False),
                             ("pipeline float masks + keep_const (no host stores)", {"OFD_HOST_MASK_BYTES": "0"}, True)):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        pipe = ops.PairPipeline(local, H, W, chunk_frames=8)
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        for _ in range(2):
            pipe.run(h_img, h_dep, h_s, *h_out, keep_const_planes=keep)
        barrier()
        t0 = time.perf_counter()
        K = 6
        for _ in range(K):
            pipe.run(h_img, h_dep, h_s, *h_out, keep_const_planes=keep)
        dt = (time.perf_counter() - t0) / K
        pipe.close()
        barrier()
        g = gather(Fe / dt)
        say(f"{label:52s} pairs/s per rank {min(g):7.0f} .. {max(g):7.0f}   node total {sum(g):8.0f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
