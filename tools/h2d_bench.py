#!/usr/bin/env python
"""Pinned host -> device copy bandwidth of this box, one GPU or all of them AT THE SAME TIME (run under torchrun, one rank per
GPU): the bound of bench.py's e2e once the kernels are faster than the copies. Also the host side of the story: how fast the
box's cores read memory (numpy sum over a buffer larger than the caches, all ranks at once), alone and while the copies run.
    python tools/h2d_bench.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/h2d_bench.py"""
import json
import os
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1_580_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.zero_()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hn = h.numpy().view(np.uint64)[: n // 8]


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def copies(reps=5):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return n / dt / 1e9


def host_read(reps=3):
    hn.sum()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        hn.sum()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return n / dt / 1e9


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]
    return [x]


h2d = gather(copies())
cpu = gather(host_read())
# both at once: a reader thread per rank while the copies run
stop = False
acc = [0]


def reader():
    while not stop:
        hn[: n // 64].sum()
        acc[0] += n // 8


barrier()
th = threading.Thread(target=reader)
t0 = time.perf_counter()
th.start()
both_h2d = copies(reps=5)
stop = True
th.join()
both_cpu = acc[0] / (time.perf_counter() - t0) / 1e9
both = gather(both_h2d)
both_c = gather(both_cpu)
if rank == 0:
    print(json.dumps({"gpus": world, "host_cores": os.cpu_count(), "bytes_per_copy": n,
                      "h2d_GB_per_s_per_gpu": h2d, "h2d_GB_per_s_total": sum(h2d),
                      "host_read_GB_per_s_per_rank_one_thread": cpu, "host_read_GB_per_s_total": sum(cpu),
                      "with_a_reader_thread_per_rank": {"h2d_GB_per_s_total": sum(both), "host_read_GB_per_s_total": sum(both_c)}}))
if world > 1:
    dist.destroy_process_group()
