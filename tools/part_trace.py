#!/usr/bin/env python
"""Per-step timeline of the fused partitioned path on real peers (torchrun, one rank per GPU): for each return path and mode
one traced batch (BLIGHT_PART_TRACE=1: csrc/part_session.cu prints, per rank, when every dispatch / wait / lookup / scatter
ended), next to an NCCL all-to-all of the same id volume as a yardstick of what NVLink gives. Diagnostic, not a test.
Environment: TRACE_GENOME (default 1 G), TRACE_READS (per rank, 4 M), TRACE_SUB (64 M), TRACE_RETURNS (stream,pull)."""
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blight_b200 import api, synth  # noqa: E402
from blight_b200 import dist as bdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    gl = int(os.environ.get("TRACE_GENOME", 1_000_000_000))
    n_reads = int(os.environ.get("TRACE_READS", 4_000_000))
    sub = int(os.environ.get("TRACE_SUB", 64 << 20))
    returns = os.environ.get("TRACE_RETURNS", "stream,pull").split(",")
    orders = os.environ.get("TRACE_ORDERS", "serial").split(",")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wd = [None]
    if rank == 0:
        wd = [tempfile.mkdtemp(prefix="blight_trace_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)]
    dist.broadcast_object_list(wd, src=0)
    g = synth.random_genome(gl, seed=42)
    flat = None
    if rank == 0:
        st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
        flat = api.FlatIndex.build_spans(g, st, ln, 31, 9, 10, 3, 6, threads=os.cpu_count() or 1)
    part = bdist.PartitionedSet.from_full(flat, local, wd[0])
    del flat
    d_genome = torch.from_numpy(g).to(dev)
    del g
    bases = synth.torch_simulate_reads(d_genome, n_reads, 150, 0.01, 0.5, seed=144 + rank)
    del d_genome
    roff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 150
    koff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 120
    total = n_reads * 120

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # yardstick: the ids of one batch (4 bytes per k-mer) as ONE NCCL all-to-all, and as one per sub-batch
    for label, n in (("a2a_whole_batch", total), ("a2a_one_sub_batch", int(total * sub / bases.numel()))):
        per = (n // world) // 4 * 4
        src = torch.empty(per * world, dtype=torch.int32, device=dev)
        dst = torch.empty_like(src)
        ms = timed(lambda: dist.all_to_all_single(dst, src))
        if rank == 0:
            print(json.dumps({label: {"bytes_per_gpu": per * world * 4, "ms": ms, "GB_per_s_per_gpu_sent_to_peers": per * (world - 1) * 4 / ms / 1e6}}), flush=True)
        del src, dst
    for order in orders:
        for rp in returns:
            part.enable_fused(sub_positions=sub, ids_capacity=total, order=order, return_path=rp)
            ids_ms = timed(lambda: part.query_reads_fused(bases, roff, koff, total, check_overflow=False))
            cnt_ms = timed(lambda: part.query_reads_fused(bases, roff, want_ids=False, check_overflow=False))
            if rank == 0:
                print(json.dumps({"order": order, "return": rp, "sub": part._sub, "ids_ms": ids_ms, "counting_ms": cnt_ms}), flush=True)
            torch.cuda.synchronize(); dist.barrier()
            os.environ["BLIGHT_PART_TRACE"] = "1"
            sys.stderr.write(json.dumps({"traced": {"order": order, "return": rp}}) + "\n"); sys.stderr.flush()
            part.query_reads_fused(bases, roff, koff, total, check_overflow=False)
            torch.cuda.synchronize(); dist.barrier()
            if rp == returns[0]:
                part.query_reads_fused(bases, roff, want_ids=False, check_overflow=False)
                torch.cuda.synchronize(); dist.barrier()
            os.environ["BLIGHT_PART_TRACE"] = "0"
    part.disable_fused()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        import shutil
        shutil.rmtree(wd[0], ignore_errors=True)


if __name__ == "__main__":
    main()
