#!/usr/bin/env python
"""Multi-GPU parity + throughput check / BASELINE configs[4] driver (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Every rank holds its own reads. Replica mode (whole index per GPU) is the ground truth and the 1-GPU rate; partition
mode — plain NCCL exchange of (canon, minimizer), and the fused peer-memory path of csrc/part_kernels.cu — must return
exactly the same ids for every rank's reads. Prints one JSON line from rank 0.
Environment: BLIGHT_CHECK_GENOME (bases, default 20 M), BLIGHT_CHECK_READS (per rank, default 2 M), BLIGHT_CHECK_M / _N / _B
(index shape, default 9 / 10 / 6), BLIGHT_CHECK_SUB (positions per sub-batch of the fused path), BLIGHT_CHECK_PLAIN=0 skips
the plain exchange (it needs 20 B of staging per k-mer)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blight_b200 import api, synth  # noqa: E402
from blight_b200 import dist as bdist  # noqa: E402


def timed(fn, reps, dev):
    """ms per call: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    genome_len = int(os.environ.get("BLIGHT_CHECK_GENOME", 20_000_000))
    n_reads = int(os.environ.get("BLIGHT_CHECK_READS", 2_000_000))
    m, n, b = (int(os.environ.get("BLIGHT_CHECK_" + x, d)) for x, d in (("M", 9), ("N", 10), ("B", 6)))
    subs = [int(x) for x in os.environ.get("BLIGHT_CHECK_SUB", str(64 << 20)).split(",")]
    plain = os.environ.get("BLIGHT_CHECK_PLAIN", "1") != "0"
    reps = int(os.environ.get("BLIGHT_CHECK_REPS", 3))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # every rank holds the genome, the whole flat index (replica = ground truth) and its slice on the host for a while:
    # refuse to start rather than drive the box out of memory
    need_gb = world * (genome_len * 14e-9 + 2) + genome_len * 10e-9
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 1e9
    except Exception:
        avail_gb = float("inf")
    if avail_gb < need_gb:
        if rank == 0:
            print(json.dumps({"skipped": f"needs about {need_gb:.0f} GB of host memory, {avail_gb:.0f} GB available"}))
        dist.destroy_process_group()
        return

    # rank 0 builds the index once and cuts it; everybody loads the whole blob (replica = ground truth) and its slice
    wd = [None]
    if rank == 0:
        wd = [tempfile.mkdtemp(prefix="blight_part_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)]
    dist.broadcast_object_list(wd, src=0)
    blob = os.path.join(wd[0], "full.blflat")
    g = synth.random_genome(genome_len, seed=42)
    t0 = time.time()
    flat = None
    if rank == 0:
        st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
        flat = api.FlatIndex.build_spans(g, st, ln, 31, m, n, 3, b, threads=os.cpu_count() or 1)
        flat.save(blob)
    build_s = time.time() - t0
    part = bdist.PartitionedSet.from_full(flat, local, wd[0])
    if rank != 0:
        flat = api.FlatIndex.load(blob)
    N = flat.info()["number_kmer"]
    rep = bdist.ReplicaSet(flat, local)
    del flat

    # this rank's reads
    d_genome = torch.from_numpy(g).to(dev)
    bases = synth.torch_simulate_reads(d_genome, n_reads, 150, 0.01, 0.5, seed=44 + rank)
    del d_genome
    roff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 150
    koff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 120
    total = n_reads * 120

    # ground truth: this GPU, whole index
    ids_rep, ctr_rep = rep.index.query_reads(bases, roff, koff, total)
    torch.cuda.synchronize()
    ids_rep = ids_rep[:total].clone()
    ctr_all = ctr_rep.clone()
    dist.all_reduce(ctr_all)
    scratch = torch.empty(total, dtype=torch.int64, device=dev)
    rep_ids_ms = timed(lambda: rep.index.query_reads(bases, roff, koff, total, ids=scratch), reps, dev)
    rep_cnt_ms = timed(lambda: rep.index.query_reads(bases, roff, want_ids=False), reps, dev)
    del scratch

    out = {"return_path": os.environ.get("BLIGHT_PART_PIPELINE", "session") + ":" + os.environ.get("BLIGHT_PART_RETURN", "stream") + "+" + os.environ.get("BLIGHT_PART_ORDER", "serial"), "world": world, "index_kmers": N, "reads_per_rank": n_reads, "kmers_per_rank": total, "shape": {"m": m, "n": n, "b": b},
           "build_seconds": build_s, "cuts": part.plan.cuts, "replica_index_bytes": rep.index.info["device_bytes"],
           "local_index_bytes": part.index.info["device_bytes"],
           "replica_ids_ms": rep_ids_ms, "replica_ids_kmers_per_s": world * total / (rep_ids_ms * 1e-3),
           "replica_count_ms": rep_cnt_ms, "replica_count_kmers_per_s": world * total / (rep_cnt_ms * 1e-3)}
    ok = True

    # partition mode, fused peer-memory path (one pass per sub-batch size asked for; the last one is the headline)
    ids_buf = torch.empty(total, dtype=torch.int64, device=dev)
    out["fused_ids_equal_replica"] = out["fused_counters_equal_replica"] = True
    out["fused_by_sub"] = {}
    for sub in subs:
        part.enable_fused(sub_positions=sub, ids_capacity=total)
        ids_buf.fill_(-7)
        ids_f, ctr_f = part.query_reads_fused(bases, roff, koff, total, ids=ids_buf)
        torch.cuda.synchronize()
        okf = torch.tensor([1 if torch.equal(ids_f, ids_rep) else 0], device=dev)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        _, ctr_c = part.query_reads_fused(bases, roff, want_ids=False)
        torch.cuda.synchronize()
        out["fused_ids_equal_replica"] &= bool(okf.item())
        out["fused_counters_equal_replica"] &= bool(torch.equal(ctr_f.cpu()[:3], ctr_all.cpu()[:3]) and torch.equal(ctr_c.cpu()[:3], ctr_all.cpu()[:3]))
        stream_mode = os.environ.get("BLIGHT_PART_PIPELINE") == "legacy"  # the session path answers in its own id array: no copy
        f_ids_ms = timed(lambda: part.query_reads_fused(bases, roff, koff, total, ids=ids_buf if stream_mode else None, check_overflow=False), reps, dev)
        f_cnt_ms = timed(lambda: part.query_reads_fused(bases, roff, want_ids=False, check_overflow=False), reps, dev)
        ovf = torch.tensor([1 if part.overflowed() else 0], device=dev)
        dist.all_reduce(ovf, op=dist.ReduceOp.MAX)
        out["fused_by_sub"][str(part._sub)] = {"ids_ms": f_ids_ms, "count_ms": f_cnt_ms, "overflow": bool(ovf.item())}
        ok = ok and not bool(ovf.item())
        out.update({"fused_sub_positions": part._sub, "fused_ids_ms": f_ids_ms, "fused_ids_kmers_per_s": world * total / (f_ids_ms * 1e-3),
                    "fused_count_ms": f_cnt_ms, "fused_count_kmers_per_s": world * total / (f_cnt_ms * 1e-3),
                    "found_total": int(ctr_f[0]), "replica_found_total": int(ctr_all[0])})
    ok = ok and out["fused_ids_equal_replica"] and out["fused_counters_equal_replica"]
    out["fused_ids_vs_one_gpu"] = out["fused_ids_kmers_per_s"] / (total / (rep_ids_ms * 1e-3))
    out["fused_count_vs_one_gpu"] = out["fused_count_kmers_per_s"] / (total / (rep_cnt_ms * 1e-3))
    part.disable_fused()

    # partition mode, plain NCCL exchange of (canon, minimizer) pairs
    if plain:
        ids_p, ctr_p = part.query_reads(bases, roff, koff, total)
        torch.cuda.synchronize()
        okp = torch.tensor([1 if torch.equal(ids_p, ids_rep) else 0], device=dev)
        dist.all_reduce(okp, op=dist.ReduceOp.MIN)
        out["plain_ids_equal_replica"] = bool(okp.item())
        ok = ok and out["plain_ids_equal_replica"]
        del ids_p
        p_ms = timed(lambda: part.query_reads(bases, roff, koff, total), reps, dev)
        out.update({"plain_ids_ms": p_ms, "plain_ids_kmers_per_s": world * total / (p_ms * 1e-3)})

    if rank == 0:
        print(json.dumps(out))
        assert ok, out
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        import shutil
        shutil.rmtree(wd[0], ignore_errors=True)


if __name__ == "__main__":
    main()
