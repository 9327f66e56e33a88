#!/usr/bin/env python
"""Multi-GPU parity + throughput check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Replica mode and partition mode must both return, for every rank's shard of the reads, exactly the ids a single GPU
holding the whole index returns. Prints one JSON line from rank 0."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    genome_len = int(os.environ.get("BLIGHT_CHECK_GENOME", 20_000_000))
    n_reads = int(os.environ.get("BLIGHT_CHECK_READS", 2_000_000))
    m, n, b = 9, 10, 6
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = synth.random_genome(genome_len, seed=42)
    st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
    flat = api.FlatIndex.build_spans(g, st, ln, 31, m, n, 3, b, threads=0)
    N = flat.info()["number_kmer"]

    # the whole batch, identical on every rank
    d_genome = torch.from_numpy(g).to(dev)
    bases = synth.torch_simulate_reads(d_genome, n_reads, 150, 0.01, 0.5, seed=44)
    roff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 150
    lo, hi = bdist.shard_range(n_reads, rank, world)

    # ground truth for my shard: this GPU, whole index
    rep = bdist.ReplicaSet(flat, local)
    ids_rep, ctr_rep = rep.query_reads_sharded(bases, roff, want_ids=True)
    torch.cuda.synchronize()
    koff_all = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 120
    if rank == 0:
        full_ids, full_ctr = rep.index.query_reads(bases, roff, koff_all, n_reads * 120)
        torch.cuda.synchronize()
        assert torch.equal(full_ids[lo * 120:hi * 120], ids_rep)
        assert torch.equal(full_ctr.cpu(), ctr_rep.cpu()), (full_ctr, ctr_rep)

    # partition mode
    wd = [None]
    if rank == 0:
        wd = [tempfile.mkdtemp(prefix="blight_part_")]
    dist.broadcast_object_list(wd, src=0)
    part = bdist.PartitionedSet.from_full(flat if rank == 0 else None, local, wd[0])
    my_bases = bases[lo * 150:hi * 150].clone()
    my_roff = (roff[lo:hi + 1] - lo * 150).contiguous()
    my_koff = (koff_all[lo:hi + 1] - lo * 120).contiguous()
    total = (hi - lo) * 120
    ids_part, ctr_part = part.query_reads(my_bases, my_roff, my_koff, total)
    torch.cuda.synchronize()
    ok = bool(torch.equal(ids_part, ids_rep))
    # timing of partition mode (whole pipeline incl. both all-to-alls), max over ranks
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        part.query_reads(my_bases, my_roff, my_koff, total)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    tot = ctr_part.clone()
    dist.all_reduce(tot)
    if rank == 0:
        print(json.dumps({"world": world, "index_kmers": N, "reads": n_reads, "partition_ids_equal_replica": bool(okt.item()),
                          "partition_ms": float(t.item()), "partition_kmers_per_s": n_reads * 120 / (float(t.item()) * 1e-3),
                          "found_total": int(tot[0]), "replica_found_total": int(ctr_rep[0]),
                          "local_index_bytes": part.index.info["device_bytes"], "cuts": part.plan.cuts}))
        assert bool(okt.item()) and int(tot[0]) == int(ctr_rep[0])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
