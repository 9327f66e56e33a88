#!/usr/bin/env python
"""Loop-back tuning of the fused partitioned path on ONE GPU (world = 1: the rank dispatches to itself, so the two kernels,
their ordering and the return path run exactly as on a box, minus NVLink): sub-batch size, kernel order, CTA split of the
overlap order, return path — against the one-kernel read path on the same index. One JSON line per variant."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blight_b200 import api, synth  # noqa: E402
from blight_b200 import dist as bdist  # noqa: E402

genome_len = int(os.environ.get("TUNE_GENOME", 100_000_000))
n_reads = int(os.environ.get("TUNE_READS", 4_000_000))
m, n, b = (int(x) for x in os.environ.get("TUNE_SHAPE", "9,10,6").split(","))
reps = int(os.environ.get("TUNE_REPS", 5))
torch.cuda.set_device(0)
g = synth.random_genome(genome_len, seed=42)
st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
flat = api.FlatIndex.build_spans(g, st, ln, 31, m, n, 3, b, threads=os.cpu_count() or 1)
bases = synth.torch_simulate_reads(torch.from_numpy(g).cuda(), n_reads, 150, 0.01, 0.5, seed=44)
roff = torch.arange(0, n_reads + 1, device="cuda", dtype=torch.int64) * 150
koff = torch.arange(0, n_reads + 1, device="cuda", dtype=torch.int64) * 120
total = n_reads * 120
plan = bdist.PartitionPlan([0, flat.info()["n_mphf"]], 2 * m - 1 - n)
ps = bdist.PartitionedSet(plan, flat, 0, 31, m)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


scratch = torch.empty(total, dtype=torch.int64, device="cuda")
want, _ = ps.index.query_reads(bases, roff, koff, total)
torch.cuda.synchronize()
want = want[:total].clone()
one_ids = timed(lambda: ps.index.query_reads(bases, roff, koff, total, ids=scratch))
one_cnt = timed(lambda: ps.index.query_reads(bases, roff, want_ids=False))
del scratch
print(json.dumps({"variant": "one kernel (k_reads_sk)", "ids_ms": one_ids, "counting_ms": one_cnt}), flush=True)
variants = []
if os.environ.get("TUNE_SUBS"):
    variants = [(int(x) << 20, "serial", None, "stream") for x in os.environ["TUNE_SUBS"].split(",")]
for sub in (() if os.environ.get("TUNE_SUBS") else (32 << 20, 64 << 20, 128 << 20, 256 << 20)):
    variants.append((sub, "serial", None, "stream"))
if not os.environ.get("TUNE_SUBS"):
    variants += [(64 << 20, "ahead", None, "stream"), (64 << 20, "serial", None, "direct"), (64 << 20, "serial", None, "pull"), (32 << 20, "serial", None, "pull")]
if os.environ.get("TUNE_OVERLAP"):
    for split in ("2,2", "3,1", "3,2", "2,1", "4,1"):
        variants.append((64 << 20, "overlap", split, "stream"))
    variants.append((128 << 20, "overlap", "3,1", "stream"))
for sub, order, split, ret in variants:
    if split:
        os.environ["BLIGHT_PART_SPLIT"] = split
    ps.enable_fused(sub_positions=sub, ids_capacity=total, order=order, return_path=ret,
                    records_per_position=float(os.environ["TUNE_RPP"]) if os.environ.get("TUNE_RPP") else None)
    ids, _ = ps.query_reads_fused(bases, roff, koff, total)
    torch.cuda.synchronize()
    ok = bool(torch.equal(ids, want))
    ids_ms = timed(lambda: ps.query_reads_fused(bases, roff, koff, total, check_overflow=False))
    cnt_ms = timed(lambda: ps.query_reads_fused(bases, roff, want_ids=False, check_overflow=False))
    print(json.dumps({"sub": ps._sub, "order": order, "split": split, "return": ret, "ids_ms": ids_ms, "counting_ms": cnt_ms, "ids_equal": ok,
                      "overflow": ps.overflowed(), "ids_vs_one_kernel": one_ids / ids_ms, "counting_vs_one_kernel": one_cnt / cnt_ms}), flush=True)
