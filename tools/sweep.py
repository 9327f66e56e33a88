#!/usr/bin/env python
"""BASELINE.json configs[2]: b sweep 0-8 and m sweep {7,9,11} on the 100 M-k-mer synthetic index, one GPU.
For every shape: throughput (id mode and counting mode, CUDA events, inputs resident) and a parity check of a read sample against the
oracle (C port). Prints one JSON line per shape; profiles/ keeps the table."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (checker only)
from blight_b200 import api, synth  # noqa: E402

genome_len = int(os.environ.get("SWEEP_GENOME", 100_000_000))
n_reads = int(os.environ.get("SWEEP_READS", 4_000_000))
sample = int(os.environ.get("SWEEP_SAMPLE", 3000))
shapes = [(7, 5, b) for b in range(0, 9)] + [(9, 5, 6), (11, 5, 6), (9, 17, 6)]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
g = synth.random_genome(genome_len, seed=42)
st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
d_genome = torch.from_numpy(g).to(dev)
bases = synth.torch_simulate_reads(d_genome, n_reads, 150, 0.01, 0.5, seed=44)
roff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 150
koff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * 120
ids = torch.empty(n_reads * 120, dtype=torch.int64, device=dev)
hb = bases[: sample * 150].cpu().numpy()
ho = np.arange(sample + 1, dtype=np.uint64) * np.uint64(150)
for (m, n, b) in shapes:
    t0 = time.time()
    flat = api.FlatIndex.build_spans(g, st, ln, 31, m, n, min(n, 3), b, threads=os.cpu_count() or 1)
    tb = time.time() - t0
    idx = flat.upload(0)
    ctr = torch.zeros(4, dtype=torch.int64, device=dev)
    for _ in range(2):
        idx.query_reads(bases, roff, koff, n_reads * 120, ids=ids, ctr=ctr)
    torch.cuda.synchronize()
    ctr.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        idx.query_reads(bases, roff, koff, n_reads * 120, ids=ids, ctr=ctr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        idx.query_reads(bases, roff, want_ids=False, ctr=ctr)
    e1.record()
    torch.cuda.synchronize()
    ms_count = e0.elapsed_time(e1) / reps
    ctr.zero_()
    idx.query_reads(bases, roff, koff, n_reads * 120, ids=ids, ctr=ctr)
    torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as td:
        blob = os.path.join(td, "x.blflat")
        flat.save(blob)
        port = oracle.CPort(blob)
        want, wctr = port.query_reads(hb, ho)
    got = ids[: sample * 120].cpu().numpy()
    info = idx.info
    print(json.dumps({"k": 31, "m": m, "n": n, "b": b, "kmers_per_s": n_reads * 120 / (ms * 1e-3), "ms": ms,
                      "count_kmers_per_s": n_reads * 120 / (ms_count * 1e-3), "count_ms": ms_count,
                      "parity_sample_kmers": int(len(want)), "parity_ok": bool(np.array_equal(got, want)),
                      "oracle_found": int(wctr[0]), "found_fraction": float(ctr[0]) / float(ctr[2]),
                      "device_MB": info["device_bytes"] / 1e6, "bits_per_kmer": 8.0 * info["device_bytes"] / info["number_kmer"],
                      "build_s": tb}), flush=True)
    del idx, flat
