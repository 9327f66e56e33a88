#!/usr/bin/env python
"""End-to-end file_query(path) throughput: a 2-line FASTA file of simulated reads on tmpfs, plain and
gzip, through kmer_Set_Light::file_query's replacement blight_query_file_host (streaming reader -> parallel record cut
-> H2D / kernel overlap). Counters are checked against the in-memory entry point. One JSON line."""
import gzip
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blight_b200 import api, synth  # noqa: E402

genome_len = int(os.environ.get("FQ_GENOME", 100_000_000))
n_reads = int(os.environ.get("FQ_READS", 4_000_000))
n_gz = int(os.environ.get("FQ_GZ_READS", 400_000))
torch.cuda.set_device(0)
g = synth.random_genome(genome_len, seed=42)
st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
flat = api.FlatIndex.build_spans(g, st, ln, 31, 7, 5, 3, 6, threads=os.cpu_count() or 1)
idx = flat.upload(0)
rb = synth.torch_simulate_reads(torch.from_numpy(g).cuda(), n_reads, 150, 0.01, 0.5, seed=44).cpu().numpy()
# FASTA text: ">123456789\n" + 150 bases + "\n" per read, built with numpy
hdr = np.frombuffer(b">read00000\n", dtype=np.uint8)
rec = np.empty((n_reads, len(hdr) + 151), dtype=np.uint8)
rec[:, :len(hdr)] = hdr
rec[:, len(hdr):len(hdr) + 150] = rb.reshape(n_reads, 150)
rec[:, -1] = ord("\n")
text = rec.reshape(-1)
d = tempfile.mkdtemp(prefix="blight_fq_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
plain, gz = os.path.join(d, "reads.fa"), os.path.join(d, "reads.fa.gz")
text.tofile(plain)
with gzip.open(gz, "wb", compresslevel=1) as f:
    f.write(text[: n_gz * rec.shape[1]].tobytes())
want = idx.query_fasta_host(text)
out = {"reads": n_reads, "file_bytes": int(text.size), "kmers": n_reads * 120, "host_cores": os.cpu_count()}
for name, path, nk in (("plain", plain, n_reads * 120), ("gzip", gz, n_gz * 120)):
    idx.query_file_host(path)  # warm-up: page cache, pinned buffers
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        c = idx.query_file_host(path)
    dt = (time.perf_counter() - t0) / reps
    out[name] = {"seconds": dt, "kmers_per_s": nk / dt, "file_GB_per_s": os.path.getsize(path) / dt / 1e9,
                 "found": int(c[0]), "not_found": int(c[1])}
out["plain"]["counters_equal_in_memory_path"] = bool(int(out["plain"]["found"]) == int(want[0]) and int(out["plain"]["not_found"]) == int(want[1]))
os.environ["BLIGHT_FILE_QUERY"] = "whole"
t0 = time.perf_counter()
c = idx.query_file_host(plain)
out["plain_whole_file_in_memory_first"] = {"seconds": time.perf_counter() - t0, "kmers_per_s": n_reads * 120 / (time.perf_counter() - t0)}
for p in (plain, gz):
    os.remove(p)
os.rmdir(d)
print(json.dumps(out))
