#!/usr/bin/env python
"""construct_index: the GPU builder (csrc/gpu_builder.cu) next to the host builder (csrc/builder.cpp) on the synthetic unitig
graphs of BASELINE.json — same flat image (checked), wall clock of either call, and the time the GPU builder spends between
its first H2D copy and its last kernel. One JSON line per shape. BUILD_GENOMES=100000000,1000000000 picks the sizes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blight_b200 import api, synth  # noqa: E402

sizes = [int(x) for x in os.environ.get("BUILD_GENOMES", "100000000").split(",")]
shapes = [(7, 5, 6), (9, 10, 6), (11, 12, 6)]
g0 = synth.random_genome(2_000_000, seed=1)
s0, l0 = synth.cut_unitigs(g0, 31, 2000, seed=2)
api.FlatIndex.build_gpu(g0, s0, l0, 31, 7, 5, 3, 6)  # warm-up: context, kernels
for G in sizes:
    g = synth.random_genome(G, seed=42)
    st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
    for m, n, b in shapes:
        t0 = time.perf_counter()
        host = api.FlatIndex.build_spans(g, st, ln, 31, m, n, 3, b, threads=os.cpu_count() or 1)
        t_host = time.perf_counter() - t0
        t0 = time.perf_counter()
        dev = api.FlatIndex.build_gpu(g, st, ln, 31, m, n, 3, b)
        t_gpu = time.perf_counter() - t0
        same = dev.equals(host)
        print(json.dumps({"kmers": dev.info()["number_kmer"], "k": 31, "m": m, "n": n, "b": b, "host_cores": os.cpu_count(),
                          "host_builder_s": t_host, "gpu_builder_wall_s": t_gpu, "gpu_builder_device_s": dev.gpu_build_seconds,
                          "same_image": bool(same), "speedup_wall": t_host / t_gpu}), flush=True)
        del host, dev
