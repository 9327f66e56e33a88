#!/usr/bin/env python
"""bench.py — queried k-mers/s of the batched k-mer query path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]/[3], the one the north-star target is quoted on): a synthetic random-genome
unitig graph with 100 M 31-mers, index k=31 m=7 n=5 s=3 b=6, replicated on every GPU; each GPU queries its own
10 M simulated 150 bp reads (1 % substitutions, half reverse-complemented) = 1.2 G k-mers per step per GPU
(weak scaling, no data-path collective).  A step is one pass of the hot path (read tiles -> 2-bit pack -> rolling
canonical k-mers + minimizers -> MPHF -> positions -> 2^b-window compare -> int64 ids) over that batch.

  value      k-mers/s, all GPUs, inputs resident in HBM, CUDA events, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of the reads + kernels + D2H of the
             counters inside the timed region (file_query semantics: Good / Erroneous counts)
  roofline   algorithmic bytes per k-mer (158 B for b<=6) x k-mers / kernel time vs measured HBM peak
  cpu_baseline   the reference's own query code (oracle/_ref) on the host cores, on a bounded sample of the reads
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--m", type=int, default=7)
    ap.add_argument("--n", type=int, default=5)
    ap.add_argument("--s", type=int, default=3)
    ap.add_argument("--b", type=int, default=6)
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ids-only", action="store_true", help="diagnostic: time only the id (hash) mode")
    ap.add_argument("--count-only", action="store_true", help="diagnostic: time only the counting (bool) mode")
    ap.add_argument("--partition", action="store_true",
                    help="N > 1: minimizer-bucket partitioned index (BASELINE configs[4]) instead of a replica per GPU; the "
                         "exchange is fused into the kernels (peer-memory stores over NVLink)")
    return ap.parse_args()


def b_alg(b: int, k: int = 31) -> float:
    """Algorithmic bytes per queried k-mer: 32 B x (1.65 levels + rank + position + sequence sectors)
    + 1.25 B of ASCII in + 8 B id out."""
    seq_sectors = -(-2 * (k + (1 << b) - 1) // 256)
    return 32.0 * (1.65 + 1 + 1 + seq_sectors) + 1.25 + 8.0


def ncu_traffic(kname: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel, from the committed `ncu --set full`
    capture of this exact workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kname, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload_index(args, rank, world, tmpdir):
    """Rank 0 builds the flat index with the product's host builder and saves it; the others load the blob."""
    from blight_b200 import api, synth
    if world > 1:
        import torch.distributed as dist
        box = [tmpdir]
        dist.broadcast_object_list(box, src=0)  # every rank must look in rank 0's directory
        tmpdir = box[0]
    blob = os.path.join(tmpdir, "bench_index.blflat")
    g = synth.random_genome(args.genome, seed=42)
    t0 = time.time()
    if rank == 0:
        st, ln = synth.cut_unitigs(g, args.k, 2000, seed=43)
        # torchrun exports OMP_NUM_THREADS=1: ask for the host's cores explicitly
        flat = api.FlatIndex.build_spans(g, st, ln, args.k, args.m, args.n, args.s, args.b, threads=os.cpu_count() or 1)
        flat.save(blob)
    if world > 1:
        dist.barrier()
        if rank != 0:
            flat = api.FlatIndex.load(blob)
    return g, flat, blob, time.time() - t0


def run_reference(args, rank):
    """--impl reference: the reference's own CPU query code (oracle/_ref: /root/reference + fixes P1/P2, compiled by
    oracle/build_ref.sh) on the host cores, all threads, on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    import oracle
    from blight_b200 import api, synth
    if not oracle.reference_available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libblight_ref.so was not built (needs /root/reference at build time)"})
        return
    with tempfile.TemporaryDirectory() as td:
        g, flat, blob, _ = build_workload_index(args, 0, 1, td)
        ref = oracle.Reference.from_blob(blob, args.k, args.m)  # reference object holding the identical index
    threads = os.cpu_count() or ref.max_threads()
    n_reads = args.cpu_sample_reads
    rb, ro = synth.simulate_reads(g, n_reads, args.read_len, 0.01, 0.5, seed=44)
    kmers = n_reads * (args.read_len - args.k + 1)
    for _ in range(args.warmup):
        ref.query_reads(rb[: ro[2000]], ro[:2001], threads=threads, want_ids=False)
    t = 0.0
    for _ in range(args.steps):
        _, f, nf, sec = ref.query_reads(rb, ro, threads=threads, want_ids=False)
        t += sec
    val = kmers * args.steps / t
    line = {
        "impl": "reference", "metric": "queried k-mers/s", "value": val, "unit": "k-mers/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, 1) | {"sample": f"{n_reads} reads ({kmers} k-mers) per step"},
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": threads, "kind": "reference",
                         "sample": f"{n_reads} reads x {args.read_len} bp = {kmers} k-mers per step, in-memory OpenMP loop over query_sequence_bool"},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "found": f, "not_found": nf,
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"synthetic {args.genome / 1e6:g} Mbp random-genome unitig graph ({args.genome - args.k + 1} {args.k}-mers) replicated per GPU, "
                    f"{args.reads} simulated {args.read_len} bp reads per GPU per step (1% subst., 50% revcomp), file_query semantics (found / not-found counts; the id mode is reported under ids_mode)",
        "k": args.k, "m": args.m, "n": args.n, "s": args.s, "b": args.b,
        "kmers_per_step_per_gpu": args.reads * (args.read_len - args.k + 1),
        "parallelism": (f"partition x{world}: MPHF groups sharded, super-k-mers and ids exchanged as peer-memory stores inside the kernels"
                        if getattr(args, "partition", False) and world > 1 else f"replica x{world}, reads sharded, no data-path collective"),
        "cache": "inputs (1.5 GB of reads per step) and the index (1.8 GB in HBM) exceed the 126 MB L2; no explicit flush",
    }


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its version banner there)
    is sent to stderr, and the line goes to the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


OUT = None


def emit(line: dict):
    print(json.dumps(line), file=OUT or sys.stdout, flush=True)


def main():
    global OUT
    OUT = claim_stdout()
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from blight_b200 import api, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (blight_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    tmpdir = tempfile.mkdtemp(prefix="blight_bench_")
    g, flat, blob, build_s = build_workload_index(args, rank, world, tmpdir)
    info = flat.info()
    part = None
    if args.partition and world > 1:
        from blight_b200 import dist as bdist
        part = bdist.PartitionedSet.from_full(flat if rank == 0 else None, local, os.path.dirname(blob))
        part.enable_fused()
        idx = part.index
    else:
        idx = flat.upload(local)
    kpr = args.read_len - args.k + 1
    total_kmers = args.reads * kpr

    # synthetic reads, generated on the device (data preparation, untimed)
    d_genome = torch.from_numpy(g).to(dev)
    d_bases = synth.torch_simulate_reads(d_genome, args.reads, args.read_len, 0.01, 0.5, seed=44 + rank)
    del d_genome
    d_roff = torch.arange(0, args.reads + 1, device=dev, dtype=torch.int64) * args.read_len
    d_koff = torch.arange(0, args.reads + 1, device=dev, dtype=torch.int64) * kpr
    d_ids = torch.empty(total_kmers, dtype=torch.int64, device=dev)
    d_ctr = torch.zeros(api.N_CTR, dtype=torch.int64, device=dev)

    def step_count():
        if part is not None:
            _, c = part.query_reads_fused(d_bases, d_roff, want_ids=False, check_overflow=False)
            d_ctr.add_(c // world)  # the fused path returns the counters summed over the ranks
        else:
            idx.query_reads(d_bases, d_roff, want_ids=False, ctr=d_ctr)

    def step_ids():
        if part is not None:
            _, c = part.query_reads_fused(d_bases, d_roff, d_koff, total_kmers, ids=d_ids, check_overflow=False)
            d_ctr.add_(c // world)
        else:
            idx.query_reads(d_bases, d_roff, d_koff, total_kmers, ids=d_ids, ctr=d_ctr)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step):
        """W untimed warm-up steps, then K steps between CUDA events on the launching stream; max over ranks."""
        for _ in range(max(args.warmup, 3)):
            step()
        sync_all()
        d_ctr.zero_()
        sync_all()
        l0 = api.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        sync_all()
        launches = api.launch_count() - l0
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, d_ctr.cpu().numpy().copy()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # headline: file_query semantics (query_sequence_bool over the batch -> Good / Erroneous counters, blight.cpp:780-789)
    ms, launches, ctr = (0.0, 0, None)
    if not args.ids_only:
        ms, launches, ctr = timed(step_count)
    # the identifier mode (query_sequence_hash: one int64 id per k-mer, blight.cpp:575-591)
    ids_ms, ids_launches, ids_ctr = (0.0, 0, None)
    if not args.count_only:
        ids_ms, ids_launches, ids_ctr = timed(step_ids)
    if args.ids_only:
        ms, launches, ctr = ids_ms, ids_launches, ids_ctr
    clocks = sampler.stop() if rank == 0 else None
    found_frac = float(ctr[api.CTR_FOUND]) / max(1.0, float(ctr[api.CTR_QUERIES]))
    value = world * total_kmers * args.steps / (ms * 1e-3)

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + kernels + D2H inside the timed region ----
    if part is not None and part.overflowed():
        raise SystemExit("bench.py --partition: an inbox region overflowed, the timed steps dropped records")
    e2e = None
    if not args.no_e2e and part is None:
        h_bases = torch.empty(d_bases.numel(), dtype=torch.uint8, pin_memory=True)
        h_bases.copy_(d_bases)
        h_roff = torch.empty(args.reads + 1, dtype=torch.int64, pin_memory=True)
        h_roff.copy_(d_roff)
        torch.cuda.synchronize()
        hb, hr = h_bases.numpy(), h_roff.numpy().view(np.uint64)
        e_steps = args.steps
        for _ in range(2):
            idx.query_reads_host(hb, hr, want_ids=False)  # warm-up (sizes the library's device workspace)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            _, ectr = idx.query_reads_host(hb, hr, want_ids=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * total_kmers * e_steps / dt, "unit": "k-mers/s",
               "h2d_bytes_per_step": int(h_bases.numel() + 8 * (args.reads + 1)), "d2h_bytes_per_step": 8 * api.N_CTR,
               "steps": e_steps, "ms_per_step": 1e3 * dt / e_steps, "mode": "bool (file_query counters), pinned host reads",
               "found": int(ectr[api.CTR_FOUND]), "not_found": int(ectr[api.CTR_NOT_FOUND])}

    # ---- CPU baseline: the reference's query code on the host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            n_s = min(args.cpu_sample_reads, args.reads)
            sb = d_bases[: n_s * args.read_len].cpu().numpy()
            so = np.arange(n_s + 1, dtype=np.uint64) * np.uint64(args.read_len)
            if oracle.reference_available():
                if not os.path.exists(blob):
                    flat.save(blob)
                ref = oracle.Reference.from_blob(blob, args.k, args.m)
                threads = os.cpu_count() or ref.max_threads()
                ref.query_reads(sb[: 2000 * args.read_len], so[:2001], threads=threads, want_ids=False)
                _, f, nf, sec = ref.query_reads(sb, so, threads=threads, want_ids=False)
                kind = "reference"
            else:
                port = oracle.CPort(blob)
                threads = 1
                n_s = min(n_s, 20000)
                t0 = time.perf_counter()
                _, c3 = port.query_reads(sb[: n_s * args.read_len], so[: n_s + 1], want_ids=False)
                sec = time.perf_counter() - t0
                kind = "port"
            cpu = {"value": n_s * kpr / sec, "unit": "k-mers/s", "cores": threads, "kind": kind,
                   "sample": f"first {n_s} reads of the step ({n_s * kpr} k-mers), in-memory OpenMP loop over query_sequence_bool, {sec:.2f} s"}
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        peak, peak_src = measured_peak()
        headline_ids = args.ids_only
        balg = b_alg(args.b, args.k) - (0.0 if headline_ids else 8.0 - 0.125)  # counters instead of an int64 id per k-mer
        kernel_ms = ms / args.steps  # the step is exactly one launch of the read kernel per GPU
        achieved = balg * total_kmers / (kernel_ms * 1e-3) / 1e9
        kname = "k_reads_sk<ids>" if headline_ids else "k_reads_sk<count>"
        if part is not None:
            kname = "k_dispatch_runs + k_runs_lookup" + (" + k_scatter_runs" if headline_ids else "")
        default_cfg = (args.genome, args.reads, args.read_len, args.k, args.m, args.n, args.b) == (100_000_000, 10_000_000, 150, 31, 7, 5, 6)
        traffic = ncu_traffic(kname) if default_cfg and part is None else None
        line = {
            "metric": "queried k-mers/s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, profiles/)",
                         "algorithmic_bytes_per_launch": balg * total_kmers, "peak_source": peak_src, "kernel": kname,
                         "bytes_per_kmer_algorithmic": balg, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "ids_mode": None if (args.count_only or args.ids_only) else {
                "value": world * total_kmers * args.steps / (ids_ms * 1e-3), "unit": "k-mers/s", "ms_per_step": ids_ms / args.steps,
                "kernel": "k_reads_sk<ids>" if part is None else "k_dispatch_runs + k_runs_lookup + k_scatter_runs", "note": "query_sequence_hash semantics: one int64 id per k-mer written to HBM (9.6 GB per step)"},
            "found_fraction": found_frac, "index": {"number_kmer": info["number_kmer"], "device_bytes": idx.info["device_bytes"],
                                                     "build_seconds": build_s},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    import shutil
    shutil.rmtree(tmpdir, ignore_errors=True)  # rank 0's holds the index blob (and the slices of --partition)


if __name__ == "__main__":
    main()
