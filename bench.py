#!/usr/bin/env python
"""bench.py — queried k-mers/s of the batched k-mer query path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]/[3], the one the north-star target is quoted on): a synthetic random-genome
unitig graph with 100 M 31-mers, index k=31 m=7 n=5 s=3 b=6, replicated on every GPU; each GPU queries its own
10 M simulated 150 bp reads (1 % substitutions, half reverse-complemented) = 1.2 G k-mers per step per GPU
(weak scaling, no data-path collective).  A step is one pass of the hot path (read tiles -> 2-bit pack -> rolling
canonical k-mers + minimizers -> MPHF -> positions -> 2^b-window compare -> counters / int64 ids) over that batch.

  value           k-mers/s, all GPUs, inputs resident in HBM, CUDA events, max over ranks
  e2e             same metric through the C ABI with HOST (pinned) buffers: H2D of the reads + kernels + D2H of the
                  counters inside the timed region (file_query semantics: Good / Erroneous counts); bytes counted by the
                  library from the copies it actually issued
  file_query      kmer_Set_Light::file_query(path) end to end from a 2-line FASTA file (N = 1)
  roofline        algorithmic bytes per k-mer x k-mers / kernel time vs measured HBM peak
  cpu_baseline    the reference's own query code (oracle/_ref) on the host cores, on a bounded sample of the reads:
                  the in-memory OpenMP loop and the reference's own file_query -t $(nproc)
  partition_mode  (N > 1) BASELINE configs[4] shape: a 1 G-k-mer index (k31 m9 n10 b6) cut by minimizer bucket over the N
                  GPUs, super-k-mers and ids exchanged as peer-memory stores inside the kernels; k-mers/s with ids and
                  counting, ratio to ONE GPU holding the whole index, ids compared with that GPU's in the run
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
    ap.add_argument("--no-file-query", action="store_true")
    ap.add_argument("--ids-only", action="store_true", help="diagnostic: time only the id (hash) mode")
    ap.add_argument("--count-only", action="store_true", help="diagnostic: time only the counting (bool) mode")
    ap.add_argument("--packed-input", action="store_true", help="diagnostic: the reads are held 2-bit packed on the device (blight_query_reads_packed)")
    ap.add_argument("--compact", action="store_true", help="diagnostic: upload without the derived tables (the reference's arrays only)")
    ap.add_argument("--no-partition", action="store_true", help="N > 1: skip the bucket-partitioned leg")
    ap.add_argument("--partition-only", action="store_true", help="diagnostic (N > 1): run the bucket-partitioned leg alone and print its record")
    ap.add_argument("--partition-returns", default="stream", help="also measure these return paths of the partitioned id mode at the default sub-batch size (comma list of stream,pull,direct; the library's default is direct)")
    ap.add_argument("--partition-genome", type=int, default=1_000_000_000)
    ap.add_argument("--partition-reads", type=int, default=4_000_000, help="reads per GPU per batch of the partitioned leg")
    ap.add_argument("--partition-shape", default="9,10,6", help="m,n,b of the partitioned index")
    ap.add_argument("--build-blob", default=None, help=argparse.SUPPRESS)  # internal: build the workload index, save it, exit
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


def bind_near_gpu(local: int):
    """Runs this process (and hence first-touches its pinned buffers) on the cores NVML lists as local to its GPU: with 8
    ranks on a two-socket box, host buffers on the far socket cost every H2D copy a trip over the socket link."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def build_flat(args):
    """The workload's index, built by the product's host builder from the seeded synthetic unitigs."""
    from blight_b200 import api, synth
    g = synth.random_genome(args.genome, seed=42)
    st, ln = synth.cut_unitigs(g, args.k, 2000, seed=43)
    # torchrun exports OMP_NUM_THREADS=1: ask for the host's cores explicitly
    return g, api.FlatIndex.build_spans(g, st, ln, args.k, args.m, args.n, args.s, args.b, threads=os.cpu_count() or 1)


def build_workload_index(args, rank, world, tmpdir):
    """Rank 0 builds the flat index and saves it; the others load the blob."""
    from blight_b200 import api, synth
    if world > 1:
        import torch.distributed as dist
        box = [tmpdir]
        dist.broadcast_object_list(box, src=0)  # every rank must look in rank 0's directory
        tmpdir = box[0]
    blob = os.path.join(tmpdir, "bench_index.blflat")
    t0 = time.time()
    if rank == 0:
        g, flat = build_flat(args)
        flat.save(blob)
    else:
        g = synth.random_genome(args.genome, seed=42)
    if world > 1:
        dist.barrier()
        if rank != 0:
            flat = api.FlatIndex.load(blob)
    return g, flat, blob, time.time() - t0


def fasta_of(reads_2d: np.ndarray) -> np.ndarray:
    """2-line FASTA text of fixed-length reads: '>\\n' + bases + '\\n' per read (the reference skips the header line
    whatever it holds, blight.cpp:760-772)."""
    n, L = reads_2d.shape
    rec = np.empty((n, L + 3), dtype=np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("\n")
    rec[:, 2:2 + L] = reads_2d
    rec[:, -1] = ord("\n")
    return rec.reshape(-1)


def shm_dir():
    return "/dev/shm" if os.path.isdir("/dev/shm") else None


def reference_measure(blob, k, m, sample_bases: np.ndarray, n_s: int, read_len: int, warm: bool = True):
    """The reference's own code on the host cores, on n_s reads: (in-memory OpenMP loop over query_sequence_bool,
    the reference's own file_query -t cores). Returns dicts."""
    import oracle
    kpr = read_len - k + 1
    so = np.arange(n_s + 1, dtype=np.uint64) * np.uint64(read_len)
    threads = os.cpu_count() or 1
    ref = oracle.Reference.from_blob(blob, k, m, cores=threads)  # file_query uses the object's core count (blight.h:18)
    if warm:
        ref.query_reads(sample_bases[: 2000 * read_len], so[:2001], threads=threads, want_ids=False)
    _, f, nf, sec = ref.query_reads(sample_bases, so, threads=threads, want_ids=False)
    mem = {"value": n_s * kpr / sec, "unit": "k-mers/s", "cores": threads, "kind": "reference", "seconds": sec, "found": f, "not_found": nf,
           "sample": f"first {n_s} reads of the step ({n_s * kpr} k-mers), in-memory OpenMP loop over query_sequence_bool, {sec:.2f} s"}
    fq = None
    try:
        d = tempfile.mkdtemp(prefix="blight_ref_fq_", dir=shm_dir())
        path = os.path.join(d, "sample.fa")
        fasta_of(sample_bases.reshape(n_s, read_len)).tofile(path)
        q0 = int(ref.L.blref_number_query(ref.h))
        t0 = time.perf_counter()
        rc = ref.file_query(path)
        dt = time.perf_counter() - t0
        done = int(ref.L.blref_number_query(ref.h)) - q0
        os.remove(path)
        os.rmdir(d)
        if rc == 0 and done > 0:
            fq = {"value": done / dt, "unit": "k-mers/s", "cores": threads, "seconds": dt,
                  "sample": f"kmer_Set_Light::file_query on a FASTA file of the same {n_s} reads (tmpfs), -t {threads}, wall clock around the call (file reading included, blight.cpp:746-799)"}
    except Exception as ex:  # reported, never required
        fq = {"value": None, "error": repr(ex)}
    return mem, fq


def run_reference(args, rank):
    """--impl reference: the reference's own CPU query code (oracle/_ref: /root/reference + fixes P1/P2, compiled by
    oracle/build_ref.sh) on the host cores, all threads, on a bounded sample of the same workload per step. This process
    never maps the product library: the index blob is built by a child process and imported into the reference object."""
    if rank != 0:
        return
    import oracle
    from blight_b200 import synth
    if not oracle.reference_available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libblight_ref.so was not built (needs /root/reference at build time)"})
        return
    td = tempfile.mkdtemp(prefix="blight_refarm_", dir=shm_dir())
    blob = os.path.join(td, "bench_index.blflat")
    cmd = [sys.executable, os.path.abspath(__file__), "--build-blob", blob, "--genome", str(args.genome), "--k", str(args.k), "--m", str(args.m),
           "--n", str(args.n), "--s", str(args.s), "--b", str(args.b)]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    g = synth.random_genome(args.genome, seed=42)
    threads = os.cpu_count() or 1
    ref = oracle.Reference.from_blob(blob, args.k, args.m, cores=threads)
    n_reads = args.cpu_sample_reads
    rb, ro = synth.simulate_reads(g, n_reads, args.read_len, 0.01, 0.5, seed=44)
    kmers = n_reads * (args.read_len - args.k + 1)
    for _ in range(args.warmup):
        ref.query_reads(rb[: ro[2000]], ro[:2001], threads=threads, want_ids=False)
    t, step_s = 0.0, []
    for _ in range(args.steps):
        _, f, nf, sec = ref.query_reads(rb, ro, threads=threads, want_ids=False)
        t += sec
        step_s.append(round(sec, 4))
    val = kmers * args.steps / t
    _, fq = reference_measure(blob, args.k, args.m, rb, n_reads, args.read_len, warm=False)
    os.remove(blob)
    os.rmdir(td)
    mapped = [ln.split()[-1] for ln in open("/proc/self/maps") if "libblight_b200" in ln]
    line = {
        "impl": "reference", "metric": "queried k-mers/s", "value": val, "unit": "k-mers/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, 1) | {"sample": f"{n_reads} reads ({kmers} k-mers) per step"},
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": threads, "kind": "reference",
                         "sample": f"{n_reads} reads x {args.read_len} bp = {kmers} k-mers per step, in-memory OpenMP loop over query_sequence_bool",
                         "file_query": fq},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "found": f, "not_found": nf, "product_library_mapped": bool(mapped),
        "step_seconds": step_s, "host": {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "loadavg": os.getloadavg()},
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"synthetic {args.genome / 1e6:g} Mbp random-genome unitig graph ({args.genome - args.k + 1} {args.k}-mers) replicated per GPU, "
                    f"{args.reads} simulated {args.read_len} bp reads per GPU per step (1% subst., 50% revcomp), file_query semantics (found / not-found counts; the id mode is reported under ids_mode)",
        "k": args.k, "m": args.m, "n": args.n, "s": args.s, "b": args.b,
        "kmers_per_step_per_gpu": args.reads * (args.read_len - args.k + 1),
        "parallelism": f"replica x{world}, reads sharded, no data-path collective" + ("; partition_mode: MPHF groups sharded, super-k-mers and ids exchanged as peer-memory stores inside the kernels" if world > 1 else ""),
        "cache": "inputs (1.5 GB of reads per step) and the index (2 GB in HBM) exceed the 126 MB L2; no explicit flush",
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


def read_kernel_name(info, want_ids, k, m):
    """The kernel the launcher picks for a mode (kernels.cu: launch_reads_t / use_superkmer_kernel), from the layout in force."""
    from blight_b200 import api
    forced = os.environ.get("BLIGHT_READS_KERNEL", "")[:1]
    sk = k - m + 1 >= 8 and not (want_ids and not (info["layout"] & api.LAYOUT_POS_ID))
    if sk and forced == "p":
        sk = False
    return ("k_reads_sk" if sk else "k_reads") + ("<ids>" if want_ids else "<count>")


def partition_leg(args, rank, world, local, dev):
    """BASELINE configs[4] shape under the driver's own launch: a 1 G-k-mer index cut by minimizer bucket over the ranks."""
    import torch
    import torch.distributed as dist
    from blight_b200 import api, synth
    from blight_b200 import dist as bdist
    m, n, b = (int(x) for x in args.partition_shape.split(","))
    gl, n_reads = args.partition_genome, args.partition_reads
    try:
        import psutil
        avail_gb = psutil.virtual_memory().available / 1e9
    except Exception:
        avail_gb = float("inf")
    need_gb = world * (gl * 5e-9 + 2) + gl * 12e-9
    if avail_gb < need_gb:
        return {"skipped": f"needs about {need_gb:.0f} GB of host memory, {avail_gb:.0f} GB available"}
    t_leg = time.time()
    wd = [None]
    if rank == 0:
        wd = [tempfile.mkdtemp(prefix="blight_part_", dir=shm_dir())]
    dist.broadcast_object_list(wd, src=0)
    blob = os.path.join(wd[0], "full.blflat")
    g = synth.random_genome(gl, seed=42)
    flat = None
    t0 = time.time()
    if rank == 0:
        st, ln = synth.cut_unitigs(g, args.k, 2000, seed=43)
        flat = api.FlatIndex.build_spans(g, st, ln, args.k, m, n, 3, b, threads=os.cpu_count() or 1)
        flat.save(blob)
    build_s = time.time() - t0
    part = bdist.PartitionedSet.from_full(flat, local, wd[0])
    if rank != 0:
        flat = api.FlatIndex.load(blob)
    N = flat.info()["number_kmer"]
    whole = flat.upload(local)  # every rank also holds the whole index: ground truth for its own reads, and the one-GPU rate
    del flat
    d_genome = torch.from_numpy(g).to(dev)
    del g
    kpr = args.read_len - args.k + 1
    bases = synth.torch_simulate_reads(d_genome, n_reads, args.read_len, 0.01, 0.5, seed=144 + rank)
    del d_genome
    roff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * args.read_len
    koff = torch.arange(0, n_reads + 1, device=dev, dtype=torch.int64) * kpr
    total = n_reads * kpr
    reps = max(3, min(args.steps, 10))

    def timed(fn):
        for _ in range(3):
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

    ids_one, ctr_one = whole.query_reads(bases, roff, koff, total)
    torch.cuda.synchronize()
    ids_one = ids_one[:total]
    ctr_all = ctr_one.clone()
    dist.all_reduce(ctr_all)
    scratch = torch.empty(total, dtype=torch.int64, device=dev)
    one_ids_ms = timed(lambda: whole.query_reads(bases, roff, koff, total, ids=scratch))
    one_cnt_ms = timed(lambda: whole.query_reads(bases, roff, want_ids=False))
    whole_bytes = whole.info["device_bytes"]

    del scratch
    variants, same_all, ctr_ok_all, ovf_all = {}, True, True, False
    def measure(sub, want_ids_session, return_path=None):
        part.enable_fused(want_ids=want_ids_session, sub_positions=sub, ids_capacity=total if want_ids_session else 0, return_path=return_path)
        out = {"sub_positions": part._sub}
        ok_ids = ok_ctr = True
        if want_ids_session:
            ids_f, ctr_f = part.query_reads_fused(bases, roff, koff, total)
            torch.cuda.synchronize()
            same = torch.tensor([1 if torch.equal(ids_f, ids_one) else 0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            ok_ids = bool(same.item())
            ok_ctr = bool(torch.equal(ctr_f.cpu()[:3], ctr_all.cpu()[:3]))
            out["ids_ms"] = timed(lambda: part.query_reads_fused(bases, roff, koff, total, check_overflow=False))
            out["ids_vs_one_gpu"] = world * one_ids_ms / out["ids_ms"]
        _, ctr_c = part.query_reads_fused(bases, roff, want_ids=False)
        torch.cuda.synchronize()
        ok_ctr = ok_ctr and bool(torch.equal(ctr_c.cpu()[:3], ctr_all.cpu()[:3]))
        out["counting_ms"] = timed(lambda: part.query_reads_fused(bases, roff, want_ids=False, check_overflow=False))
        out["counting_vs_one_gpu"] = world * one_cnt_ms / out["counting_ms"]
        ovf = torch.tensor([1 if part.overflowed() else 0], device=dev)
        dist.all_reduce(ovf, op=dist.ReduceOp.MAX)
        out.update({"ids_equal_replica": ok_ids, "counters_equal_replica": ok_ctr, "overflow": bool(ovf.item())})
        return out

    variants, same_all, ctr_ok_all, ovf_all = {}, True, True, False
    for sub in sorted({64 << 20, bdist.DEFAULT_SUB_IDS, bdist.DEFAULT_SUB_COUNTING}):
        v = measure(sub, True)
        variants[f"{sub >> 20}M"] = v
        same_all &= v["ids_equal_replica"]; ctr_ok_all &= v["counters_equal_replica"]; ovf_all |= v["overflow"]
    other_returns = {}
    for rp in [x for x in args.partition_returns.split(",") if x]:
        for sub in (bdist.DEFAULT_SUB_IDS,):
            v = measure(sub, True, rp)
            other_returns[f"{rp}/{sub >> 20}M"] = v
            same_all &= v["ids_equal_replica"]; ctr_ok_all &= v["counters_equal_replica"]; ovf_all |= v["overflow"]
    f_ids_ms = variants[f"{bdist.DEFAULT_SUB_IDS >> 20}M"]["ids_ms"]
    f_cnt_ms = variants[f"{bdist.DEFAULT_SUB_COUNTING >> 20}M"]["counting_ms"]
    default_order = f"ids: sub-batches of {bdist.DEFAULT_SUB_IDS >> 20} M positions; counting: {bdist.DEFAULT_SUB_COUNTING >> 20} M (the library's defaults per mode, blight_b200/dist.py)"
    del ids_one, whole
    torch.cuda.empty_cache()
    local_bytes = part.index.info["device_bytes"]
    part.disable_fused()
    dist.barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(wd[0], ignore_errors=True)
    return {
        "workload": f"synthetic {gl / 1e6:g} Mbp random-genome unitig graph ({N} {args.k}-mers), index k={args.k} m={m} n={n} b={b} cut into {world} contiguous "
                    f"ranges of MPHF groups (one per GPU); every GPU holds {n_reads} reads ({total} k-mers) per batch; CUDA events, max over ranks, {reps} batches",
        "ids": {"value": world * total / (f_ids_ms * 1e-3), "unit": "k-mers/s", "ms_per_batch": f_ids_ms},
        "counting": {"value": world * total / (f_cnt_ms * 1e-3), "unit": "k-mers/s", "ms_per_batch": f_cnt_ms},
        "one_gpu_whole_index": {"ids": total / (one_ids_ms * 1e-3), "counting": total / (one_cnt_ms * 1e-3), "unit": "k-mers/s",
                                "ids_ms": one_ids_ms, "counting_ms": one_cnt_ms, "device_bytes": whole_bytes},
        "ids_vs_one_gpu": world * one_ids_ms / f_ids_ms, "counting_vs_one_gpu": world * one_cnt_ms / f_cnt_ms,
        "ids_equal_replica": same_all, "counters_equal_replica": ctr_ok_all, "overflow": ovf_all,
        "variant": default_order, "by_sub_batch_size": variants, **({"by_return_path": other_returns} if other_returns else {}),
        "device_bytes_per_gpu": local_bytes, "cuts": part.plan.cuts,
        "return_path": "direct (the library's default): the owner stores int64 ids straight into the source's id array over NVLink, run by run; owners take their sources round-robin so that no GPU is the target of all the others at once; ordering between GPUs by device-side flags (csrc/part_session.cu). by_return_path: the other return paths at the default sub-batch size",
        "build_seconds": build_s, "leg_seconds": time.time() - t_leg,
    }


def main():
    global OUT
    args = parse()
    if args.build_blob:
        _, flat = build_flat(args)
        flat.save(args.build_blob)
        return
    OUT = claim_stdout()
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
    cpus = bind_near_gpu(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    if args.partition_only:
        if world < 2:
            raise SystemExit("--partition-only needs torchrun with at least 2 ranks")
        pm = partition_leg(args, rank, world, local, dev)
        if rank == 0:
            OUT.write(json.dumps(pm) + "\n")
            OUT.flush()
        dist.barrier()
        dist.destroy_process_group()
        return
    tmpdir = tempfile.mkdtemp(prefix="blight_bench_", dir=shm_dir())
    g, flat, blob, build_s = build_workload_index(args, rank, world, tmpdir)
    info = flat.info()
    t0 = time.time()
    idx = flat.upload(local, api.UploadOptions.compact() if args.compact else None)
    torch.cuda.synchronize()
    upload_s = time.time() - t0
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

    d_packed = None
    if args.packed_input:
        code = ((d_bases >> 1) & 3).to(torch.int64)
        pad = (-code.numel()) % 16
        if pad:
            code = torch.cat([code, torch.zeros(pad, dtype=torch.int64, device=dev)])
        sh = torch.arange(30, -2, -2, device=dev, dtype=torch.int64)
        d_packed = ((code.view(-1, 16) << sh).sum(1) & 0xFFFFFFFF).to(torch.int32)  # 16 bases per word, first base in the high bits
        d_packed = torch.cat([d_packed, torch.zeros(64, dtype=torch.int32, device=dev)])
        del code

    def step_count():
        if d_packed is not None:
            idx.query_reads_packed(d_packed, d_roff, d_bases.numel(), want_ids=False, ctr=d_ctr)
        else:
            idx.query_reads(d_bases, d_roff, want_ids=False, ctr=d_ctr)

    def step_ids():
        if d_packed is not None:
            idx.query_reads_packed(d_packed, d_roff, d_bases.numel(), d_koff, total_kmers, ids=d_ids, ctr=d_ctr)
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
    del d_ids

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + kernels + D2H inside the timed region ----
    e2e = None
    h_bases = None
    if not args.no_e2e:
        h_bases = torch.empty(d_bases.numel(), dtype=torch.uint8, pin_memory=True)
        h_bases.copy_(d_bases)
        h_roff = torch.empty(args.reads + 1, dtype=torch.int64, pin_memory=True)
        h_roff.copy_(d_roff)
        torch.cuda.synchronize()
        hb, hr = h_bases.numpy(), h_roff.numpy().view(np.uint64)
        e_steps = args.steps
        for _ in range(2):
            idx.query_reads_host(hb, hr, want_ids=False)  # warm-up (sizes the library's device workspace and staging)
        sync_all()
        x0 = api.transfer_bytes()
        k0 = api.host_pack_stats()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            _, ectr = idx.query_reads_host(hb, hr, want_ids=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        x1 = api.transfer_bytes()
        k1 = api.host_pack_stats()
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        # what the same bytes cost as plain pinned copies on this box, all ranks at once: the floor of e2e once the kernels are
        # faster than the link (on the 8-GPU box half of the GPUs get 23 GB/s, tools/h2d_bench.py)
        d_sink = torch.empty(h_bases.numel(), dtype=torch.uint8, device=dev)
        d_sink.copy_(h_bases, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(3):
            d_sink.copy_(h_bases, non_blocking=True)
        torch.cuda.synchronize()
        ht = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ht, op=dist.ReduceOp.MAX)
        del d_sink
        e2e = {"value": world * total_kmers * e_steps / dt, "unit": "k-mers/s", "ascii_copy_only_ms_per_step": 1e3 * float(ht.item()),
               "h2d_bytes_per_step": (x1[0] - x0[0]) // e_steps, "d2h_bytes_per_step": (x1[1] - x0[1]) // e_steps,
               "ascii_bytes_per_step": int(h_bases.numel()), "steps": e_steps, "ms_per_step": 1e3 * dt / e_steps,
               "mode": "bool (file_query counters), pinned host reads; part of the batch crosses PCIe 2-bit packed by the host cores (bytes as counted by the library)",
               "host_threads_near_gpu": len(cpus) if cpus else None,
               "host_packer": {"threads": k1[2], "bases_packed_per_step": (k1[0] - k0[0]) // e_steps,
                               "GB_per_s_while_packing": (k1[0] - k0[0]) / max(1e-9, k1[1] - k0[1]) / 1e9},
               "found": int(ectr[api.CTR_FOUND]), "not_found": int(ectr[api.CTR_NOT_FOUND])}

    # ---- file_query(path): 2-line FASTA file on tmpfs through kmer_Set_Light::file_query's replacement (N = 1) ----
    fq = None
    if rank == 0 and world == 1 and not args.no_file_query:
        try:
            src = h_bases.numpy() if h_bases is not None else d_bases.cpu().numpy()
            path = os.path.join(tmpdir, "reads.fa")
            fasta_of(src.reshape(args.reads, args.read_len)).tofile(path)
            fbytes = os.path.getsize(path)
            idx.query_file_host(path)  # warm-up: page cache, pinned buffers
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                c = idx.query_file_host(path)
            dt = (time.perf_counter() - t0) / reps
            os.remove(path)
            fq = {"value": total_kmers / dt, "unit": "k-mers/s", "seconds": dt, "file_bytes": fbytes, "file_GB_per_s": fbytes / dt / 1e9,
                  "found": int(c[0]), "not_found": int(c[1]), "note": "blight_query_file_host: streaming reader -> record cut -> H2D / kernel overlap, wall clock"}
        except Exception as ex:
            fq = {"value": None, "error": repr(ex)}

    # ---- CPU baseline: the reference's query code on the host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import oracle
            n_s = min(args.cpu_sample_reads, args.reads)
            sb = d_bases[: n_s * args.read_len].cpu().numpy()
            if not os.path.exists(blob):
                flat.save(blob)
            if oracle.reference_available():
                cpu, cpu_fq = reference_measure(blob, args.k, args.m, sb, n_s, args.read_len)
                cpu["file_query"] = cpu_fq
            else:
                port = oracle.CPort(blob)
                n_s = min(n_s, 20000)
                so = np.arange(n_s + 1, dtype=np.uint64) * np.uint64(args.read_len)
                t0 = time.perf_counter()
                port.query_reads(sb[: n_s * args.read_len], so, want_ids=False)
                sec = time.perf_counter() - t0
                cpu = {"value": n_s * kpr / sec, "unit": "k-mers/s", "cores": 1, "kind": "port",
                       "sample": f"first {n_s} reads of the step ({n_s * kpr} k-mers), single-threaded C restatement, {sec:.2f} s"}
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    dev_bytes, layout = idx.info["device_bytes"], idx.info["layout"]
    # ---- partition mode (N > 1) ----
    pm = None
    if world > 1 and not args.no_partition:
        del idx, d_bases, d_roff, d_koff, h_bases
        torch.cuda.empty_cache()
        try:
            pm = partition_leg(args, rank, world, local, dev)
        except Exception as ex:
            pm = {"error": repr(ex)}

    if rank == 0:
        peak, peak_src = measured_peak()
        headline_ids = args.ids_only
        balg = b_alg(args.b, args.k) - (0.0 if headline_ids else 8.0 - 0.125)  # counters instead of an int64 id per k-mer
        kernel_ms = ms / args.steps  # the step is exactly one launch of the read kernel per GPU
        achieved = balg * total_kmers / (kernel_ms * 1e-3) / 1e9
        info_d = {"layout": layout}
        kname = read_kernel_name(info_d, headline_ids, args.k, args.m)
        default_cfg = (args.genome, args.reads, args.read_len, args.k, args.m, args.n, args.b) == (100_000_000, 10_000_000, 150, 31, 7, 5, 6)
        default_layout = layout == (api.LAYOUT_POS_ID | api.LAYOUT_FILTER | api.LAYOUT_EXACT_POS) and not os.environ.get("BLIGHT_FILTER_BITS")
        traffic = ncu_traffic(kname) if default_cfg and default_layout else None
        line = {
            "metric": "queried k-mers/s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, profiles/)",
                         "algorithmic_bytes_per_launch": balg * total_kmers, "peak_source": peak_src, "kernel": kname,
                         "bytes_per_kmer_algorithmic": balg, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu, "e2e": e2e, "file_query": fq, "gpu_launches": int(launches), "clocks": clocks,
            "ids_mode": None if (args.count_only or args.ids_only) else {
                "value": world * total_kmers * args.steps / (ids_ms * 1e-3), "unit": "k-mers/s", "ms_per_step": ids_ms / args.steps,
                "kernel": read_kernel_name(info_d, True, args.k, args.m), "note": "query_sequence_hash semantics: one int64 id per k-mer written to HBM (9.6 GB per step)"},
            "found_fraction": found_frac,
            "index": {"number_kmer": info["number_kmer"], "device_bytes": dev_bytes, "bits_per_kmer": 8.0 * dev_bytes / max(1, info["number_kmer"]),
                      "layout": {"pos_id": bool(layout & api.LAYOUT_POS_ID), "filter": bool(layout & api.LAYOUT_FILTER), "exact_pos": bool(layout & api.LAYOUT_EXACT_POS)},
                      "reference_arrays_bits_per_kmer": (info["positions_bits"] + info["mphf_bits"] + 2 * info["total_nuc"]) / max(1, info["number_kmer"]),
                      "build_seconds": build_s, "upload_seconds": upload_s},
            "partition_mode": pm,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    import shutil
    shutil.rmtree(tmpdir, ignore_errors=True)  # rank 0's holds the index blob


if __name__ == "__main__":
    main()
