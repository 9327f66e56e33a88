#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed) into the text committed under profiles/:
key raw metrics of the first kernel in the report + instruction / stall share per CUDA source line.
usage: tools/ncu_summary.py report.ncu-rep out.txt [kmers_per_launch]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
kmers = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_requests_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = [f"# ncu summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)"]
for k in keys:
    if k in M:
        lines.append(f"{k:75s} {M[k][0]} {M[k][1]}")
if kmers:
    try:
        t = float(M["gpu__time_duration.sum"][0]); tu = M["gpu__time_duration.sum"][1]
        t_s = t * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(tu, 1e-3)
        dr = float(M["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12}.get(M["dram__bytes_read.sum"][1], 1)
        dw = float(M["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12}.get(M["dram__bytes_write.sum"][1], 1)
        lines += ["", f"k-mers per launch                         {kmers:.0f}",
                  f"k-mers/s under ncu (not a bench value)    {kmers / t_s:.4g}",
                  f"DRAM traffic per launch (read+write)      {(dr + dw) / 1e9:.1f} GB = {(dr + dw) / kmers:.1f} B per k-mer (algorithmic: 158 B at b<=6)",
                  f"DRAM sectors read per k-mer               {float(M['dram__sectors_read.sum'][0]) / kmers:.2f}",
                  f"L1->L2 sectors per k-mer                  {float(M['l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum'][0]) / kmers:.2f}",
                  f"L2 lookup sectors per L1 request sector   {float(M['lts__t_sectors_srcunit_tex_op_read.sum'][0]) / float(M['l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum'][0]):.2f}",
                  f"warp instructions per 32 k-mers           {float(M['smsp__inst_executed.sum'][0]) / (kmers / 32):.0f}"]
    except Exception as e:  # pragma: no cover
        lines.append(f"(derived figures unavailable: {e})")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
agg = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] in ("Line No", "Function Name") or r[0] == "":
        continue
    try:
        agg.append((cur, int(r[0]), r[1].strip(), int(r[6]), int(r[7]), int(r[8])))
    except ValueError:
        pass
ts = sum(a[3] for a in agg) or 1
ti = sum(a[4] for a in agg) or 1
lines += ["", "## top source lines by executed warp instructions (share of instructions | share of stall samples | active lanes per instruction)"]
for a in sorted(agg, key=lambda a: -a[4])[:30]:
    lines.append(f"{a[0]}:{a[1]:<4d} {100 * a[4] / ti:5.1f}% {100 * a[3] / ts:5.1f}% {a[5] / max(a[4], 1):5.1f}  {a[2][:95]}")
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 4:
    # usage: ... kmers_per_launch kernel_key  -> records the DRAM traffic bench.py reports as roofline.traffic
    import json, os
    tf = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
    d = json.load(open(tf)) if os.path.exists(tf) else {}
    d[sys.argv[4]] = {"dram_bytes_per_launch": dr + dw, "source": os.path.basename(out), "kmers_per_launch": kmers}
    json.dump(d, open(tf, "w"), indent=1, sort_keys=True)
print("\n".join(lines[:40]))
