#!/usr/bin/env python
"""Instruction / stall share of an .ncu-rep per phase of k_reads_sk (phases found by the '// A.' ... comments in
kernels.cu) plus the stall-reason mix. usage: tools/ncu_phases.py report.ncu-rep"""
import csv, io, subprocess, sys, os
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 10 or r[0] in ("Line No", "Function Name") or r[0] == "": continue
    try: agg.append((cur, int(r[0]), r[1].strip(), int(r[6]), int(r[7]), int(r[8])))
    except ValueError: pass
ti = sum(a[4] for a in agg) or 1; ts = sum(a[3] for a in agg) or 1
lines = open(os.path.join(ROOT, "blight_b200/csrc/kernels.cu")).read().split("\n")
def find(s, start=0):
    for i in range(start, len(lines)):
        if s in lines[i]: return i + 1
    return 10**9
sk = find("k_reads_sk(DevIndexView I")
marks = [("prolog", sk)] + [(n, find(t, sk)) for n, t in [("A pack", "// A. pack"), ("B keys", "// B. m-mer keys"), ("C1", "// C1."), ("C3", "// C3."), ("C4a", "// C4a."), ("C2/C4b", "// C2 (phase 0")]]
marks.append(("epilog", find("for (int o = 16; o > 0; o >>= 1)", marks[-1][1])))
ph = {}
for a in agg:
    f, l = a[0], a[1]
    name = f
    if f == "kernels.cu":
        name = "other-k"
        for (n, s), (n2, e) in zip(marks, marks[1:] + [("x", 10**9)]):
            if s <= l < e: name = n
    d = ph.setdefault(name, [0, 0, 0]); d[0] += a[4]; d[1] += a[3]; d[2] += a[5]
for k, v in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} instr {100*v[0]/ti:5.1f}%  stalls {100*v[1]/ts:5.1f}%  lanes {v[2]/max(v[0],1):5.1f}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
M = dict(zip(rows[0], rows[2]))
tot = 0; st = {}
for k, v in M.items():
    if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
        st[k[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(v); tot += float(v)
print("stall mix:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
for k in ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]:
    print(k, M.get(k))
