#!/usr/bin/env python
"""SASS size of one kernel attributed to CUDA source lines (no GPU needed).
usage: tools/sass_lines.py build/obj/kernels.cu.o k_reads_skILi2ELb1 [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
cur, sec, cnt = None, False, collections.Counter()
for l in txt.split("\n"):
    if l.startswith("//---") and ".text." in l:
        sec = pat in l
        continue
    if not sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print("total SASS instructions", tot, "=", tot * 16, "bytes")
byfile = collections.Counter()
for k, v in cnt.items():
    byfile[k[0] if k else None] += v
print(byfile.most_common())
for k, v in cnt.most_common(top):
    print(v, k)
