#!/usr/bin/env python
"""One-line digest of a bench.py record: tools/show_bench.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e, i = d.get("e2e") or {}, d.get("ids_mode") or {}
print(sys.argv[1], "count %.4e (%.2f ms)" % (d["value"], d["ms_per_step"]), "ids %s (%s ms)" % (i.get("value"), i.get("ms_per_step")),
      "e2e %s (%s ms)" % (e.get("value"), e.get("ms_per_step")), "frac %.3f" % d["roofline"]["frac"])
