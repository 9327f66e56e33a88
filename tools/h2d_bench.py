#!/usr/bin/env python
"""Pinned host -> device copy bandwidth of this box (the bound of bench.py's e2e once the kernels are faster than the copy)."""
import json, time, torch
torch.cuda.set_device(0)
n = 1_580_000_000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for chunk_mb in (0, 64, 16):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        if chunk_mb == 0:
            d.copy_(h, non_blocking=True)
        else:
            c = chunk_mb << 20
            for o in range(0, n, c):
                d[o:o + c].copy_(h[o:o + c], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out["whole" if chunk_mb == 0 else f"chunks_{chunk_mb}MB"] = {"ms": 1e3 * dt, "GB_per_s": n / dt / 1e9}
print(json.dumps(out))
