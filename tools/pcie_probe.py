#!/usr/bin/env python3
"""Bare cudaMemcpyAsync bandwidth of this box for the buffer sizes the host-call path moves:
pinned host <-> device, one direction alone and both directions at once, one copy per step vs split in lanes.
Prints one JSON line."""
import json
import sys
import time

import torch

dev = torch.device("cuda", 0)
res = {"gpu": torch.cuda.get_device_name(0)}
for mb in (1, 2, 4, 8, 16, 64, 256):
    n = mb * (1 << 20)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = max(5, min(200, 2048 // mb))

    def run(h2d, d2h, split=1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for k in range(split):
                a, b = k * n // split, (k + 1) * n // split
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in[a:b].copy_(h_in[a:b], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out[a:b].copy_(d_out[a:b], non_blocking=True)
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t0) / 1e9

    run(True, True)
    res[f"{mb}MB"] = {"h2d_GBps": round(run(True, False), 2), "d2h_GBps": round(run(False, True), 2),
                      "both_each_GBps": round(run(True, True), 2), "h2d_split4_GBps": round(run(True, False, 4), 2)}
# pageable for comparison
n = 8 << 20
hp = torch.empty(n, dtype=torch.uint8); d = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    d.copy_(hp)
torch.cuda.synchronize()
res["8MB_pageable_h2d_GBps"] = round(n * 20 / (time.perf_counter() - t0) / 1e9, 2)
try:
    import subprocess
    res["pcie"] = subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv,noheader"],
                                 capture_output=True, text=True).stdout.strip().splitlines()[0]
    res["numa"] = subprocess.run("lscpu | grep -i 'numa node' | head -4", shell=True, capture_output=True, text=True).stdout.strip()
except Exception as ex:  # noqa: BLE001
    res["pcie"] = str(ex)
print(json.dumps(res))
