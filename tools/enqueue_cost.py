#!/usr/bin/env python3
"""Host-side cost of one step (CPU time to enqueue a frame, GPU idle-independent): the async host call vs
the device-pointer call.  If enqueue time per step approaches the GPU's step time, e2e is CPU-bound."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import rnnoise_b200 as rb

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = rb.Model(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "models", "default.bin"))
b = rb.Batch(model, S)
pcm = torch.randn(8, S, 480).mul_(1000).pin_memory()
out = [torch.empty(S, 480).pin_memory() for _ in range(4)]
vad = [torch.empty(S).pin_memory() for _ in range(4)]
for name, N in (("warm", 20), ("host_async", 200)):
    b.sync(); t0 = time.perf_counter()
    for i in range(N):
        b.process_ptr_async(out[i % 4].data_ptr(), pcm[i % 8].data_ptr(), vad[i % 4].data_ptr())
    t1 = time.perf_counter(); b.sync(); t2 = time.perf_counter()
    print(f"{name}: enqueue {1e3 * (t1 - t0) / N:.4f} ms/step, total {1e3 * (t2 - t0) / N:.4f} ms/step")
d = pcm.cuda(); do = torch.empty(S, 480, device="cuda"); dv = torch.empty(S, device="cuda")
st = torch.cuda.Stream(priority=-1)
b.set_stream(st.cuda_stream)
for name, N in (("warm", 20), ("device", 200), ("device+hint", 200)):
    b.sync(); t0 = time.perf_counter()
    for i in range(N):
        if name == "device+hint":
            if i == 0:
                b.prefilter_device(d[0].data_ptr())
            b.prefilter_device(d[(i + 1) % 8].data_ptr())
        b.process_device(do.data_ptr(), d[i % 8].data_ptr(), dv.data_ptr())
    t1 = time.perf_counter(); b.sync(); t2 = time.perf_counter()
    print(f"{name}: enqueue {1e3 * (t1 - t0) / N:.4f} ms/step, total {1e3 * (t2 - t0) / N:.4f} ms/step")
# raw PCIe copy time for one step's PCM (pinned <-> device), alone and both directions at once
h = torch.empty(S, 480).pin_memory(); g = torch.empty(S, 480, device="cuda"); g2 = torch.empty(S, 480, device="cuda"); h2 = torch.empty(S, 480).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name in ("h2d", "d2h", "both"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(50):
        if name in ("h2d", "both"):
            with torch.cuda.stream(s1): g.copy_(h, non_blocking=True)
        if name in ("d2h", "both"):
            with torch.cuda.stream(s2): h2.copy_(g2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print(f"copy {name}: {dt * 1e3:.4f} ms per {S * 480 * 4 / 1e6:.1f} MB -> {S * 480 * 4 / dt / 1e9:.1f} GB/s per direction")
