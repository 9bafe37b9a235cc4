#!/usr/bin/env python3
"""Same-box A/B of engine variants selected by environment variables at batch creation: per-kernel CUDA-event times
(serialised profile pass) and the pipelined device-resident step, default model.  Every configuration is run
twice, interleaved.
usage: python tools/ab_env.py [--streams S] "VAR=a;VAR2=b" "VAR=c" ""      (one argument per configuration)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rnnoise_b200  # noqa: E402
from bench import make_pool  # noqa: E402

args = sys.argv[1:]
S = 4096
if args and args[0] == "--streams":
    S = int(args[1]); args = args[2:]
configs = args or [""]
VARS = sorted({kv.split("=")[0] for c in configs for kv in c.split(";") if kv})
dev = torch.device("cuda", 0)
pool = torch.from_numpy(make_pool(S)).to(dev)
F = pool.shape[0]
out = torch.empty(S, 480, device=dev); vad = torch.empty(S, device=dev)
model = rnnoise_b200.Model(os.path.join(ROOT, "tests", "golden", "models", "default.bin"))
res = {}
for mode in configs + configs:
    for v in VARS:
        os.environ.pop(v, None)
    for kv in mode.split(";"):
        if kv:
            os.environ[kv.split("=")[0]] = kv.split("=", 1)[1]
    try:
        b = rnnoise_b200.Batch(model, S, 0)
    except Exception as ex:  # noqa: BLE001
        res.setdefault(mode, []).append({"error": str(ex)})
        continue
    st = torch.cuda.Stream(dev); torch.cuda.set_stream(st); b.set_stream(st.cuda_stream)
    b.prefilter_device(pool[0].data_ptr())
    def step(i):
        b.prefilter_device(pool[(i + 1) % F].data_ptr())
        b.process_device(out.data_ptr(), pool[i % F].data_ptr(), vad.data_ptr())
    for i in range(50):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 1000
    e0.record(st)
    for i in range(K):
        step(50 + i)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    b.process_device(out.data_ptr(), pool[(50 + K) % F].data_ptr(), vad.data_ptr())
    b.sync()
    b.profile(True)
    for i in range(30):
        b.process_device(out.data_ptr(), pool[i % F].data_ptr(), vad.data_ptr())
    times, n = b.profile_read()
    b.profile(False)
    res.setdefault(mode, []).append({"step_ms": ms, "frames_per_s": S / ms * 1e3, "kernels_us": {k: round(1e3 * v / n, 1) for k, v in times.items()}})
    b.destroy()
print(json.dumps({"streams": S, "results": res}))
