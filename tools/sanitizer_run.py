#!/usr/bin/env python3
"""A short frame loop through the C ABI for compute-sanitizer (memcheck / racecheck): S streams (a partial 128-tile,
a partial pitch group, two lanes when S >= 1024), F frames, host-buffer and device-pointer calls, multi-frame call.
usage: compute-sanitizer --tool memcheck python tools/sanitizer_run.py [S] [F]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rnnoise_b200  # noqa: E402
from rnnoise_b200.synth_pcm import batch_pcm  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 300
F = int(sys.argv[2]) if len(sys.argv) > 2 else 6
base = batch_pcm(min(S, 32), F)
pcm = np.ascontiguousarray(np.tile(base, (1, (S + base.shape[1] - 1) // base.shape[1], 1))[:, :S])
model = rnnoise_b200.Model(os.path.join(ROOT, "tests", "golden", "models", "default.bin"))
b = rnnoise_b200.Batch(model, S)
for f in range(F):
    out, vad = b.process(pcm[f])
x = np.ascontiguousarray(pcm.transpose(1, 0, 2).reshape(S, F * 480))
out, vad = b.process_frames(x)
b.process_s16(pcm[0].astype(np.int16))
b.destroy()
model.free()
print("sanitizer_run done:", S, "streams,", F, "frames; checksum", float(np.abs(out).sum()))
