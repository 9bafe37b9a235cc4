#!/usr/bin/env python3
"""Stage timeline of the pipelined host call (float or int16 PCM): where each frame's H2D, prefilter, pitch,
spectrum, network+synthesis and D2H start and end.  usage: python tools/timeline.py [s16] [streams]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["RNNOISE_B200_TIMELINE"] = "48"
import torch
import rnnoise_b200 as rb

s16 = "s16" in sys.argv[1:]
S = next((int(a) for a in sys.argv[1:] if a.isdigit()), 4096)
model = rb.Model(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "models", "default.bin"))
b = rb.Batch(model, S)
dt = torch.int16 if s16 else torch.float32
pcm = (torch.randn(8, S, 480) * 1000).to(dt).pin_memory()
out = [torch.empty(S, 480, dtype=dt).pin_memory() for _ in range(4)]
vad = [torch.empty(S).pin_memory() for _ in range(4)]
fn = b.process_ptr_s16_async if s16 else b.process_ptr_async
for i in range(48):
    fn(out[i % 4].data_ptr(), pcm[i % 8].data_ptr(), vad[i % 4].data_ptr())
t = b.timeline()
names = ["h2d0", "h2d1", "bq1", "pitch1", "front1", "back0", "back1", "d2h1"]
print("frame " + " ".join(f"{n:>8s}" for n in names) + "   | h2d  front-after-bq  back  d2h  period(back1)")
for f in range(24, 40):
    r = t[f]
    print(f"{f:5d} " + " ".join(f"{x:8.3f}" for x in r) + f"   | {r[1]-r[0]:.3f} {r[4]-r[2]:.3f} {r[6]-r[5]:.3f} {r[7]-r[6]:.3f} {t[f][6]-t[f-1][6]:.3f}")
