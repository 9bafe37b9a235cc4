#!/usr/bin/env python3
"""Per-phase clock64() profile of one k_pitch2 CTA (diagnostics builds: build.py --variant X --flags "-DPITCH_TIMING=<cta>").
usage: RNNOISE_B200_LIB_PATH=rnnoise_b200/librnnoise_b200_pgt.so python tools/pitch_timing.py [streams]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

os.environ["RNNOISE_B200_LANES"] = "1"
os.environ["RNNOISE_B200_OVERLAP"] = "0"
import rnnoise_b200  # noqa: E402
from rnnoise_b200.synth_pcm import batch_pcm  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L = rnnoise_b200.lib()
L.b200_debug_pitch_timing.argtypes = [C.POINTER(C.c_longlong), C.c_int]
base = batch_pcm(32, 12)
pcm = np.ascontiguousarray(np.tile(base, (1, S // 32, 1)))
model = rnnoise_b200.Model(os.path.join(ROOT, "tests", "golden", "models", "default.bin"))
b = rnnoise_b200.Batch(model, S)
names = ["P1 append+decimate", "P2 autocorr", "P3 lpc", "P4 fir", "P5 decimate2", "P6 coarse|chains", "P7 scan", "P8 fine", "P9 pick",
         "P10 rd dots", "P11a gains", "P11b decision", "P12 refine", "P13 final"]
acc = np.zeros(len(names))
n = 0
for f in range(12):
    b.process(pcm[f])
    buf = (C.c_longlong * 32)()
    assert L.b200_debug_pitch_timing(buf, 32) == 32
    t = np.array(buf[:len(names) + 1], np.float64)
    if f >= 4:
        acc += np.diff(t); n += 1
acc /= n
print(json.dumps({"lib": os.path.basename(rnnoise_b200.LIB_PATH), "streams": S, "cycles_per_phase": {k: round(v) for k, v in zip(names, acc)},
                  "total_cycles": round(float(acc.sum())), "total_us_at_1965MHz": round(float(acc.sum()) / 1965, 1)}))
b.destroy(); model.free()
