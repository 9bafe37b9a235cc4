#!/usr/bin/env python3
"""Generates tests/golden/ref_<model>.npz from the UNMODIFIED reference build (oracle/_ref, RTCD/AVX2
path on this container's Intel Xeon) -- run in the build container, where /root/reference exists:

    python oracle/make_models.py && python oracle/build_ref.py && python tests/golden/make_golden.py

Per model and per test stream: FRAMES frames of rnnoise_b200.synth_pcm.stream_pcm(stream, FRAMES),
with per-frame features[65], pitch period, silence flag, Ex[32], raw network gains[32], VAD and the
denoised PCM[480] of the reference; plus the same run of the generic-C build (true division in the
activations), which defines the tolerance envelope E_ref used by the parity tests.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refbind import RefLib  # noqa: E402
from rnnoise_b200.synth_pcm import stream_pcm  # noqa: E402

FRAMES = 60
STREAMS = (0, 15, 7)
HERE = os.path.dirname(os.path.abspath(__file__))


def run(lib, pcm):
    st = lib.create()
    keys = ("features", "Ex", "g_raw", "out", "lastg")
    acc = {k: [] for k in keys}
    acc.update(pitch=[], silence=[], vad=[])
    for f in range(pcm.shape[0]):
        t = lib.process_frame_traced(st, pcm[f])
        for k in keys:
            acc[k].append(t[k])
        acc["pitch"].append(t["pitch"]); acc["silence"].append(t["silence"]); acc["vad"].append(t["vad"])
    lib.destroy(st)
    return {k: np.asarray(v) for k, v in acc.items()}


def main():
    for name in ("default", "hot", "little"):
        mp = os.path.join(HERE, "models", name + ".bin")
        rt, ge = RefLib(mp, "rtcd"), RefLib(mp, "generic")
        out = {}
        for s in STREAMS:
            # 15 is a "gap" stream: use enough frames to enter digital silence
            pcm = stream_pcm(s, FRAMES)
            a, b = run(rt, pcm), run(ge, pcm)
            for k, v in a.items():
                out[f"s{s}_{k}"] = v
            for k in ("g_raw", "out", "vad", "lastg"):
                out[f"s{s}_generic_{k}"] = b[k]
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, frames=FRAMES, streams=np.array(STREAMS), **out)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
