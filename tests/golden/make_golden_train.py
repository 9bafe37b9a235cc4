#!/usr/bin/env python3
"""Generates tests/golden/ref_train.npz: training-feature records of the UNMODIFIED reference built with
-DTRAINING=1 (oracle/_ref/librnnoise_ref_training.so, see oracle/build_ref.py and oracle/ref_train.c) on
the seeded synthetic (clean, noisy) pairs of rnnoise_b200.synth_pcm.train_pair.  Run here, where
/root/reference exists; the GPU box only reads the .npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import build_ref, trainbind  # noqa: E402
from rnnoise_b200.synth_pcm import train_pair, train_params  # noqa: E402

STREAMS = [0, 1, 2, 3, 4, 9, 15, 20, 33, 79]   # 79: noise-free with a digital-silence gap -> quiet frames
FRAMES = 50


def vad_target(f, s):
    return float((f // 7 + s) % 2)


def main():
    build_ref.build()
    rec = np.zeros((FRAMES, len(STREAMS), 98), np.float32)
    quiet = np.zeros((FRAMES, len(STREAMS)), np.int32)
    for q, s in enumerate(STREAMS):
        clean, noisy = train_pair(s, FRAMES)
        lp, blp, nf = train_params(s)
        r = trainbind.RefTrain()
        for f in range(FRAMES):
            rec[f, q], quiet[f, q], _ = r.frame(clean[f], noisy[f], vad_target(f, s), nf, lp, blp)
        r.close()
    out = os.path.join(HERE, "ref_train.npz")
    np.savez_compressed(out, streams=np.array(STREAMS), frames=FRAMES, rec=rec, quiet=quiet)
    print(out, os.path.getsize(out), "bytes; undefined gains:", int(np.sum(rec[:, :, 65:97] == -1)), "quiet frames:", int(quiet.sum()))


if __name__ == "__main__":
    main()
