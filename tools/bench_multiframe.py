#!/usr/bin/env python3
"""Offline (file) denoising throughput: rnnoise_process_frames_batch (T frames per call) against a loop
of frame-at-a-time host calls, on the same pinned host buffers.  Prints a markdown table.
usage: python tools/bench_multiframe.py [--frames 600] [--streams 16 64 256 1024]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--streams", type=int, nargs="+", default=[16, 64, 256, 1024])
    args = ap.parse_args()
    import torch
    import rnnoise_b200 as rb
    from rnnoise_b200.synth_pcm import batch_pcm
    model = rb.Model(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "models", "default.bin"))
    T = args.frames
    print("| streams | frames/call | per-frame async calls: frames/s | multi-frame float: frames/s | multi-frame int16: frames/s | x real time per stream (multi int16) |")
    print("|---:|---:|---:|---:|---:|---:|")
    for S in args.streams:
        base = batch_pcm(min(S, 64), 50)                                  # [50][<=64][480]
        sig = np.tile(base, (T // 50 + 1, S // base.shape[1] + 1, 1))[:T, :S]   # [T][S][480]
        by_frame = torch.from_numpy(np.ascontiguousarray(sig)).pin_memory()
        by_stream = torch.from_numpy(np.ascontiguousarray(sig.transpose(1, 0, 2).reshape(S, T * 480))).pin_memory()
        by_stream16 = by_stream.to(torch.int16).pin_memory()
        out_f = torch.empty_like(by_frame).pin_memory(); vad_f = torch.empty(T, S).pin_memory()
        out_s = torch.empty_like(by_stream).pin_memory(); vad_s = torch.empty(S, T).pin_memory()
        out_16 = torch.empty_like(by_stream16).pin_memory()
        L = rb.lib()
        res = []
        for mode in ("frame", "multi", "multi16"):
            b = rb.Batch(model, S)
            best = 0.0
            for rep in range(3):
                b.sync(); t0 = time.perf_counter()
                if mode == "frame":
                    for f in range(T):
                        b.process_ptr_async(out_f[f].data_ptr(), by_frame[f].data_ptr(), vad_f[f].data_ptr())
                    b.sync()
                elif mode == "multi":
                    assert L.rnnoise_process_frames_batch(b.handle, out_s.data_ptr(), by_stream.data_ptr(), vad_s.data_ptr(), T) == 0
                else:
                    assert L.rnnoise_process_frames_batch_s16(b.handle, out_16.data_ptr(), by_stream16.data_ptr(), vad_s.data_ptr(), T) == 0
                dt = time.perf_counter() - t0
                best = max(best, S * T / dt)
            res.append(best)
            b.destroy()
        print(f"| {S} | {T} | {res[0]:.4g} | {res[1]:.4g} | {res[2]:.4g} | {res[2] / S / 100:.1f} |")


if __name__ == "__main__":
    main()
