"""Synthetic 48 kHz mono PCM for parity tests and benchmarks.

Per stream s: rng = default_rng(20260922 + s); voiced harmonic source with an f0 glide
U(80,400) Hz +-20 % at 0.5 Hz, sum_{k<20} sin(k*phi)/k, gated by a 1.5 Hz square envelope,
amplitude U(1000,8000), plus white Gaussian noise at SNR U(0,20) dB.  Every 16th stream has an
exact-zero 1 s gap (exercises the silence mask, reference src/denoise.c:389).  Samples are float32
in int16 units rounded to integers, as examples/rnnoise_demo.c:53-56 feeds them.
"""
import numpy as np

FRAME = 480
FS = 48000.0


def stream_pcm(s, frames, base_seed=20260922):
    rng = np.random.default_rng(base_seed + int(s))
    n = frames * FRAME
    t = np.arange(n) / FS
    f0 = rng.uniform(80.0, 400.0)
    f = f0 * (1.0 + 0.2 * np.sin(2 * np.pi * 0.5 * t + rng.uniform(0, 2 * np.pi)))
    phi = 2 * np.pi * np.cumsum(f) / FS
    v = np.zeros(n)
    for k in range(1, 20):
        v += np.sin(k * phi) / k
    gate = (np.sin(2 * np.pi * 1.5 * t + rng.uniform(0, 2 * np.pi)) > 0).astype(np.float64)
    amp = rng.uniform(1000.0, 8000.0)
    sig = amp * gate * v
    snr_db = rng.uniform(0.0, 20.0)
    p_sig = max(np.mean(sig ** 2), 1.0)
    noise = rng.standard_normal(n) * np.sqrt(p_sig / 10 ** (snr_db / 10))
    x = sig + noise
    if s % 16 == 15 and n >= 2 * int(FS):
        x[int(FS // 2):int(FS // 2) + int(FS)] = 0.0  # 1 s of digital silence
    elif s % 16 == 15:
        x[n // 3: 2 * n // 3] = 0.0
    x = np.clip(np.round(x), -32768, 32767)
    return x.astype(np.float32).reshape(frames, FRAME)


def batch_pcm(streams, frames, base_seed=20260922, first_stream=0):
    """float32 [frames][streams][480] (frame-major: one contiguous [S][480] block per step)."""
    out = np.empty((frames, streams, FRAME), np.float32)
    for i in range(streams):
        out[:, i, :] = stream_pcm(first_stream + i, frames, base_seed)
    return out
