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


def train_pair(stream, frames):
    """Deterministic (clean, noisy) float32 [frames][480] pair for the training-feature tests: the clean
    signal is one synthetic stream, the noisy one adds a second, scaled, stream (float32 elementwise
    arithmetic only, so every machine produces the same bits).  Every 5th stream is noise-free."""
    clean = stream_pcm(stream, frames) * np.float32(0.5)
    if stream % 5 == 4:
        return clean, clean.copy()
    noise = stream_pcm(stream + 7919, frames) * np.float32(0.25)
    return clean, (clean + noise).astype(np.float32)


def train_params(stream):
    """(lowpass bin, band_lp, noise_free) per stream, spanning the tool's range (dump_features.c:400-406)."""
    eband = [0, 2, 4, 6, 8, 10, 12, 15, 18, 21, 24, 28, 32, 36, 41, 47, 53, 60, 68, 77, 87, 98, 110, 124, 140, 157, 176, 198, 223,
             251, 282, 317, 356, 400]
    lowpass = [481, 60, 120, 200, 300, 481, 90][stream % 7]
    band_lp = next((i for i in range(32) if eband[i] > lowpass), 32)
    return lowpass, band_lp, int(stream % 5 == 4)
