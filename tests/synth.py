"""Seeded synthetic songs for the parity tests (numpy, host side).

Signals follow SURVEY.md §8d: band-limited noise + a few sinusoids + amplitude-modulated bursts at a
per-song tempo, zero-mean, peak <= 0.5 full scale, never silent at either end.
"""
import numpy as np


def song_f32(seed, seconds, rate=44100):
    """44.1 kHz mono float32 in [-0.5, 0.5]."""
    rng = np.random.default_rng(0x5EED0000 + seed)
    n = int(seconds * rate)
    t = np.arange(n, dtype=np.float64) / rate
    noise = rng.standard_normal(n)
    # crude band limit: moving average of width 3..9
    w = int(rng.integers(3, 10))
    noise = np.convolve(noise, np.ones(w) / w, mode="same")
    sig = rng.uniform(0.02, 0.12) * noise / (noise.std() + 1e-12)
    for _ in range(int(rng.integers(2, 5))):
        f = rng.uniform(60, 8000)
        sig += rng.uniform(0.01, 0.08) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
    bpm = rng.uniform(60, 180)
    phase = (t * bpm / 60.0) % 1.0
    burst = np.exp(-phase * rng.uniform(4, 12))
    sig += rng.uniform(0.05, 0.25) * burst * np.sin(2 * np.pi * rng.uniform(80, 400) * t)
    sig -= sig.mean()
    peak = np.abs(sig).max()
    if peak > 0.5:
        sig *= 0.5 / peak
    return sig.astype(np.float32)


def song_s16(seed, seconds, decorrelate=False, gain=1.0):
    """int16 / 22 050 Hz / stereo interleaved, the analysers' native input."""
    x = song_f32(seed, seconds, rate=22050).astype(np.float64) * gain
    left = np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
    if decorrelate:
        rng = np.random.default_rng(0xABCD + seed)
        right = np.clip(left.astype(np.int32) + rng.integers(-400, 401, len(left)), -32768, 32767).astype(np.int16)
    else:
        right = left
    return np.stack([left, right], axis=1).reshape(-1)
