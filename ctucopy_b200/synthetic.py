"""Deterministic synthetic 16 kHz speech-like audio.

Utterance k: rng = default_rng(1000 + k); content is one of
  0: three harmonic tones (f0 in [90, 250] Hz) with 4 Hz amplitude modulation,
  1: a linear chirp 100 -> 4000 Hz,
  2: tone + chirp,
peak about 8000, gated on after a 0.30 s noise-only lead-in; pink noise (1/f shaped white)
at an SNR from {0, 5, 10, 20} dB, or a -60 dBFS pink floor on "clean" items.  Never
digital silence (the reference emits -inf/NaN on all-zero frames, src/fea/fea_impl.cc:109).
Output is int16, to be used with `-dither 0`.
"""
from __future__ import annotations

import numpy as np

SNRS_DB = (0.0, 5.0, 10.0, 20.0, None)  # None = "clean" (-60 dBFS floor)


def _pink(rng: np.random.Generator, n: int) -> np.ndarray:
    w = rng.standard_normal(n)
    F = np.fft.rfft(w)
    f = np.arange(len(F), dtype=np.float64)
    f[0] = 1.0
    F /= np.sqrt(f)
    y = np.fft.irfft(F, n)
    return y / (np.sqrt(np.mean(y * y)) + 1e-30)


def utterance(k: int, seconds: float = 10.0, fs: int = 16000) -> np.ndarray:
    rng = np.random.default_rng(1000 + k)
    n = int(round(seconds * fs))
    t = np.arange(n, dtype=np.float64) / fs
    kind = k % 3
    sig = np.zeros(n)
    if kind in (0, 2):
        f0 = rng.uniform(90.0, 250.0)
        am = 1.0 + 0.5 * np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi))
        for h, a in ((1, 1.0), (2, 0.6), (3, 0.35)):
            sig += a * am * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi))
    if kind in (1, 2):
        T = max(t[-1], 1e-9)
        ph = 2 * np.pi * (100.0 * t + 0.5 * (4000.0 - 100.0) / T * t * t)
        sig += np.sin(ph)
    sig *= 8000.0 / (np.max(np.abs(sig)) + 1e-30)
    lead = min(int(0.30 * fs), n // 3)
    gate = np.ones(n)
    gate[:lead] = 0.0
    ramp = min(int(0.01 * fs), n - lead)
    if ramp > 0:
        gate[lead:lead + ramp] = np.linspace(0.0, 1.0, ramp)
    sig *= gate
    snr = SNRS_DB[(k // 3) % len(SNRS_DB)]
    noise = _pink(rng, n)
    if snr is None:
        noise *= 32768.0 * 10 ** (-60.0 / 20.0)
    else:
        ps = np.mean(sig[lead:] ** 2) if n > lead else 1.0
        noise *= np.sqrt(ps / (10 ** (snr / 10.0)))
    x = np.clip(np.round(sig + noise), -32768, 32767)
    return x.astype(np.int16)


def batch(n_utts: int, seconds: float = 10.0, fs: int = 16000, unique: int = 0, first: int = 0):
    """Returns (pcm int16 [sum N], lengths int64 [n_utts]).  With unique > 0 only that many
    distinct utterances are synthesised and tiled (the throughput set: 10 000 x 10 s would
    otherwise take minutes of host time to generate).  `first`: number of the first utterance
    synthesised (a shard of a larger list: rank r of the benchmark starts at 16 r)."""
    u = n_utts if unique <= 0 else min(unique, n_utts)
    base = [utterance(first + k, seconds, fs) for k in range(u)]
    pcm = np.concatenate([base[i % u] for i in range(n_utts)]) if n_utts else np.zeros(0, np.int16)
    lens = np.array([len(base[i % u]) for i in range(n_utts)], dtype=np.int64)
    return pcm, lens
