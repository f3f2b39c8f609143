"""Synthetic 16 kHz test audio (SURVEY.md section 8d "Synthetic inputs").

Even clip indices are white noise 0.1*N(0,1) from ``torch.Generator(1234+idx)``;
odd indices are constant-amplitude linear chirps 0.5*sin(2*pi*(f0*t + (f1-f0)*t^2/(2D))).
Both have a constant envelope, so ``trim(top_db=20)`` is the identity and frame
counts are a function of the length alone.
"""
from __future__ import annotations

import numpy as np
import torch

SR = 16000


def noise(idx: int, n: int) -> np.ndarray:
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + idx)
    return (0.1 * torch.randn(n, generator=g, dtype=torch.float32)).numpy()


def chirp(idx: int, n: int) -> np.ndarray:
    rng = np.random.RandomState(4321 + idx)
    f0 = rng.uniform(50.0, 500.0)
    f1 = rng.uniform(2000.0, 7500.0)
    t = np.arange(n, dtype=np.float64) / SR
    dur = max(n / SR, 1e-9)
    return (0.5 * np.sin(2.0 * np.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * dur)))).astype(np.float32)


def clip(idx: int, n: int) -> np.ndarray:
    return noise(idx, n) if idx % 2 == 0 else chirp(idx, n)


def mixed(idx: int, n: int) -> np.ndarray:
    """Chirp over a -34 dB noise floor (no numerically empty spectral bins), or plain noise for even indices."""
    if idx % 2 == 0:
        return noise(idx, n)
    return (chirp(idx, n) + 0.1 * noise(idx, n)).astype(np.float32)


def ragged_lengths(n_clips: int, lo_s: float = 3.0, hi_s: float = 30.0, seed: int = 2024) -> np.ndarray:
    rng = np.random.RandomState(seed)
    return np.round(SR * rng.uniform(lo_s, hi_s, size=n_clips)).astype(np.int64)


def with_silence(idx: int, n: int, lead: int, tail: int) -> np.ndarray:
    """A clip with leading/trailing near-silence, so that trim() is not the identity."""
    y = clip(idx, n).copy()
    y[:lead] *= 1e-4
    if tail:
        y[-tail:] *= 1e-4
    return y
