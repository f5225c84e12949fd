"""Synthetic 16 kHz noisy-speech clips: the fixed recipe of SURVEY 8(d) used by the tests and bench.py.

clean = sum_{h=1..20} (1/h) sin(2 pi h f0 t + phi_h) * 0.5 (1 + sin(2 pi 4 t + phi_e)),  f0 ~ U(90, 250) Hz;
noise = white gaussian scaled to an SNR drawn from {0, 5, 10, 15} dB (README.md:46); the mix is peak-normalised
with the reference's `minMaxNorm` (dataloader/dataloader.py:11-15) and clipped to [-1, 1] (:71)."""
from __future__ import annotations

import numpy as np

FS = 16000


def min_max_norm(wav: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    mx, mn = np.max(np.abs(wav)), np.min(np.abs(wav))
    return np.clip((wav - mn) / (mx - mn + eps), -1.0, 1.0)


def synth_clip(index: int, n_samples: int) -> np.ndarray:
    rng = np.random.default_rng(1234 + index)
    t = np.arange(n_samples) / FS
    f0 = rng.uniform(90.0, 250.0)
    clean = np.zeros(n_samples)
    for h in range(1, 21):
        clean += np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi)) / h
    clean *= 0.5 * (1.0 + np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi)))
    noise = rng.standard_normal(n_samples)
    snr = rng.choice([0.0, 5.0, 10.0, 15.0])
    noise *= np.sqrt(np.mean(clean ** 2) / (np.mean(noise ** 2) * 10 ** (snr / 10)))
    return min_max_norm(clean + noise).astype(np.float32)


def synth_clips(batch: int, n_samples: int, first_clip: int = 0) -> np.ndarray:
    return np.stack([synth_clip(first_clip + i, n_samples) for i in range(batch)])
