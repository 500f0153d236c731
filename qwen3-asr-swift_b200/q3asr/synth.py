"""Synthetic 16 kHz audio for benchmarks and parity runs: input data only, no model arithmetic.

Recipe of the reference's own fixture generator
(/root/reference/scripts/kws/generate_fbank_reference.py:33-40):
0.4*sin(2*pi*440 t) + 0.2*sin(2*pi*1200 t) + 0.05*N(0,1), default_rng(20260418 + i).
bench.py, the tools and (through oracle/synth.py) the tests share it.
"""
import numpy as np

SEED = 20260418


def clip(i, n_samples, sr=16000):
    rng = np.random.default_rng(SEED + int(i))
    t = np.arange(n_samples, dtype=np.float64) / sr
    x = 0.4 * np.sin(2 * np.pi * 440.0 * t) + 0.2 * np.sin(2 * np.pi * 1200.0 * t)
    x = x + 0.05 * rng.standard_normal(n_samples)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
