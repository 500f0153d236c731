#!/usr/bin/env python
"""Stress run: many ragged utterances (0.2-30 s, mixed sample rates) through the pool on the 0.6B configuration, a sample of them
checked against a single handle one at a time.  Usage: python tools/stress_ragged.py [n_clips]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(1)
rates = rng.choice([16000, 16000, 24000, 8000, 44100], size=n_clips)
secs = np.concatenate([rng.uniform(0.2, 30.0, size=n_clips - 4), [0.011, 30.0, 29.99, 0.1]])
clips = [synth.clip(i, max(int(s * r), int(0.0101 * r) + 1)) for i, (s, r) in enumerate(zip(secs, rates))]
pool = q3asr.Pool("0.6B", devices=(0, 0))
t0 = time.perf_counter()
got = pool.transcribe_ids(clips, max_tokens=24, stop_on_eos=False, sample_rates=rates.tolist())
dt = time.perf_counter() - t0
pool.close()
audio = float(sum(c.size / r for c, r in zip(clips, rates)))
print(f"{n_clips} clips, {audio:.0f} s of audio in {dt:.2f} s ({audio / dt:.0f} audio-s/s incl. pool scheduling, two workers on one GPU)")
assert all(len(g) == 24 for g in got)
single = q3asr.Qwen3ASRModel.random_init("0.6B")
bad = 0
for i in list(range(0, n_clips, 13)) + [n_clips - 4, n_clips - 1]:
    want = single.transcribe_ids([clips[i]], max_tokens=24, stop_on_eos=False, sample_rates=[int(rates[i])])[0]
    if want.tolist() != got[i].tolist():
        bad += 1
        print("mismatch", i, clips[i].size, rates[i], want[:8].tolist(), got[i][:8].tolist())
single.close()
print("mismatches vs single handle:", bad)
sys.exit(1 if bad else 0)
