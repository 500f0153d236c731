#!/usr/bin/env python
"""Mel stage time for the bench batch (64 x 30 s): CUDA events around the stage, best of 10.  Usage: python tools/mel_time.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

m = q3asr.Qwen3ASRModel("tiny")
clips = [synth.clip(i, 480000) for i in range(64)]
m.batch_upload(clips)
best = 1e9
for _ in range(10):
    m.flush_l2()
    m.batch_run(q3asr.STAGE_MEL, 0, False)
    m.sync()
    m.batch_download(64, 0)
    best = min(best, float(m.stage_ms()[0]))
by = sum(4.0 * c.size + 4.0 * 128 * (c.size // 160) for c in clips)
print(f"mel stage {best * 1000:.1f} us  {by / best / 1e6:.1f} GB/s  ({by / 1e6:.1f} MB algorithmic)")
m.close()
