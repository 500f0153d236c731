#!/usr/bin/env python
"""End-to-end time of q3asr_transcribe_ids from host buffers (64 x 30 s, 128 tokens), best of 5, and the upload share.
Usage: python tools/e2e_time.py [label]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

label = sys.argv[1] if len(sys.argv) > 1 else ""
m = q3asr.Qwen3ASRModel.random_init("0.6B")
clips = [synth.clip(i, 480000) for i in range(64)]
m.transcribe_ids(clips, 128, stop_on_eos=False)
best, up = 1e9, 1e9
for _ in range(5):
    t0 = time.perf_counter()
    m.transcribe_ids(clips, 128, stop_on_eos=False)
    best = min(best, time.perf_counter() - t0)
    t0 = time.perf_counter()
    m.batch_upload(clips)
    up = min(up, time.perf_counter() - t0)
print(f"{label:24s} e2e {best * 1000:7.2f} ms  ({1920 / best:7.1f} audio-s/s)  upload alone {up * 1000:6.2f} ms", flush=True)
m.close()
