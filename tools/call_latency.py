#!/usr/bin/env python
"""Wall-clock latency of the blocking call at small batches (median of 9): what the per-call fixed costs (step-graph build, planning,
upload) add up to.  Usage: python tools/call_latency.py [tokens]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import numpy as np  # noqa: E402
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = q3asr.Qwen3ASRModel.random_init("0.6B")
clips = [synth.clip(i, 480000) for i in range(8)]
for n, tk in ((1, 2), (1, tokens), (8, tokens)):
    lat = []
    ref = None
    for _ in range(10):
        t0 = time.perf_counter()
        ids = m.transcribe_ids(clips[:n], max_tokens=tk, stop_on_eos=False)
        lat.append(time.perf_counter() - t0)
        ref = ref or [t.tolist() for t in ids]
        assert [t.tolist() for t in ids] == ref
    print(f"{n} clip(s), {tk:3d} tokens: median {np.median(lat[1:]) * 1000:7.2f} ms  min {min(lat[1:]) * 1000:7.2f} ms", flush=True)
m.close()
