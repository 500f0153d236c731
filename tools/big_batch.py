#!/usr/bin/env python
"""Throughput of ONE blocking q3asr_transcribe_ids call as the request grows past the 128 decode rows the weight-streaming decode
step is built for.  Usage: python tools/big_batch.py [tokens]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = q3asr.Qwen3ASRModel.random_init("0.6B")
clips = [synth.clip(i, 480000) for i in range(256)]
ref = None
for n in (64, 128, 160, 256):
    m.transcribe_ids(clips[:n], max_tokens=tokens, stop_on_eos=False)
    t0 = time.perf_counter()
    ids = m.transcribe_ids(clips[:n], max_tokens=tokens, stop_on_eos=False)
    sec = time.perf_counter() - t0
    if ref is None:
        ref = [t.tolist() for t in ids]
    same = all(ids[i].tolist() == ref[i] for i in range(len(ref)))
    print(f"{n:4d} clips: {sec * 1000:8.1f} ms  {n * 30 / sec:8.0f} audio-s/s  first 64 ids as in the 64-clip call: {same}", flush=True)
m.close()
