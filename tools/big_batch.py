#!/usr/bin/env python
"""Throughput of ONE blocking q3asr_transcribe_ids call as the request grows (the weight-streaming decode step takes up to 256 rows;
larger requests are served as equal sub-batches), and every utterance's ids against those it got in the other requests.  Usage: python tools/big_batch.py [tokens]"""
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
clips = [synth.clip(i, 480000) for i in range(300)]
ref = {}
for n in (64, 128, 256, 160, 200, 300):
    m.transcribe_ids(clips[:n], max_tokens=tokens, stop_on_eos=False)
    t0 = time.perf_counter()
    ids = m.transcribe_ids(clips[:n], max_tokens=tokens, stop_on_eos=False)
    sec = time.perf_counter() - t0
    for i, t in enumerate(ids):  # every utterance's ids must not depend on the request it arrived in
        ref.setdefault(i, t.tolist())
    same = all(ids[i].tolist() == ref[i] for i in range(n))
    print(f"{n:4d} clips: {sec * 1000:8.1f} ms  {n * 30 / sec:8.0f} audio-s/s  ids as in the earlier (smaller or chunked) requests: {same}", flush=True)
m.close()
