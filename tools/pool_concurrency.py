#!/usr/bin/env python
"""Aggregate throughput of one GPU when the pool runs W workers (handles, streams) on it: batches of 64 x 30 s clips, 128 tokens.
W = 1 runs the batches one after the other; W >= 2 overlaps the latency-bound decode steps of different batches.
Usage: python tools/pool_concurrency.py [clips] [workers ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
workers = [int(a) for a in sys.argv[2:]] or [1, 2, 3]
max_batch = int(os.environ.get("POOL_MAX_BATCH", "64"))
clips = [synth.clip(i % 64, 480000) for i in range(n)]
ref = None
for w in workers:
    pool = q3asr.Pool("0.6B", devices=(0,) * w)
    pool.transcribe_ids(clips[:max_batch * w], max_tokens=128, stop_on_eos=False, max_batch_per_gpu=max_batch)  # warm-up
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        out = pool.transcribe_ids(clips, max_tokens=128, stop_on_eos=False, max_batch_per_gpu=max_batch)
        best = min(best, time.perf_counter() - t0)
    pool.close()
    ids = [t.tolist() for t in out]
    ref = ref or ids
    print(f"max_batch {max_batch} workers {w}: {n} clips in {best * 1000:.0f} ms = {n * 30 / best:.0f} audio-s/s  ids_same {ids == ref}", flush=True)
