#!/usr/bin/env python
"""Decode-stage time per step (best of 3) for one configuration; environment knobs are read by the library at first use, so
tuning sweeps run this once per setting.  Usage: python tools/decode_time.py [size] [clips] [seconds] [tokens] [label]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("Q3ASR_PKG") or os.path.join(ROOT, "qwen3-asr-swift_b200"))  # Q3ASR_PKG: A/B against another build
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

size = sys.argv[1] if len(sys.argv) > 1 else "0.6B"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tokens = int(sys.argv[4]) if len(sys.argv) > 4 else 128
label = sys.argv[5] if len(sys.argv) > 5 else ""
m = q3asr.Qwen3ASRModel.random_init(size)
m.batch_upload([synth.clip(i, 16000 * secs) for i in range(clips)])
best = 1e9
for _ in range(3):
    m.batch_run(q3asr.STAGE_ALL, tokens, False)
    m.sync()
    ids = m.batch_download(clips, tokens)
    best = min(best, m.stage_ms()[3])
per = best / (tokens - 1) * 1000.0
print(f"{size} {label:28s} decode {best:8.2f} ms  {per:8.1f} us/step  {per / 28:6.2f} us/layer  ids[0][:6] {ids[0][:6].tolist()}", flush=True)
m.close()
