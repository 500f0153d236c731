#!/usr/bin/env python
"""Timing ablation of the decode step (debug): drops one kernel of the layer at a time (Q3ASR_DEC_SKIP bit mask; results
are wrong while a bit is set) and reports the decode-stage time per step.  Needs the ablation build of the library
(`make -C qwen3-asr-swift_b200 ablation` -> lib/libq3asr_ablation.so): the shipped libq3asr.so has no such switch.
Usage: python tools/decode_ablation.py [clips] [tokens] [size] [seconds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
q3asr.LIB_PATH = os.path.join(os.path.dirname(q3asr.LIB_PATH), "libq3asr_ablation.so")  # before the first call loads the library
from q3asr import synth  # noqa: E402  (input data only)

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = sys.argv[3] if len(sys.argv) > 3 else "0.6B"
seconds = int(sys.argv[4]) if len(sys.argv) > 4 else 30
m = q3asr.Qwen3ASRModel.random_init(size)
x = [synth.clip(i, 16000 * seconds) for i in range(clips)]
m.batch_upload(x)
names = ["none", "qkv", "attn", "o", "norm1", "gateup", "down", "norm2", "all-gemm", "all-but-attn", "everything"]
masks = [0, 1, 2, 4, 8, 16, 32, 64, 1 | 4 | 16 | 32, 1 | 4 | 8 | 16 | 32 | 64, 127]
base = None
for name, mask in zip(names, masks):
    os.environ["Q3ASR_DEC_SKIP"] = str(mask)
    best = 1e9
    for _ in range(3):
        m.batch_run(q3asr.STAGE_ALL, tokens, False)
        m.sync()
        m.batch_download(clips, tokens)
        best = min(best, m.stage_ms()[3])
    per = best / (tokens - 1) * 1000.0
    base = base or per
    print(f"skip {name:14s} decode {best:8.2f} ms  {per:8.1f} us/step  {per / 28:6.2f} us/layer  delta {(base - per) / 28:6.2f} us/layer", flush=True)
m.close()
