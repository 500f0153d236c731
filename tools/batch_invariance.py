#!/usr/bin/env python
"""Are greedy ids independent of the batch an utterance is decoded in?  One handle, ragged 0.6B clips: a batch of N against the same
clips in sub-batches of 64 / 40 / 1.  Kernel configurations differ between batch sizes (decode attention: 2 warps per (sequence, head)
from 37 sequences up, 8 below; UMMA N of the decode GEMMs), so fp32 summation order can differ across them.
Usage: python tools/batch_invariance.py [n] [tokens]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rng = np.random.default_rng(2)
clips = [synth.clip(i, int(rng.uniform(0.3, 30.0) * 16000)) for i in range(n)]
m = q3asr.Qwen3ASRModel.random_init("0.6B")
full = m.transcribe_ids(clips, max_tokens=tokens, stop_on_eos=False)
for sub in (64, 40, 16, 1):
    bad = 0
    idx = range(n) if sub > 1 else range(0, n, 5)
    for s in (range(0, n, sub) if sub > 1 else idx):
        part = m.transcribe_ids(clips[s:s + sub], max_tokens=tokens, stop_on_eos=False)
        for k, t in enumerate(part):
            bad += int(t.tolist() != full[s + k].tolist())
    total = n if sub > 1 else len(list(idx))
    print(f"batch of {n} vs sub-batches of {sub}: {bad} of {total} utterances differ", flush=True)
m.close()
