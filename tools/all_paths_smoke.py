#!/usr/bin/env python
"""A short pass over every path of the tiny configurations (compute-sanitizer is closed on this pool, so this is a plain smoke run):
transcription (greedy, decoder knobs, other sample rates), long-form windows, forced aligner, sampler hook, mel edge sizes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402

m = q3asr.Qwen3ASRModel.random_init("tiny")
clips = [synth.clip(i, n) for i, n in enumerate([16000 * 2 + 77, 1600, 161, 40000])]
for c in clips + [synth.clip(9, 5121)]:
    m.extract_features(c)
print("greedy", [t.tolist()[:4] for t in m.transcribe_ids(clips, max_tokens=6, stop_on_eos=False)])
o = q3asr.Qwen3DecodingOptions(repetition_penalty=1.3, no_repeat_ngram_size=2, temperature=0.5, seed=1)
print("knobs", [t.tolist()[:4] for t in m.transcribe_ids(clips, max_tokens=6, stop_on_eos=False, options=o)])
x24 = [synth.clip(i, 24000 + 100 * i) for i in range(2)]
print("24k", [t.tolist()[:4] for t in m.transcribe_ids(x24, max_tokens=4, stop_on_eos=False, sample_rates=[24000, 24000])])
print("long", len(m.transcribe_long(synth.clip(3, 16000 * 3 + 50), window_seconds=1.0, max_tokens=3, batch=2)))
print("pick", m.pick_next_token(np.arange(100, dtype=np.float32), [99, 98], q3asr.Qwen3DecodingOptions(no_repeat_ngram_size=1)))
m.close()
a = q3asr.Qwen3ASRModel.random_init("tiny-aligner")
print("align", a.align_indices([clips[0], clips[3]], [[2007, 5, 2007], [2007, 6, 7, 2007, 2007, 8, 2007]], [[0, 2], [0, 3, 4, 6]]))
a.close()
p = q3asr.Pool("tiny", devices=(0, 0))
print("pool", [t.tolist()[:3] for t in p.transcribe_ids(clips, max_tokens=4, stop_on_eos=False, max_batch_per_gpu=2)])
p.close()
print("done")
