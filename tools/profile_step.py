#!/usr/bin/env python
"""One short pass of the hot path for ncu: 0.6B, a few 30 s clips, a few decode steps (so the launch list stays
in the hundreds).  Usage: python tools/profile_step.py [clips] [max_tokens] [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 4
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 1
os.environ.setdefault("Q3ASR_NO_GRAPH", "1")  # ncu sees plain launches
m = q3asr.Qwen3ASRModel.random_init("0.6B")
rate = int(os.environ.get("PROFILE_RATE", "16000"))  # another rate exercises the device sample-rate converter
x = [synth.clip(i, 30 * rate) for i in range(clips)]
if os.environ.get("PROFILE_KNOBS"):  # decoder knobs: full logits + sample_kernel instead of the fused argmax
    opts = q3asr.Qwen3DecodingOptions(repetition_penalty=1.2, no_repeat_ngram_size=3, temperature=0.7, seed=1)
    for _ in range(repeats):
        ids = m.transcribe_ids(x, max_tokens=tokens, stop_on_eos=False, options=opts, sample_rates=[rate] * clips)
else:
    m.batch_upload(x) if rate == 16000 else None
    for _ in range(repeats):
        if rate != 16000:
            ids = m.transcribe_ids(x, max_tokens=tokens, stop_on_eos=False, sample_rates=[rate] * clips)
        else:
            m.batch_run(q3asr.STAGE_ALL, tokens, False)
            m.sync()
    if rate == 16000:
        ids = m.batch_download(clips, tokens)
print("ok", [t.tolist() for t in ids[:2]], "launches", m.launch_count)
m.close()
