#!/usr/bin/env python
"""Stage times (encoder / prefill) of the 64 x 30 s workload under an env knob, e.g. Q3ASR_ATTN_NO_PAIR=1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

m = q3asr.Qwen3ASRModel.random_init("0.6B")
x = [synth.clip(i, 480000) for i in range(64)]
m.batch_upload(x)
best = None
for _ in range(4):
    m.batch_run(q3asr.STAGE_MEL | q3asr.STAGE_ENCODER | q3asr.STAGE_PREFILL, 1, False)
    m.sync()
    m.batch_download(64, 1)
    st = m.stage_ms()
    best = st if best is None or st[2] < best[2] else best
print({k: os.environ.get(k) for k in os.environ if k.startswith("Q3ASR_")}, "mel/enc/prefill ms:", [round(float(v), 2) for v in best[:3]])
m.close()
