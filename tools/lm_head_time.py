#!/usr/bin/env python
"""LM-head time per launch (profiling mode, CUDA events): general kernel vs lmhead.cuh.
Usage: python tools/lm_head_time.py [size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402  (input data only)

size = sys.argv[1] if len(sys.argv) > 1 else "0.6B"
m = q3asr.Qwen3ASRModel.random_init(size)
m.batch_upload([synth.clip(i, 16000 * 5) for i in range(64)])
ref = None
for name, env in (("general kernel", dict(Q3ASR_LM_GENERAL="1")), ("lmhead.cuh", dict(Q3ASR_LM_GENERAL="0")), ("general kernel", dict(Q3ASR_LM_GENERAL="1"))):
    os.environ.update(env)
    m.profile(True)
    for _ in range(5):
        m.batch_run(q3asr.STAGE_ALL, 3, False)
        m.sync()
    rep = m.profile_report()
    m.profile(False)
    ids = m.batch_download(64, 3)
    flat = [t.tolist() for t in ids]
    ref = ref or flat
    r = rep["lm_head"]
    print(f"{size} {name:18s} lm_head {r['ms'] / r['launches'] * 1000:7.2f} us/launch ({r['launches']} launches)  ids_same {flat == ref}", flush=True)
m.close()
