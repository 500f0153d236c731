#!/usr/bin/env python
"""Timeline of the persistent decode-step kernel (debug): runs one batch with Q3ASR_MEGA_TRACE and prints, per phase kind and
sub-batch, when the first / last CTA began and ended (us from the kernel's first stamp), for a few layers.
Usage: python tools/mega_trace.py [clips] [size] [seconds]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
path = "/tmp/mega_trace.bin"
os.environ["Q3ASR_MEGA_TRACE"] = path
import q3asr  # noqa: E402
from q3asr import synth  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = sys.argv[2] if len(sys.argv) > 2 else "0.6B"
seconds = int(sys.argv[3]) if len(sys.argv) > 3 else 30
os.environ["Q3ASR_MEGA"] = "1"
m = q3asr.Qwen3ASRModel.random_init(size)
x = [synth.clip(i, 16000 * seconds) for i in range(clips)]
for _ in range(2):
    m.transcribe_ids(x, max_tokens=4, stop_on_eos=False)
m.close()
raw = open(path, "rb").read()
n_phases, G, n_sub, kinds = np.frombuffer(raw[:16], dtype=np.int32)
t = np.frombuffer(raw[16:], dtype=np.uint64).reshape(n_phases, G, 2).astype(np.float64)
t0 = t[t > 0].min()
names = ["qkv", "attn", "o", "norm1", "gu", "down", "norm2"]
print(f"{n_phases} phases, {G} CTAs, {n_sub} sub-batches; kernel span {(t.max() - t0) / 1000:.1f} us")
for layer in (0, 1, 13, 27):
    for kind in range(7):
        for s in range(n_sub):
            p = (layer * 7 + kind) * n_sub + s
            b, e = t[p, :, 0], t[p, :, 1]
            ok = e > 0
            print(f"L{layer:2d} {names[kind]:5s} sub{s}: begin {((b[ok].min() - t0) / 1000):8.2f} .. {((b[ok].max() - t0) / 1000):8.2f}   "
                  f"end {((e[ok].min() - t0) / 1000):8.2f} .. {((e[ok].max() - t0) / 1000):8.2f}   busy(max) {((e[ok] - b[ok]).max() / 1000):6.2f}")
# per-kind totals: time between the last end of the previous phase in program order and the last end of this phase
ends = t[:, :, 1].max(axis=1)
prev = np.concatenate([[t0], ends[:-1]])
for kind in range(7):
    idx = [p for p in range(n_phases) if (p // n_sub) % 7 == kind]
    print(f"{names[kind]:5s}: sum of (last end - previous phase's last end) over layers and sub-batches {np.sum(ends[idx] - prev[idx]) / 1000:8.1f} us")
