#!/bin/bash
# end-to-end demo of the front door: N synthetic 30 s WAV files (mixed 16 / 24 kHz) -> transcribe_batch -> JSONL
N=${1:-128}
DIR=$(mktemp -d)
python - "$DIR" "$N" <<'PY'
import struct, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "qwen3-asr-swift_b200"))
from q3asr import synth
d, n = sys.argv[1], int(sys.argv[2])
for i in range(n):
    rate = 24000 if i % 4 == 3 else 16000
    x = synth.clip(i, 30 * rate)
    pcm = np.clip(np.round(x * 32767.0), -32768, 32767).astype("<i2")
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + pcm.nbytes, b"WAVE", b"fmt ", 16, 1, 1, rate, rate * 2, 2, 16, b"data", pcm.nbytes)
    open(os.path.join(d, f"clip{i:04d}.wav"), "wb").write(hdr + pcm.tobytes())
PY
g++ -std=c++17 -O2 -o qwen3-asr-swift_b200/build/transcribe_batch qwen3-asr-swift_b200/host/transcribe_batch.cpp -Lqwen3-asr-swift_b200/lib -lq3asr -Wl,-rpath,$PWD/qwen3-asr-swift_b200/lib || exit 1
qwen3-asr-swift_b200/build/transcribe_batch "$DIR" --jsonl --max-tokens 128 --batch 64 > gpurun_out/cli_demo.log 2>&1; echo "exit $?"
grep -c '"text"' gpurun_out/cli_demo.log; tail -n 5 gpurun_out/cli_demo.log
rm -rf "$DIR"
