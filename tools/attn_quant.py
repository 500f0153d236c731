import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import q3asr
q3asr.LIB_PATH = os.path.join(os.path.dirname(q3asr.LIB_PATH), "libq3asr_ablation.so")
from q3asr import synth
m = q3asr.Qwen3ASRModel.random_init("0.6B")
tokens = 64
for clips in (37, 55, 56, 64, 74, 75, 92, 111):
    x = [synth.clip(i, 480000) for i in range(clips)]
    m.batch_upload(x)
    for name, mask in (("attn only", 1 | 4 | 8 | 16 | 32 | 64), ("full", 0)):
        os.environ["Q3ASR_DEC_SKIP"] = str(mask)
        best = 1e9
        for _ in range(3):
            m.batch_run(q3asr.STAGE_ALL, tokens, False); m.sync(); m.batch_download(clips, tokens)
            best = min(best, m.stage_ms()[3])
        per = best / (tokens - 1) * 1000.0
        print(f"{clips:4d} seqs ({clips*8:4d} items, {clips*8/148:.2f}/SM) {name:10s} {per:8.1f} us/step {per/28:6.2f} us/layer  {per/28/clips*1000:6.1f} ns/layer/seq", flush=True)
m.close()
