"""Generates the committed golden fixtures from the CPU oracle (run here, on CPU):
    python tests/golden/make_golden.py [--big]
mel_golden.npz      inputs + oracle log-mel for four small clips
tiny_golden.npz     encoder output + greedy ids + teacher-forced argmax of the tiny configuration
q06b_clip5s.npz     (--big) greedy ids + encoder slice of Qwen3-ASR-0.6B dims on one 5 s clip
The reference itself cannot be imported (Swift + MLX); these are outputs of the restatement in oracle/.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mel as omel  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import synth, weights  # noqa: E402


def main():
    out = {}
    imp = np.zeros(4000, np.float32)
    imp[2000] = 1.0
    for name, x in (("mel_clip0_1600", synth.clip(0, 1600)), ("mel_clip1_16000", synth.clip(1, 16000)),
                    ("mel_zeros_3200", np.zeros(3200, np.float32)), ("mel_impulse_4000", imp)):
        out[name + "_x"] = x
        out[name + "_y"] = omel.mel(x)
    np.savez_compressed(os.path.join(HERE, "mel_golden.npz"), **out)

    cfg = weights.preset("tiny")
    orc = omodel.Oracle(cfg, weights.random_state_dict(cfg, 20260418))
    x = synth.clip(0, 16000 * 3 + 777)
    enc = orc.encode(omel.mel(x))
    ids, tops, margins = orc.greedy(enc, 32, stop_on_eos=False)
    forced = np.random.default_rng(11).integers(0, 2000, size=24).astype(np.int32)
    fids, ftops, fmargins = orc.greedy(enc, 0, forced=forced)
    np.savez_compressed(os.path.join(HERE, "tiny_golden.npz"), seed=20260418, clip_index=0, n_samples=x.size, encoder=enc, ids=ids,
                        tops=tops, margins=margins, forced=forced, forced_ids=fids, forced_tops=ftops, forced_margins=fmargins)

    if "--big" in sys.argv:
        cfg = weights.preset("0.6B")
        orc = omodel.Oracle(cfg, weights.random_state_dict(cfg, 20260418))
        x = synth.clip(7, 80000)
        enc = orc.encode(omel.mel(x))
        ids, tops, margins = orc.greedy(enc, 24, stop_on_eos=False)
        np.savez_compressed(os.path.join(HERE, "q06b_clip5s.npz"), seed=20260418, clip_index=7, n_samples=x.size,
                            encoder_first64=enc[:, :64], ids=ids, tops=tops, margins=margins)
        print("0.6B ids", ids.tolist(), "min margin", margins.min())


if __name__ == "__main__":
    main()
