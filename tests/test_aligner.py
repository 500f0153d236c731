"""Forced aligner: the timestamp-correction integer code against the reference's own unit tests
(Tests/Qwen3ASRTests/ForcedAlignerTests.swift:213-259, 441-492; C ABI and oracle restatement, no GPU), and the device path
(one prefill + classification head + argmax at the timestamp slots) against the oracle."""
import numpy as np
import pytest

import q3asr
from oracle import mel as omel
from oracle import synth
from oracle import timestamps as ots

IMPLS = [("c_abi", q3asr.enforce_monotonicity, q3asr.lis_positions, q3asr.trailing_plateau_start),
         ("oracle", ots.enforce_monotonicity, ots.lis_positions, ots.trailing_plateau_start)]


@pytest.mark.parametrize("name,fix,lis,plateau", IMPLS)
def test_timestamp_correction_reference_cases(name, fix, lis, plateau):
    assert fix([1, 3, 5, 7, 9, 11]) == [1, 3, 5, 7, 9, 11]                    # testTimestampCorrectionAlreadyMonotonic
    out = fix([1, 3, 2, 7, 9, 11])                                             # testTimestampCorrectionSingleOutOfOrder
    assert all(b >= a for a, b in zip(out, out[1:]))
    assert fix([5, 5, 5, 5]) == [5, 5, 5, 5]                                   # testTimestampCorrectionAllSame
    out = fix([10, 8, 6, 4, 2])                                                # testTimestampCorrectionDescending
    assert all(b >= a for a, b in zip(out, out[1:]))
    arr = [3, 1, 4, 1, 5, 9, 2, 6]                                             # testLISBasic
    pos = lis(arr)
    assert len(pos) >= 4 and all(arr[a] < arr[b] for a, b in zip(pos, pos[1:]))
    assert fix([]) == [] and fix([7]) == [7] and lis([]) == []


@pytest.mark.parametrize("name,fix,lis,plateau", IMPLS)
def test_trailing_plateau_reference_cases(name, fix, lis, plateau):
    healthy = [0.5 * i for i in range(20)]                                     # testNoPlateauOnHealthyAlignment
    assert plateau(healthy, 0.1, 5) == 20
    assert plateau([0.5 * i for i in range(15)] + [12.0] * 10, 0.1, 5) == 15   # testTrailingPlateauDetected
    assert plateau([0.5 * i for i in range(10)] + [6.0] * 3, 0.1, 5) == 13     # testPlateauBelowMinSizeIgnored
    assert plateau([0.5 * i for i in range(10)] + [6.0 + 0.05 * i for i in range(8)], 0.1, 5) == 10  # testToleranceAcceptsTinyDrift


def test_timestamp_correction_c_matches_oracle_on_random_input():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(0, 60))
        mode = int(rng.integers(0, 3))
        if mode == 0:
            raw = rng.integers(0, 5000, size=n)
        elif mode == 1:   # mostly increasing with outliers: what the aligner produces
            raw = np.sort(rng.integers(0, 5000, size=n))
            k = rng.integers(0, n + 1, size=max(1, n // 5))
            raw[k[k < n]] = rng.integers(0, 5000, size=int((k < n).sum()))
        else:             # a good prefix, then garbage (the plateau case of alignLong)
            raw = np.concatenate([np.sort(rng.integers(0, 3000, size=n // 2)), rng.integers(0, 50, size=n - n // 2)])
        raw = [int(v) for v in raw]
        c, o = q3asr.enforce_monotonicity(raw), ots.enforce_monotonicity(raw)
        assert c == o, (raw, c, o)
        assert all(b >= a for a, b in zip(c, c[1:]))
        assert q3asr.lis_positions(raw) == ots.lis_positions(raw)
        t = (np.array(c, dtype=np.float32) * np.float32(0.08)).tolist()
        assert q3asr.trailing_plateau_start(t, 0.1, 5) == ots.trailing_plateau_start(t, 0.1, 5)


def test_offset_words_reference_cases():
    """testOffsetWordsByZeroNoOp / testOffsetWordsAddsConstant (ForcedAlignerTests.swift:494-511)"""
    words = [dict(text="w", start_time=1.0, end_time=1.5), dict(text="w", start_time=2.0, end_time=2.5)]
    assert q3asr.Qwen3ASRModel.offset_words(words, 0) is words
    moved = q3asr.Qwen3ASRModel.offset_words(words, 10)
    assert [(w["start_time"], w["end_time"]) for w in moved] == [(11.0, 11.5), (12.0, 12.5)]
    assert words[0]["start_time"] == 1.0  # the input is not modified


class _ScriptedAligner(q3asr.Qwen3ASRModel):
    """align_long's loop without a GPU: align_text is scripted — every word gets `step` seconds, and words whose start would pass
    `reliable_s` collapse onto the last reliable timestamp (what the monotonicity pass makes of a saturated head)."""

    def __init__(self, reliable_s, step=1.0):
        self.reliable_s, self.step, self.calls = reliable_s, step, []

    def align_text(self, audio, text, language="English", sample_rate=16000):
        words = [w for w in text.split(" ") if w]
        self.calls.append((len(audio), len(words)))
        out, t = [], 0.0
        for w in words:
            if t < self.reliable_s:
                out.append(dict(text=w, start_time=t, end_time=t + self.step))
                t += self.step
            else:
                out.append(dict(text=w, start_time=t, end_time=t))
        return out


def test_align_long_loop():
    """Qwen3ForcedAligner.alignLong (ForcedAligner.swift:100-181) around a scripted align."""
    sr = 100  # samples per second: keeps the arrays small, the loop only looks at counts
    text = " ".join(f"w{i}" for i in range(600))
    # short audio: one pass, no plateau detection even though the tail is flat
    a = _ScriptedAligner(reliable_s=50.0)
    got = a.align_long(np.zeros(200 * sr, np.float32), text, sample_rate=sr)
    assert a.calls == [(200 * sr, 600)] and len(got) == 600 and got[-1]["start_time"] == 50.0
    # long and healthy: one pass
    a = _ScriptedAligner(reliable_s=1e9)
    got = a.align_long(np.zeros(600 * sr, np.float32), text, sample_rate=sr)
    assert a.calls == [(600 * sr, 600)] and [w["start_time"] for w in got] == [float(i) for i in range(600)]
    # long, saturating after 250 s: the first 250 words are kept, the rest re-aligned on the audio from 250 s on, then once more
    a = _ScriptedAligner(reliable_s=250.0)
    msgs = []
    got = a.align_long(np.zeros(600 * sr, np.float32), text, sample_rate=sr, progress=msgs.append)
    assert a.calls == [(600 * sr, 600), (350 * sr, 350), (100 * sr, 100)]
    assert [w["text"] for w in got] == [f"w{i}" for i in range(600)]
    assert [w["start_time"] for w in got] == [float(i) for i in range(600)]
    assert msgs == ["Audio 600.0s saturated after word 250 (250.0s); chunking remaining 350.0s (pass 2)",
                    "Audio 350.0s saturated after word 250 (250.0s); chunking remaining 100.0s (pass 3)"]
    # the remainder shorter than 5 s is dropped with its words (:158)
    a = _ScriptedAligner(reliable_s=250.0)
    got = a.align_long(np.zeros(253 * sr, np.float32), text, sample_rate=sr)
    assert len(a.calls) == 1 and len(got) == 250
    # fewer than 10 words: never split
    a = _ScriptedAligner(reliable_s=2.0)
    got = a.align_long(np.zeros(300 * sr, np.float32), "a b c d e f g h i", sample_rate=sr)
    assert len(a.calls) == 1 and len(got) == 9
    # nothing aligned / empty inputs
    assert _ScriptedAligner(1.0).align_long(np.zeros(0, np.float32), text, sample_rate=sr) == []
    assert _ScriptedAligner(1.0).align_long(np.zeros(10, np.float32), "", sample_rate=sr) == []
    # at most 10 passes
    a = _ScriptedAligner(reliable_s=241.0)
    a.align_long(np.zeros(5000 * sr, np.float32), " ".join(["w"] * 6000), sample_rate=sr)
    assert len(a.calls) == 10


# ---- GPU: the classification pass ----
@pytest.fixture(scope="module")
def aligner(built_lib):
    m = built_lib.Qwen3ASRModel.random_init("tiny-aligner", seed=20260418)
    yield m
    m.close()


@pytest.fixture(scope="module")
def aligner_oracle():
    from oracle import model, weights
    cfg = weights.preset("tiny-aligner")
    return model.Oracle(cfg, weights.random_state_dict(cfg, 20260418))


def _slotted(rng, n_words, ts):
    ids, pos = [], []
    for _ in range(n_words):
        pos.append(len(ids)); ids.append(ts)
        ids += rng.integers(3, 1900, size=int(rng.integers(1, 4))).tolist()
        pos.append(len(ids)); ids.append(ts)
    return ids, pos


@pytest.mark.gpu
def test_gpu_align_indices_match_oracle(aligner, aligner_oracle):
    rng = np.random.default_rng(3)
    clips = [synth.clip(i, n) for i, n in enumerate([16000 * 3 + 500, 16000, 16000 * 2])]
    sl = [_slotted(rng, k, 2007) for k in (7, 3, 12)]
    got = aligner.align_indices(clips, [s[0] for s in sl], [s[1] for s in sl])
    checked = 0
    for x, (ids, pos), g in zip(clips, sl, got):
        raw, margins, tops = aligner_oracle.align_indices(aligner_oracle.encode(omel.mel(x)), ids, pos)
        assert g.shape == raw.shape and (g >= 0).all() and (g < 70).all()
        ulp2 = np.maximum(np.abs(tops), 2.0 ** -6) * 2.0 ** -6   # two bf16 ulps of the winning logit: closer calls are not defined by the contract
        safe = margins > ulp2
        assert np.array_equal(g[safe], raw[safe]), (g.tolist(), raw.tolist(), margins.tolist())
        checked += int(safe.sum())
    assert checked >= 30, checked
    # a clip alone gives the same classes as inside the batch; other sample rates are converted on the device
    solo = aligner.align_indices(clips[1:2], [sl[1][0]], [sl[1][1]])[0]
    assert solo.tolist() == got[1].tolist()


@pytest.mark.gpu
def test_gpu_align_words_and_errors(aligner, tiny_model):
    rng = np.random.default_rng(4)
    words = [rng.integers(3, 1900, size=2).tolist() for _ in range(6)]
    out = aligner.align(synth.clip(2, 16000 * 4), words, words=[f"w{i}" for i in range(6)])
    assert [w["text"] for w in out] == [f"w{i}" for i in range(6)]
    starts = [w["start_time"] for w in out]
    assert all(b >= a for a, b in zip(starts, starts[1:])) and all(w["end_time"] >= w["start_time"] for w in out)
    assert all(abs(w["start_time"] / 0.08 - round(w["start_time"] / 0.08)) < 1e-3 for w in out)
    with pytest.raises(q3asr.Q3Error, match="classification head"):
        tiny_model.align_indices([synth.clip(0, 16000)], [[2007, 5, 2007]], [[0, 2]])
    with pytest.raises(q3asr.Q3Error, match="outside the slotted text"):
        aligner.align_indices([synth.clip(0, 16000)], [[2007, 5, 2007]], [[0, 3]])


@pytest.mark.gpu
def test_gpu_align_text_end_to_end(aligner):
    """align(audio, text): TextPreprocessor word pairs -> slots -> classes -> LIS fix-up -> AlignedWord-like dicts."""
    class Tok:  # toy tokenizer inside the tiny vocabulary
        def encode(self, w):
            return [10 + (ord(c) % 1500) for c in w]
    assert aligner.align_text(synth.clip(1, 32000), "no tokenizer yet") == []
    aligner.tokenizer = Tok()
    try:
        out = aligner.align_text(synth.clip(1, 16000 * 3), "Hello, world! 你好 don't stop.", "English")
        assert [w["text"] for w in out] == ["Hello,", "world!", "你", "好", "don't", "stop."]
        starts = [w["start_time"] for w in out]
        assert all(b >= a for a, b in zip(starts, starts[1:])) and all(w["end_time"] >= w["start_time"] for w in out)
        assert aligner.align_text(synth.clip(1, 32000), "?!") == []
    finally:
        aligner.tokenizer = None


@pytest.mark.gpu
def test_gpu_full_size_aligner_properties(built_lib):
    """Qwen3-ForcedAligner-0.6B dimensions (24-layer 1024-wide encoder projecting to the 1024-wide decoder, 5000 classes,
    ForcedAligner.swift:57-85): a 20 s clip with 40 words — class range, batch invariance, monotone word times."""
    import time
    m = built_lib.Qwen3ASRModel.random_init("aligner", seed=7)
    try:
        rng = np.random.default_rng(8)
        words = [rng.integers(1000, 100000, size=int(rng.integers(1, 4))).tolist() for _ in range(40)]
        x = synth.clip(4, 16000 * 20)
        t0 = time.perf_counter()
        out = m.align(x, words, words=[str(i) for i in range(40)])
        dt = time.perf_counter() - t0
        assert len(out) == 40 and all(0.0 <= w["start_time"] <= w["end_time"] <= 5000 * 0.08 for w in out)
        starts = [w["start_time"] for w in out]
        assert all(b >= a for a, b in zip(starts, starts[1:]))
        ids, pos = [], []
        for w in words:
            pos.append(len(ids)); ids.append(151705); ids += w; pos.append(len(ids)); ids.append(151705)
        a = m.align_indices([x, x[:16000 * 7]], [ids, ids[:30]], [pos, [p for p in pos if p < 30]])
        b = m.align_indices([x], [ids], [pos])
        assert a[0].tolist() == b[0].tolist() and (a[0] >= 0).all() and (a[0] < 5000).all() and len(a[1]) == len([p for p in pos if p < 30])
        print(f"align of 20 s / 40 words: {dt * 1000:.1f} ms (first call, includes buffer allocation)")
    finally:
        m.close()
