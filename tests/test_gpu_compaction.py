"""Finished utterances leave the decode batch (continuous-batching half of the scheduler; reference semantics:
Qwen3ASR.swift:378-379 — the token is appended, then the utterance stops at EOS).

A token the random-init model emits at different steps for different clips is made the EOS token, so the sequences of one batch
finish at ragged steps.  Checked: (1) every utterance's ids equal the prefix, up to and including its first EOS, of the ids the
same batch produces with EOS ignored (batch invariance + EOS semantics); (2) the same ids with compaction switched off
(Q3ASR_NO_COMPACT) and one utterance at a time; (3) the decode loop really shrank: the rows it processed (q3asr_decode_stats)
follow the tokens generated, not batch x longest utterance, and at least one compaction happened; (4) a second run of the same
resident batch starts from the full batch again.
"""
import collections

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu

TOKENS = 80


def _clips(n):
    return [synth.clip(700 + i, 16000 * 2 + 3200 * (i % 7)) for i in range(n)]


@pytest.mark.parametrize("mega", ["0", "1"])
def test_ragged_eos_compacts_the_decode_batch(built_lib, monkeypatch, mega):
    monkeypatch.setenv("Q3ASR_MEGA", mega)
    n = 40
    clips = _clips(n)
    m = built_lib.Qwen3ASRModel.random_init("0.6B", seed=20260418)
    try:
        free = [t.tolist() for t in m.transcribe_ids(clips, max_tokens=TOKENS, stop_on_eos=False)]
    finally:
        m.close()
    # the EOS candidate: the token whose first occurrences are most spread over the steps (so the batch thins out gradually)
    firsts = collections.defaultdict(list)
    for ids in free:
        seen = set()
        for s, t in enumerate(ids):
            if t not in seen:
                seen.add(t)
                firsts[t].append(s)
    eos = max(firsts, key=lambda t: (len(firsts[t]) >= n // 2) * (np.std(firsts[t]) + 0.01 * len(firsts[t])))
    assert len(firsts[eos]) >= n // 2, "no token is emitted by at least half of the clips: pick other clips"
    want = [ids[:ids.index(eos) + 1] if eos in ids else ids for ids in free]
    cfg = built_lib.preset("0.6B")
    cfg.tok_eos = int(eos)
    m = built_lib.Qwen3ASRModel.random_init("0.6B", seed=20260418, config=cfg)
    try:
        got = [t.tolist() for t in m.transcribe_ids(clips, max_tokens=TOKENS, stop_on_eos=True)]
        st = m.decode_stats()
        bad = [(i, next((s for s, (a, b) in enumerate(zip(g_, w_)) if a != b), min(len(g_), len(w_))), len(g_), len(w_))
               for i, (g_, w_) in enumerate(zip(got, want)) if g_ != w_]
        assert not bad, f"(utterance, first differing step, got length, want length): {bad[:10]}; stats {st}"
        total = sum(len(w) for w in want)
        print(f"eos {eos}: {sum(eos in w for w in want)}/{n} utterances stop early, {total} tokens; decode loop: {st}")
        assert st["compactions"] >= 1 and st["rows"] < n
        # rows processed: at most the generated tokens plus the slack of polling every 16 steps and compacting in quarters
        assert st["row_steps"] <= 0.8 * n * st["steps"], st
        # second run of the same resident batch: starts from all rows again, same ids
        got2 = [t.tolist() for t in m.transcribe_ids(clips, max_tokens=TOKENS, stop_on_eos=True)]
        assert got2 == want
        monkeypatch.setenv("Q3ASR_NO_COMPACT", "1")
        got3 = [t.tolist() for t in m.transcribe_ids(clips, max_tokens=TOKENS, stop_on_eos=True)]
        st3 = m.decode_stats()
        assert got3 == want and st3["compactions"] == 0 and st3["row_steps"] == n * st3["steps"]
        monkeypatch.delenv("Q3ASR_NO_COMPACT")
        for i in (0, 7, 23):
            assert m.transcribe_ids([clips[i]], max_tokens=TOKENS, stop_on_eos=True)[0].tolist() == want[i]
    finally:
        m.close()
