"""Decoder knobs: Qwen3ASRModel.pickNextToken.  The toy-logit cases are the reference's own unit
tests (Tests/Qwen3ASRTests/Qwen3DecodingOptionsTests.swift:47-235), run against the oracle restatement on the CPU and against the
device kernel (csrc/ops.cu sample_kernel, through q3asr_pick_next_token) on the GPU."""
import numpy as np
import pytest

import q3asr
from oracle import sampler as osamp
from oracle import mel as omel
from oracle import synth


def _oracle_pick(logits, gen, opts, draw=0):
    u = osamp.gumbel_uniforms(len(logits), seed=opts.seed, step=draw) if opts.temperature > 0 else None
    return osamp.pick_next_token(logits, gen, opts.repetition_penalty, opts.no_repeat_ngram_size, opts.temperature, u)


def _cases(pick):
    O = q3asr.Qwen3DecodingOptions
    lg = np.full(64, -1.0, np.float32); lg[42] = 5.0; lg[7] = 2.0                 # testFastPathMatchesArgMax
    assert pick(lg, [], O()) == 42
    lg = np.full(32, -5.0, np.float32); lg[5] = 4.0; lg[7] = 3.0                  # testRepetitionPenaltyDemotesRepeatedPositiveLogit
    assert pick(lg, [5], O(repetition_penalty=2.0)) == 7
    assert pick(lg, [5], O()) == 5
    lg = np.full(16, -10.0, np.float32); lg[3] = -1.0; lg[11] = -2.0              # testRepetitionPenaltyHandlesNegativeLogitSign
    assert pick(lg, [3], O(repetition_penalty=3.0)) == 11
    lg = np.full(16, -1.0, np.float32); lg[9] = 3.0                               # testRepetitionPenaltyNoOpOnFirstToken
    assert pick(lg, [], O(repetition_penalty=2.5)) == 9
    A, B, C, D = 5, 6, 7, 4
    lg = np.full(32, -10.0, np.float32); lg[A] = 5.0; lg[B] = 2.0; lg[C] = 3.0    # testNoRepeatNgramMasksRepeatedTrigram
    assert pick(lg, [A, B, A, B], O(no_repeat_ngram_size=3)) == C
    lg = np.full(16, -5.0, np.float32); lg[D] = 3.0                               # testNoRepeatNgramAllowsNovelFollowup
    assert pick(lg, [1, 2, 3], O(no_repeat_ngram_size=3)) == D
    lg = np.zeros(16, np.float32); lg[4] = 2.0                                    # testTemperatureZeroRemainsDeterministic
    assert {pick(lg, [], O()) for _ in range(6)} == {4}
    lg = np.zeros(16, np.float32)                                                 # testTemperatureSamplingProducesVariety
    assert len({pick(lg, [], O(temperature=1.0, seed=3), draw=d) for d in range(50)}) >= 3
    lg = np.zeros(16, np.float32); lg[9] = 10.0                                   # testLowTemperatureStaysMostlyAtPeak
    assert sum(pick(lg, [], O(temperature=0.1, seed=4), draw=d) == 9 for d in range(50)) > 25
    # beyond the reference's cases: ties take the lowest index; an all-masked row falls back to index 0; n = 1 forbids every
    # token already generated; out-of-range ids in the history are ignored
    assert pick(np.array([1.0, 3.0, 3.0, 2.0], np.float32), [], O(repetition_penalty=1.5)) == 1
    assert pick(np.array([1.0, 2.0], np.float32), [0, 1], O(no_repeat_ngram_size=1)) == 0
    assert pick(np.array([1.0, 5.0, 4.0, 3.0], np.float32), [1, 2], O(no_repeat_ngram_size=1)) == 3
    assert pick(np.array([1.0, 5.0, 4.0], np.float32), [7, -1, 1], O(repetition_penalty=2.0)) == 2


def test_oracle_sampler_reference_cases():
    _cases(_oracle_pick)


def test_options_defaults():
    o = q3asr.Qwen3DecodingOptions()                                              # testDecodingOptionsDefaults
    assert (o.max_tokens, o.language, o.context, o.repetition_penalty, o.no_repeat_ngram_size, o.temperature) == (448, None, None, 1.0, 0, 0.0)
    assert o.is_greedy_fast_path
    o = q3asr.Qwen3DecodingOptions(128, "en", "hello", 1.2, 3, 0.7)               # testDecodingOptionsCustomInit
    assert (o.max_tokens, o.language, o.context, o.repetition_penalty, o.no_repeat_ngram_size, o.temperature) == (128, "en", "hello", 1.2, 3, 0.7)
    assert not o.is_greedy_fast_path


@pytest.mark.gpu
def test_gpu_sampler_reference_cases(tiny_model):
    _cases(lambda lg, gen, o, draw=0: tiny_model.pick_next_token(lg, gen, o, draw))


@pytest.mark.gpu
def test_gpu_sampler_matches_oracle_on_random_rows(tiny_model):
    rng = np.random.default_rng(11)
    O = q3asr.Qwen3DecodingOptions
    for trial in range(12):
        vocab = int(rng.choice([17, 1000, 2048, 151936]))
        lg = rng.normal(0, 3, vocab).astype(np.float32)
        gen = rng.integers(0, min(vocab, 40), size=int(rng.integers(0, 60))).tolist()
        o = O(repetition_penalty=float(rng.choice([1.0, 1.3, 2.0])), no_repeat_ngram_size=int(rng.choice([0, 1, 2, 3])),
              temperature=float(rng.choice([0.0, 0.0, 0.7])), seed=trial)
        got = tiny_model.pick_next_token(lg, gen, o, draw=trial)
        u = osamp.gumbel_uniforms(vocab, seed=trial, step=trial) if o.temperature > 0 else None
        scores = osamp.adjusted_scores(lg, gen, o.repetition_penalty, o.no_repeat_ngram_size, o.temperature, u)
        if o.temperature == 0:
            assert got == osamp.pick_next_token(lg, gen, o.repetition_penalty, o.no_repeat_ngram_size)
        else:  # the device uses fast logarithms: the chosen token must be the oracle's maximum up to that error
            assert scores[got] >= scores.max() - 1e-3


@pytest.mark.gpu
def test_gpu_sampler_in_the_decode_loop(tiny_model, tiny_oracle):
    """Greedy options forced through the device sampler give the fused-argmax ids; a repetition penalty / n-gram mask changes
    the stream exactly as pickNextToken applied to the oracle's teacher-forced logits says; temperature runs are reproducible."""
    O = q3asr.Qwen3DecodingOptions
    clips = [synth.clip(i, 16000 * 2 + 333 * i) for i in range(3)]
    greedy = tiny_model.transcribe_ids(clips, max_tokens=12, stop_on_eos=False)
    forced = tiny_model.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, options=O(), force_device_sampler=True)
    assert [t.tolist() for t in forced] == [t.tolist() for t in greedy]
    o = O(repetition_penalty=1.3, no_repeat_ngram_size=2)
    got = tiny_model.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, options=o)
    for ids in got:  # the n-gram mask holds along the whole stream: no bigram occurs twice
        pairs = list(zip(ids[:-1].tolist(), ids[1:].tolist()))
        assert len(pairs) == len(set(pairs))
    assert [t.tolist() for t in got] != [t.tolist() for t in greedy]  # random-init greedy streams echo, the knobs break the echo
    t1 = tiny_model.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, options=O(temperature=0.8, seed=5))
    t2 = tiny_model.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, options=O(temperature=0.8, seed=5))
    t3 = tiny_model.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, options=O(temperature=0.8, seed=6))
    assert [t.tolist() for t in t1] == [t.tolist() for t in t2] != [t.tolist() for t in t3]
    # batch independence: a sequence alone draws the same stream as in the batch (the noise is keyed by the sequence index 0 here)
    solo = tiny_model.transcribe_ids(clips[:1], max_tokens=12, stop_on_eos=False, options=o)
    assert solo[0].tolist() == got[0].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("penalty,ngram", [(1.3, 0), (1.0, 2), (1.5, 3)])
def test_gpu_slow_path_ids_match_oracle(tiny_model, tiny_oracle, penalty, ngram):
    """generateSlow (Qwen3ASR.swift:396-447): ids bit-exact against the oracle's slow loop for a fixed decode length, up to the
    first step where the oracle's own two best adjusted scores are closer than bf16 resolution."""
    O = q3asr.Qwen3DecodingOptions
    compared = 0
    for i, n in enumerate([16000 * 2 + 555, 9000, 16000 * 3]):
        x = synth.clip(i, n)
        ref, margins, tops = tiny_oracle.generate_slow(tiny_oracle.encode(omel.mel(x)), 20, penalty, ngram, stop_on_eos=False)
        got = tiny_model.transcribe_ids([x], max_tokens=20, stop_on_eos=False, options=O(repetition_penalty=penalty, no_repeat_ngram_size=ngram))[0]
        assert len(got) == len(ref) == 20
        # a near-tie (closer than two bf16 ulps of the score) may legitimately resolve either way, and then the streams part
        ulp2 = np.maximum(np.abs(tops), 2.0 ** -6) * 2.0 ** -6
        safe = np.cumprod(margins > ulp2).astype(bool)
        assert got[safe].tolist() == ref[safe].tolist(), (got.tolist(), ref.tolist(), margins.tolist())
        compared += int(safe.sum())
    assert compared >= 12, compared
