"""CPU restatement (TEST INFRASTRUCTURE ONLY) of Qwen3ASRModel.pickNextToken,
/root/reference/Sources/Qwen3ASR/Qwen3ASR.swift:449-520: repetition penalty (sign-aware), no-repeat-n-gram mask, temperature via the
Gumbel-max trick, first maximum.  Pinned by the reference's own unit tests on toy logits
(Tests/Qwen3ASRTests/Qwen3DecodingOptionsTests.swift:47-235, ported in tests/test_sampler.py).

The reference draws u ~ U[1e-6, 1] from the system RNG; the library replaces that with a counter-based stream
(csrc/ops.cu sample_kernel): u_i = 1e-6 + r_i / 2^24 * (1 - 1e-6), r_i = top 24 bits of splitmix64(key + i),
key = splitmix64(seed ^ (step << 32) ^ seq).  ``gumbel_uniforms`` restates it so that temperature runs can be compared exactly."""
import numpy as np

_M = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M
    return x ^ (x >> 31)


def gumbel_uniforms(vocab, seed=0, step=0, seq=0):
    key = _splitmix64((seed ^ ((step << 32) & _M) ^ seq) & _M)
    r = np.array([_splitmix64((key + i) & _M) >> 40 for i in range(vocab)], dtype=np.float32)
    return (np.float32(1e-6) + r * np.float32(1.0 / 16777216.0) * np.float32(1.0 - 1e-6)).astype(np.float32)


def adjusted_scores(logits, generated, repetition_penalty=1.0, no_repeat_ngram_size=0, temperature=0.0, u=None):
    scores = np.array(logits, dtype=np.float32).reshape(-1).copy()
    vocab = scores.size
    gen = [int(t) for t in generated]
    if repetition_penalty > 1.0 and gen:                         # :473-485
        for t in set(gen):
            if 0 <= t < vocab:
                v = scores[t]
                scores[t] = v / np.float32(repetition_penalty) if v > 0 else v * np.float32(repetition_penalty)
    n = no_repeat_ngram_size                                      # :489-505
    if n > 0 and len(gen) >= n - 1:
        last = gen[len(gen) - (n - 1):] if n > 1 else []
        if len(gen) >= n:
            for i in range(0, len(gen) - n + 1):
                if gen[i:i + n - 1] == last:
                    f = gen[i + n - 1]
                    if 0 <= f < vocab:
                        scores[f] = -np.inf
    if temperature > 0:                                           # :509-515
        assert u is not None and len(u) == vocab
        with np.errstate(divide="ignore"):
            scores = scores / np.float32(temperature) - np.log(-np.log(np.asarray(u, dtype=np.float32)))
    return scores.astype(np.float32)


def pick_next_token(logits, generated, repetition_penalty=1.0, no_repeat_ngram_size=0, temperature=0.0, u=None):
    scores = adjusted_scores(logits, generated, repetition_penalty, no_repeat_ngram_size, temperature, u)
    best, idx = -np.inf, 0                                        # :518-523: strict >, so the first maximum; index 0 if all -inf
    for i, v in enumerate(scores):
        if v > best:
            best, idx = v, i
    return idx
