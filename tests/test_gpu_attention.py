"""GPU parity of the segment-packed attention kernels against a float64 NumPy softmax(QK^T)V, through the C-ABI debug
hook: the tcgen05/TMEM kernel (the product path) and the mma.sync kernel kept as its checker.  Inputs are bf16; the
kernels round P to bf16 for the PV product and the output to bf16, so the tolerance is a few bf16 ulps of the row scale."""
import numpy as np
import pytest

from oracle.weights import bf16_round

pytestmark = pytest.mark.gpu


def _ref(q, k, v, segs, heads, group, causal, scale):
    rows, hd = q.shape[0], q.shape[1] // heads
    out = np.zeros((rows, heads * hd))
    mag = np.zeros((rows, heads * hd))  # softmax-weighted mean of |v|: the scale of the rounding error of P
    for r0, n in segs:
        for h in range(heads):
            kv = h // group
            Q = q[r0:r0 + n, h * hd:(h + 1) * hd].astype(np.float64)
            K = k[r0:r0 + n, kv * hd:(kv + 1) * hd].astype(np.float64)
            V = v[r0:r0 + n, kv * hd:(kv + 1) * hd].astype(np.float64)
            s = Q @ K.T * scale
            if causal:
                s = np.where(np.arange(n)[None, :] > np.arange(n)[:, None], -np.inf, s)
            p = np.exp(s - s.max(axis=1, keepdims=True))
            out[r0:r0 + n, h * hd:(h + 1) * hd] = (p @ V) / p.sum(axis=1, keepdims=True)
            mag[r0:r0 + n, h * hd:(h + 1) * hd] = (p @ np.abs(V)) / p.sum(axis=1, keepdims=True)
    return out, mag


CASES = [
    # heads, group, hd, causal, segment lengths
    (14, 1, 64, False, [104, 104, 104, 78]),      # encoder windows of a 30 s clip
    (4, 1, 64, False, [13, 1, 104, 91, 7]),       # ragged windows
    (16, 2, 128, True, [406]),                    # one 30 s prompt: 4 query tiles, up to 4 key blocks
    (4, 2, 128, True, [211, 37, 406, 1, 129]),    # mixed prompts, block-boundary lengths
    (2, 1, 128, True, [128, 256, 257]),
    (2, 2, 64, True, [300]),
    (2, 1, 128, False, [260]),
]


@pytest.mark.parametrize("kernel", [0, 1], ids=["tcgen05", "mma_sync"])
@pytest.mark.parametrize("heads,group,hd,causal,lens", CASES)
def test_attention_matches_numpy(tiny_model, kernel, heads, group, hd, causal, lens):
    rng = np.random.default_rng(heads * 1000 + hd + sum(lens))
    rows = sum(lens)
    segs, r0 = [], 0
    for n in lens:
        segs.append((r0, n))
        r0 += n
    q = bf16_round(rng.standard_normal((rows, heads * hd)).astype(np.float32))
    k = bf16_round(rng.standard_normal((rows, heads // group * hd)).astype(np.float32))
    v = bf16_round(rng.standard_normal((rows, heads // group * hd)).astype(np.float32))
    scale = 1.0 / np.sqrt(hd)
    got = tiny_model.debug_attention(q, k, v, segs, heads, group, causal, scale, kernel=kernel)
    ref, mag = _ref(q, k, v, segs, heads, group, causal, scale)
    err = np.abs(got - ref)
    # every P entry is rounded to bf16 (2^-9 relative, worst case all in the same direction: 2^-9 * sum p|v| / sum p, twice
    # because the denominator uses the unrounded P), and so is the output (2^-9 |ref|); plus fp32 accumulation noise
    tol = 2.0 ** -8 * mag + 2.0 ** -8 * np.abs(ref) + 1e-4
    assert (err <= tol).all(), (err.max(), np.argwhere(err > tol)[:5].tolist())
    assert np.sqrt((err ** 2).sum() / (ref ** 2).sum()) <= 4e-3
