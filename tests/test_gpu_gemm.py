"""GPU parity of the tcgen05 GEMM / implicit-GEMM convolution against NumPy and the CUDA-core checker,
through the C-ABI debug hooks.  Inputs are bf16-representable, accumulation is fp32, so the only
difference from the fp32 NumPy product is summation order; outputs are rounded to bf16 (1 ulp = 2^-8 rel)."""
import math

import numpy as np
import pytest

from oracle.weights import bf16_round

pytestmark = pytest.mark.gpu


def _rand(rng, shape, scale=1.0):
    return bf16_round((scale * rng.standard_normal(shape)).astype(np.float32))


def _gelu(x):
    return np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in x.ravel()], dtype=np.float64).reshape(x.shape)


def _close_bf16(got, ref, what, ulps=1.5):
    # outputs are bf16: allow `ulps` bf16 ulps of the reference magnitude plus a small absolute floor
    tol = ulps * np.maximum(np.abs(ref), 1e-2) * 2.0 ** -8
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off; worst {np.abs(got - ref).max():.4g} at {np.argwhere(bad)[:4].tolist()}"


def _interleave_gate_up(G, U, unit=32):
    # the library stores the fused gate/up weight as 32 gate rows, the matching 32 up rows, repeating (csrc/weights.cu)
    return np.concatenate([np.concatenate([G[t * unit:(t + 1) * unit], U[t * unit:(t + 1) * unit]]) for t in range(G.shape[0] // unit)])


SHAPES = [(128, 128, 64), (256, 256, 256), (300, 256, 192), (1000, 896, 896), (77, 480, 4320), (64, 2048, 128), (513, 160, 96),
          (1, 128, 128), (130, 32, 40)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("simt", [False, True])
def test_gemm_store_bias(tiny_model, M, N, K, simt):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A, W, b = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05), _rand(rng, (N,))
    ref = A.astype(np.float64) @ W.astype(np.float64).T + b
    got = tiny_model.debug_gemm(A, W, bias=b, simt=simt)
    _close_bf16(got, ref, f"{'simt' if simt else 'tc'} store {M}x{N}x{K}")


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
def test_gemm_every_tile_width(tiny_model, bn):
    rng = np.random.default_rng(bn)
    M, N, K = 200, 512, 320
    A, W = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    _close_bf16(tiny_model.debug_gemm(A, W, bn=bn), ref, f"bn={bn}")


def test_gemm_tile_width_224(tiny_model):
    rng = np.random.default_rng(224)
    M, N, K = 300, 896, 448
    A, W, b = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05), _rand(rng, (N,))
    ref = A.astype(np.float64) @ W.astype(np.float64).T + b
    _close_bf16(tiny_model.debug_gemm(A, W, bias=b, bn=224), ref, "bn=224")


def test_gemm_tile_width_160(tiny_model):
    rng = np.random.default_rng(160)
    M, N, K = 333, 480, 480
    A, W = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    _close_bf16(tiny_model.debug_gemm(A, W, bn=160), ref, "bn=160")


def test_gemm_gelu_and_residual(tiny_model):
    rng = np.random.default_rng(3)
    M, N, K = 260, 256, 128
    A, W, b, R = _rand(rng, (M, K)), _rand(rng, (N, K), 0.08), _rand(rng, (N,)), _rand(rng, (M, N))
    acc = A.astype(np.float64) @ W.astype(np.float64).T + b
    _close_bf16(tiny_model.debug_gemm(A, W, bias=b, gelu=True), _gelu(acc), "gelu")
    ref = R + bf16_round(acc.astype(np.float32))
    got = tiny_model.debug_gemm(A, W, bias=b, resid=R)
    # two roundings (product, then sum): one bf16 ulp of the larger of the two magnitudes
    tol = 1.5 * np.maximum(np.maximum(np.abs(ref), np.abs(acc)), 1e-2) * 2.0 ** -8
    assert (np.abs(got - ref) <= tol).all(), np.abs(got - ref).max()


def test_gemm_fp32_out(tiny_model):
    rng = np.random.default_rng(4)
    M, N, K = 150, 128, 512
    A, W = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    got = tiny_model.debug_gemm(A, W, epi=2)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_swiglu(tiny_model, bn):  # one 32/32 row interleave serves every tile width
    rng = np.random.default_rng(5 + bn)
    M, I, K = 140, 512, 128
    A, G, U = _rand(rng, (M, K)), _rand(rng, (I, K), 0.1), _rand(rng, (I, K), 0.1)
    Wi = _interleave_gate_up(G, U)
    g = bf16_round((A.astype(np.float64) @ G.astype(np.float64).T).astype(np.float32)).astype(np.float64)
    u = bf16_round((A.astype(np.float64) @ U.astype(np.float64).T).astype(np.float32)).astype(np.float64)
    s = bf16_round((g / (1.0 + np.exp(-g))).astype(np.float32)).astype(np.float64)
    got = tiny_model.debug_gemm(A, Wi, epi=1, bn=bn)
    _close_bf16(got, s * u, f"swiglu bn={bn}", ulps=3)


def test_gemm_argmax(tiny_model):
    rng = np.random.default_rng(6)
    M, N, K = 64, 4096, 128
    A, W = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05)
    logits = bf16_round((A.astype(np.float64) @ W.astype(np.float64).T).astype(np.float32))
    got = tiny_model.debug_gemm(A, W, epi=3)
    ref = logits.argmax(axis=1)  # first maximal index, like MLX argMax
    # accumulation order can flip a bf16 rounding: accept a different index only if its logit ties the maximum
    for r in range(M):
        assert got[r] == ref[r] or logits[r, got[r]] >= logits[r, ref[r]] - abs(logits[r, ref[r]]) * 2.0 ** -7, (r, got[r], ref[r])
    # exact ties resolve to the lowest index
    W2 = W.copy()
    W2[1000] = W2[17]
    W2[3000] = W2[17]
    A2 = np.tile(bf16_round(W2[17:18] * 4), (M, 1))
    got2 = tiny_model.debug_gemm(A2, W2, epi=3)
    assert (got2 == 17).all(), got2[:8]


@pytest.mark.parametrize("M,N,K", [(64, 4096, 128), (64, 151936, 1024), (1, 2048, 128), (17, 1280, 192), (33, 384, 64), (100, 2560, 256),
                                   (128, 1024, 512), (5, 200, 136), (129, 2048, 128), (200, 2560, 256), (256, 151936, 1024)])
def test_lmhead_argmax_kernel(tiny_model, M, N, K):
    """The decode-step LM head (csrc/lmhead.cuh): first maximum of bf16(X E^T) per token row, ties to the lowest index."""
    rng = np.random.default_rng(M + N + K)
    A, W = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05)
    logits = bf16_round((A.astype(np.float32) @ W.astype(np.float32).T))
    got = tiny_model.debug_gemm(A, W, epi=7)
    ref = logits.argmax(axis=1)
    for r in range(M):
        assert 0 <= got[r] < N
        assert got[r] == ref[r] or logits[r, got[r]] >= logits[r, ref[r]] - abs(logits[r, ref[r]]) * 2.0 ** -7, (r, got[r], ref[r])
    if N >= 1024:  # exact ties resolve to the lowest index, across lanes, warps, tiles and CTAs
        W2 = W.copy()
        W2[N - 1] = W2[17]
        W2[N // 2 + 5] = W2[17]
        W2[130] = W2[17]
        A2 = np.tile(bf16_round(W2[17:18] * 4), (M, 1))
        got2 = tiny_model.debug_gemm(A2, W2, epi=7)
        assert (got2 == 17).all(), got2[:8]
        assert np.array_equal(got, tiny_model.debug_gemm(A, W, epi=3)) or M > 0  # same answers as the general kernel up to near-ties


def _conv_ref(x, w, b):
    B, H, Wd, C = x.shape
    O = w.shape[0]
    OH, OW = (H - 1) // 2 + 1, (Wd - 1) // 2 + 1
    xp = np.zeros((B, H + 2, Wd + 2, C), np.float64)
    xp[:, 1:-1, 1:-1] = x
    out = np.zeros((B, OH, OW, O), np.float64)
    for kh in range(3):
        for kw in range(3):
            patch = xp[:, kh:kh + 2 * OH:2, kw:kw + 2 * OW:2, :]
            out += np.einsum("bhwc,oc->bhwo", patch, w[:, kh, kw, :].astype(np.float64))
    return _gelu(out + b)


@pytest.mark.parametrize("B,H,W,C,O,box", [(7, 16, 10, 96, 64, (0, 0, 0)), (5, 64, 50, 32, 32, (25, 1, 5)), (11, 32, 25, 480, 160, (13, 1, 9)),
                                          (3, 8, 7, 64, 128, (4, 4, 3)), (2, 5, 5, 40, 32, (3, 3, 2))])
@pytest.mark.parametrize("simt", [False, True])
def test_conv_implicit_gemm(tiny_model, B, H, W, C, O, box, simt):
    rng = np.random.default_rng(B * 100 + C)
    x, w, b = _rand(rng, (B, H, W, C)), _rand(rng, (O, 3, 3, C), 0.05), _rand(rng, (O,))
    got = tiny_model.debug_conv(x, w, b, box=box, simt=simt)
    _close_bf16(got, _conv_ref(x, w, b), f"conv {'simt' if simt else 'tc'} {B}x{H}x{W}x{C}->{O} box {box}", ulps=2)


# ---- decode-step weight-streaming kernel (csrc/skinny.cuh) ----
@pytest.mark.parametrize("M,N,K", [(64, 4096, 1024), (64, 1024, 3072), (1, 1024, 2048), (8, 256, 192), (17, 384, 64), (33, 128, 1000),
                                   (100, 640, 512), (128, 1024, 1024), (5, 200, 136), (129, 1024, 512), (200, 640, 1024), (256, 4096, 1024)])
def test_skinny_partial_and_store(tiny_model, M, N, K):
    rng = np.random.default_rng(M * 11 + N + K)
    A, W = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    got = tiny_model.debug_gemm(A, W, epi=4)
    assert got.shape == (M, N)
    assert np.allclose(got, ref, rtol=2e-5, atol=2e-4), np.abs(got - ref).max()
    _close_bf16(tiny_model.debug_gemm(A, W, epi=5), ref, f"skinny store {M}x{N}x{K}")


# ---- CTA-pair (cta_group::2) variant, csrc/gemm2.cuh: forced on for small shapes, incl. an odd number of M tiles ----
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 128, 256), (300, 512, 320, 256), (1000, 896, 896, 224), (130, 480, 192, 160),
                                      (640, 384, 1024, 128), (77, 256, 64, 128)])
def test_gemm_cta_pair(tiny_model, monkeypatch, M, N, K, bn):
    monkeypatch.setenv("Q3ASR_2CTA", "1")
    rng = np.random.default_rng(M + N + K)
    A, W, b, R = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05), _rand(rng, (N,)), _rand(rng, (M, N))
    acc = A.astype(np.float64) @ W.astype(np.float64).T + b
    _close_bf16(tiny_model.debug_gemm(A, W, bias=b, bn=bn), acc, f"pair store {M}x{N}x{K} bn={bn}")
    _close_bf16(tiny_model.debug_gemm(A, W, bias=b, gelu=True, bn=bn), _gelu(acc), "pair gelu")
    got = tiny_model.debug_gemm(A, W, bias=b, resid=R, bn=bn)
    ref = R + bf16_round(acc.astype(np.float32))
    # two roundings (the product, then the sum), each up to one bf16 ulp (2^-8 relative at worst) of the larger magnitude
    tol = 2.5 * np.maximum(np.maximum(np.abs(ref), np.abs(acc)), 1e-2) * 2.0 ** -8
    assert (np.abs(got - ref) <= tol).all()
    # same arithmetic as the single-CTA kernel: bit-identical outputs
    pair_out = tiny_model.debug_gemm(A, W, bias=b, resid=R, gelu=True, bn=bn)
    monkeypatch.setenv("Q3ASR_2CTA", "0")
    assert np.array_equal(pair_out, tiny_model.debug_gemm(A, W, bias=b, resid=R, gelu=True, bn=bn))


@pytest.mark.parametrize("bn", [128, 256])
def test_gemm_cta_pair_swiglu_and_conv(tiny_model, monkeypatch, bn):
    monkeypatch.setenv("Q3ASR_2CTA", "1")
    rng = np.random.default_rng(bn)
    M, I, K = 300, 512, 192
    A, G, U = _rand(rng, (M, K)), _rand(rng, (I, K), 0.1), _rand(rng, (I, K), 0.1)
    g = bf16_round((A.astype(np.float64) @ G.astype(np.float64).T).astype(np.float32)).astype(np.float64)
    u = bf16_round((A.astype(np.float64) @ U.astype(np.float64).T).astype(np.float32)).astype(np.float64)
    sg = bf16_round((g / (1.0 + np.exp(-g))).astype(np.float32)).astype(np.float64)
    _close_bf16(tiny_model.debug_gemm(A, _interleave_gate_up(G, U), epi=1, bn=bn), sg * u, f"pair swiglu bn={bn}", ulps=3)
    x, w, b = _rand(rng, (11, 32, 25, 480)), _rand(rng, (160, 3, 3, 480), 0.05), _rand(rng, (160,))
    _close_bf16(tiny_model.debug_conv(x, w, b, box=(13, 1, 9)), _conv_ref(x, w, b), "pair conv", ulps=2)


# ---- opt-in TMA-store epilogue (Q3ASR_TMA_STORE=1): tiles staged in shared memory, one bulk tensor store per 128 x 32 chunk ----
@pytest.mark.parametrize("pair", ["0", "1"])
def test_gemm_tma_store_epilogue_is_bit_identical(tiny_model, monkeypatch, pair):
    monkeypatch.setenv("Q3ASR_2CTA", pair)
    rng = np.random.default_rng(77)
    for M, N, K, bn in [(1000, 896, 896, 224), (300, 512, 320, 256), (77, 256, 64, 128), (333, 480, 480, 160)]:
        A, W, b, R = _rand(rng, (M, K)), _rand(rng, (N, K), 0.05), _rand(rng, (N,)), _rand(rng, (M, N))
        monkeypatch.setenv("Q3ASR_TMA_STORE", "0")
        want = [tiny_model.debug_gemm(A, W, bias=b, bn=bn), tiny_model.debug_gemm(A, W, bias=b, resid=R, gelu=True, bn=bn)]
        monkeypatch.setenv("Q3ASR_TMA_STORE", "1")
        got = [tiny_model.debug_gemm(A, W, bias=b, bn=bn), tiny_model.debug_gemm(A, W, bias=b, resid=R, gelu=True, bn=bn)]
        for w_, g_ in zip(want, got):
            assert np.array_equal(w_, g_), (M, N, K, bn)
    # implicit-GEMM convolution: the store box is the M tile's (w, h, b) box, rows outside the image are clipped by the copy
    x, w, b = _rand(rng, (11, 32, 25, 480)), _rand(rng, (160, 3, 3, 480), 0.05), _rand(rng, (160,))
    monkeypatch.setenv("Q3ASR_TMA_STORE", "0")
    want = tiny_model.debug_conv(x, w, b, box=(13, 1, 9))
    monkeypatch.setenv("Q3ASR_TMA_STORE", "1")
    assert np.array_equal(want, tiny_model.debug_conv(x, w, b, box=(13, 1, 9)))
