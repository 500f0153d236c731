"""GPU parity of the log-mel kernel (K1) against the CPU oracle, through the C ABI.

Tolerance (BASELINE.json north_star: "mel features within 1e-4 relative error in fp32"):
|gpu - oracle| <= 1e-4 * max(1, |oracle|) on every element.
"""
import numpy as np
import pytest

from oracle import mel as omel
from oracle import synth

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _check(gpu, ref, what, tol=TOL):
    assert gpu.shape == ref.shape, (what, gpu.shape, ref.shape)
    err = np.abs(gpu - ref) / np.maximum(1.0, np.abs(ref))
    i = np.unravel_index(np.argmax(err), err.shape) if err.size else None
    assert err.size == 0 or err.max() <= tol, f"{what}: max rel err {err.max():.3e} at {i} gpu={gpu[i]} ref={ref[i]}"


@pytest.mark.parametrize("n", [160, 161, 400, 1599, 1600, 5119, 5120, 5121, 16000, 47999, 240000, 480000])
def test_mel_matches_oracle(tiny_model, n):
    x = synth.clip(n % 7, n)
    got = tiny_model.extract_features(x)
    _check(got, omel.mel(x), f"n={n}")


def test_mel_edge_signals(tiny_model):
    zeros = np.zeros(16000, np.float32)
    _check(tiny_model.extract_features(zeros), omel.mel(zeros), "zeros")
    imp = np.zeros(8000, np.float32)
    imp[4000] = 1.0
    # an impulse has frames that are exactly silent next to frames that are not: exercises the max-8 clamp pass
    _check(tiny_model.extract_features(imp), omel.mel(imp), "impulse")
    rng = np.random.default_rng(5)
    quiet_loud = np.concatenate([1e-6 * rng.standard_normal(8000), 0.5 * rng.standard_normal(8000)]).astype(np.float32)
    _check(tiny_model.extract_features(quiet_loud), omel.mel(quiet_loud), "quiet+loud (clamp active)")


def test_mel_pure_tone_near_floor(tiny_model):
    # a noiseless tone puts most bins 60-80 dB below the peak, where fp32 round-off of ANY FFT ordering is
    # ~1e-3 relative; the double-precision oracle is the fair reference there and the tolerance is 2e-3
    t = np.arange(32000) / 16000.0
    x = (0.4 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    _check(tiny_model.extract_features(x), omel.mel(x, precise=True), "tone", tol=2e-3)


def test_mel_ragged_batch_equals_single(tiny_model):
    lens = [480000, 1600, 47999, 16000, 240000, 161]
    clips = [synth.clip(i, n) for i, n in enumerate(lens)]
    outs = tiny_model.extract_features_batch(clips)
    for c, o in zip(clips, outs):
        single = tiny_model.extract_features(c)
        assert np.array_equal(o, single)  # batching must not change a single bit
        _check(o, omel.mel(c), f"batch n={c.size}")


@pytest.mark.parametrize("count", [33, 128, 200])
def test_mel_many_short_clips(tiny_model, count):
    # Many clips of a few tiles each: every CTA changes clip at almost every tile, so the look-ahead clip lookup of the tile loop
    # (csrc/mel.cu: candidates spread over a warp up to 128 clips, one lane's search above) is what this exercises.
    rng = np.random.default_rng(count)
    lens = [int(n) for n in rng.integers(160, 6000, size=count)]
    clips = [synth.clip(1000 + i, n) for i, n in enumerate(lens)]
    outs = tiny_model.extract_features_batch(clips)
    for i in range(0, count, 7):
        _check(outs[i], omel.mel(clips[i]), f"clip {i} of {count}, n={lens[i]}")
    for i in (0, count // 2, count - 1):
        assert np.array_equal(outs[i], tiny_model.extract_features(clips[i]))


def test_mel_repeatable_under_load(tiny_model):
    # The tile loop is software-pipelined across five independent groups per CTA that alias their scratch blocks (transpose ->
    # spectrum -> parked features): a missing barrier shows up as run-to-run differences.  (compute-sanitizer is not available on
    # the GPU pool; this is the race check that is.)
    lens = [48000, 1600, 4799, 16000, 24000, 161, 160, 5121] + [int(n) for n in np.random.default_rng(1).integers(160, 20000, size=140)]
    clips = [synth.clip(i, n) for i, n in enumerate(lens)]
    first = tiny_model.extract_features_batch(clips)
    for _ in range(6):
        again = tiny_model.extract_features_batch(clips)
        assert all(np.array_equal(a, b) for a, b in zip(first, again))
    for i in (0, 5, 6, 7, 50, 147):
        _check(first[i], omel.mel(clips[i]), f"clip {i}, n={lens[i]}")


def test_mel_properties_full_size(tiny_model):
    # BASELINE config sizes (64 x 30 s): size-independent properties instead of a slow oracle pass
    clips = [synth.clip(i, 480000) for i in range(64)]
    outs = tiny_model.extract_features_batch(clips)
    for o in outs:
        assert o.shape == (128, 3000) and np.isfinite(o).all()
        # dynamic range: nothing is below (max - 8)/4 + 1 once the dropped frame is accounted for
        assert o.min() >= o.max() - 2.0 - 1e-3
    # identical clip -> identical features regardless of its slot in the batch
    again = tiny_model.extract_features_batch([clips[5], clips[0]])
    assert np.array_equal(again[0], outs[5]) and np.array_equal(again[1], outs[0])
    # amplitude scaling by 2 shifts every unclamped log-mel by log10(4)/4
    a = tiny_model.extract_features(clips[3][:48000])
    b = tiny_model.extract_features((2.0 * clips[3][:48000]).astype(np.float32))
    assert np.allclose(b - a, np.log10(4.0) / 4.0, atol=2e-4)


def test_mel_maximum_size_frame_cap(tiny_model):
    # AudioPreprocessing.swift:304: at most 120000 frames are kept (20 min); the maximum runs over every computed frame, kept or not
    n = 160 * 120000 + 160 * 37 + 5
    x = synth.clip(9, n)
    x[-2000:] *= 40.0  # the loudest frames are beyond the cap: they still set the max-8 clamp
    x = np.clip(x, -1.0, 1.0).astype(np.float32)
    got = tiny_model.extract_features(x)
    assert got.shape == (128, 120000)
    _check(got, omel.mel(x), "frame cap")
