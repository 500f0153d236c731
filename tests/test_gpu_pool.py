"""The utterance-batching scheduler (q3asr_pool_*, csrc/pool.cu) and the 1.7B configuration on a real GPU.  The pool is built
over the devices that exist: with one GPU it runs two workers on device 0, which exercises the same threads, dealing and
host-side gather as an 8-GPU box (the data path has no collective)."""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return tuple(range(n)) if n >= 2 else (0, 0)


def test_pool_matches_single_handle_and_keeps_order(built_lib):
    lens = [48000, 16000, 33333, 5000, 48000, 1600, 80000, 24000, 9000]
    clips = [synth.clip(i, n) for i, n in enumerate(lens)]
    single = built_lib.Qwen3ASRModel.random_init("tiny", seed=20260418)
    pool = built_lib.Pool("tiny", devices=_devices(), seed=20260418)
    try:
        want = [single.transcribe_ids([c], max_tokens=12, stop_on_eos=False)[0].tolist() for c in clips]
        got = pool.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, max_batch_per_gpu=3)  # several sub-batches per worker
        assert [g.tolist() for g in got] == want
        again = pool.transcribe_ids(clips[::-1], max_tokens=12, stop_on_eos=False)
        assert [g.tolist() for g in again[::-1]] == want
        with pytest.raises(built_lib.Q3Error):
            pool.transcribe_ids([np.zeros(10, np.float32)], max_tokens=4)  # shorter than one mel frame
    finally:
        pool.close()
        single.close()


def test_pool_with_sample_rates_and_decoder_knobs(built_lib):
    """q3asr_pool_transcribe_ids_opts: mixed sample rates are converted on each worker's device, and the deterministic decoder knobs
    (penalty + n-gram mask) give what a single handle gives, whatever the schedule."""
    rates = [16000, 24000, 8000, 24000, 16000, 48000]
    clips = [synth.clip(i, r * 2 + 100 * i) for i, r in enumerate(rates)]
    opts = built_lib.Qwen3DecodingOptions(repetition_penalty=1.3, no_repeat_ngram_size=2)
    single = built_lib.Qwen3ASRModel.random_init("tiny", seed=20260418)
    pool = built_lib.Pool("tiny", devices=_devices(), seed=20260418)
    try:
        want = [single.transcribe_ids([c], max_tokens=10, stop_on_eos=False, sample_rates=[r], options=opts)[0].tolist()
                for c, r in zip(clips, rates)]
        got = pool.transcribe_ids(clips, max_tokens=10, stop_on_eos=False, max_batch_per_gpu=2, sample_rates=rates, options=opts)
        assert [g.tolist() for g in got] == want
        plain = pool.transcribe_ids(clips, max_tokens=10, stop_on_eos=False, sample_rates=rates)
        assert [g.tolist() for g in plain] != want  # the knobs are really on
    finally:
        pool.close()
        single.close()


def test_pool_submit_wait(built_lib):
    """q3asr_pool_submit / q3asr_job_wait: two batches queued back to back give what the blocking call gives; errors surface at wait."""
    a = [synth.clip(i, 16000 + 700 * i) for i in range(5)]
    b = [synth.clip(10 + i, 24000 * 2) for i in range(3)]
    pool = built_lib.Pool("tiny", devices=_devices(), seed=20260418)
    try:
        want_a = pool.transcribe_ids(a, max_tokens=8, stop_on_eos=False)
        want_b = pool.transcribe_ids(b, max_tokens=6, stop_on_eos=False, sample_rates=[24000] * 3)
        ja = pool.submit(a, max_tokens=8, stop_on_eos=False, max_batch_per_gpu=2)
        jb = pool.submit(b, max_tokens=6, stop_on_eos=False, sample_rates=[24000] * 3)
        got_b, got_a = jb.wait(), ja.wait()
        assert [t.tolist() for t in got_a] == [t.tolist() for t in want_a]
        assert [t.tolist() for t in got_b] == [t.tolist() for t in want_b]
        bad = pool.submit([np.zeros(10, np.float32)], max_tokens=4)
        with pytest.raises(built_lib.Q3Error):
            bad.wait()
        assert pool.transcribe_ids(a[:1], max_tokens=8, stop_on_eos=False)[0].tolist() == want_a[0].tolist()  # the pool survives
    finally:
        pool.close()


def test_1p7b_configuration(built_lib, monkeypatch):
    """Qwen3-ASR-1.7B dimensions (encoder d 1024 / 24 layers, decoder hidden 2048): batch invariance, fixed length, and the two
    decode schedules (weight-streaming split-K path vs one kernel per op) agree."""
    m = built_lib.Qwen3ASRModel.random_init("1.7B", seed=3)
    try:
        clips = [synth.clip(i, 240000) for i in range(3)]  # 15 s utterances (BASELINE config 4)
        a = m.transcribe_ids(clips, max_tokens=10, stop_on_eos=False)
        assert all(len(t) == 10 for t in a)
        one = m.transcribe_ids([clips[2]], max_tokens=10, stop_on_eos=False)[0]
        assert one.tolist() == a[2].tolist()
        forced = np.arange(100, 108, dtype=np.int32)
        f_ids, f_top = m.decode_forced(clips[0], forced)
        monkeypatch.setenv("Q3ASR_NO_SKINNY", "1")
        p_ids, p_top = m.decode_forced(clips[0], forced)
        monkeypatch.delenv("Q3ASR_NO_SKINNY")
        assert np.abs(f_top - p_top).max() <= 16 * np.abs(p_top).max() * 2.0 ** -8  # two bf16 schedules of 28 layers: see tests/test_gpu_golden.py
        assert m.memory_footprint > 4e9
    finally:
        m.close()


def test_pool_on_distinct_devices_full_size(built_lib):
    """One process driving several GPUs at the 0.6B dimensions: every device needs its own > 48 KB dynamic-shared-memory opt-in of
    every kernel (cudaFuncSetAttribute is per device; a process-wide flag let only the first device launch).  Skipped on a
    one-GPU box; run with `gpurun --gpus 2` (tools/gpu_multi.sh)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    clips = [synth.clip(50 + i, 16000 * (2 + i % 3)) for i in range(3 * n)]
    single = built_lib.Qwen3ASRModel.random_init("0.6B", seed=20260418, device=n - 1)  # the LAST device first: it is not device 0
    try:
        want = [t.tolist() for t in single.transcribe_ids(clips, max_tokens=12, stop_on_eos=False)]
    finally:
        single.close()
    pool = built_lib.Pool("0.6B", devices=tuple(range(n)), seed=20260418)
    try:
        got = [t.tolist() for t in pool.transcribe_ids(clips, max_tokens=12, stop_on_eos=False, max_batch_per_gpu=2)]
        assert got == want
        got = [t.tolist() for t in pool.transcribe_ids(clips, max_tokens=12, stop_on_eos=True, max_batch_per_gpu=3)]
        assert all(g == w[:len(g)] for g, w in zip(got, want))
    finally:
        pool.close()
