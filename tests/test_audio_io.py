"""Front door of the batched path: WAV parser, sample-rate converter, long-form windows.
The WAV cases are the reference's own unit tests (Tests/Qwen3ASRTests/SecurityHardeningTests.swift:83-190) ported one to one:
they pin both the C parser (csrc/audio_io.cu through the C ABI; host code, runs without a GPU) and the oracle restatement."""
import struct

import numpy as np
import pytest

import q3asr
from oracle import audio_io as oio
from oracle import synth


def build_wav(sample_rate=16000, channels=1, bits=16, audio_format=1, samples=(0, 100, -100, 200, -200, 0, 0, 0),
              override_data_size=None, extra_chunk=b""):
    """SecurityHardeningTests.swift:10-71 (buildWAV)."""
    block_align = channels * (bits // 8)
    fmt = struct.pack("<HHIIHH", audio_format, channels, sample_rate, sample_rate * block_align, block_align, bits)
    payload = b"".join(struct.pack("<h", s) for s in samples)
    size = len(payload) if override_data_size is None else override_data_size
    data = b"data" + struct.pack("<I", size) + payload
    total = 4 + 8 + len(fmt) + len(extra_chunk) + len(data)
    return b"RIFF" + struct.pack("<I", total) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra_chunk + data


PARSERS = [("c_abi", q3asr.AudioFileLoader.parse_wav, q3asr.AudioLoadError), ("oracle", oio.wav_parse, oio.WavError)]


@pytest.mark.parametrize("name,parse,err", PARSERS)
def test_wav_valid_mono_and_stereo(name, parse, err):
    s, rate = parse(build_wav())                                       # testValidMonoWAV
    assert rate == 16000 and s.size == 8
    assert np.array_equal(s, np.array([0, 100, -100, 200, -200, 0, 0, 0], dtype=np.float32) / np.float32(32768.0))
    s, rate = parse(build_wav(channels=2, samples=[100, -100, 200, -200, 300, -300, 400, -400]))  # testValidStereoWAV
    assert rate == 16000 and s.size == 4
    assert abs(s[0] - 100 / 32768.0) < 1e-4 and abs(s[1] - 200 / 32768.0) < 1e-4
    s, rate = parse(build_wav(sample_rate=24000, samples=[-32768, 32767]))
    assert rate == 24000 and s[0] == -1.0 and s[1] == np.float32(32767 / 32768.0)


@pytest.mark.parametrize("name,parse,err", PARSERS)
def test_wav_rejections(name, parse, err):
    with pytest.raises(err):                                           # testTooSmallFile
        parse(bytes(20))
    bad = bytearray(build_wav()); bad[0:4] = b"NOPE"                   # testMissingRIFFHeader
    with pytest.raises(err):
        parse(bytes(bad))
    bad = bytearray(build_wav()); bad[8:12] = b"NOPE"                  # testMissingWAVEFormat
    with pytest.raises(err):
        parse(bytes(bad))
    with pytest.raises(err):                                           # testZeroChannelsRejected
        parse(build_wav(channels=0))
    with pytest.raises(err):                                           # testOversizedDataChunkSize
        parse(build_wav(override_data_size=99999))
    nodata = build_wav()                                               # testNoDataChunk
    i = nodata.index(b"data", 36)
    with pytest.raises(err):
        parse(nodata[:i] + b"xxxx" + nodata[i + 4:])
    huge = b"LIST" + struct.pack("<I", 0xFFFFFFFF) + bytes(4)          # testExtraChunkWithHugeSize
    with pytest.raises(err):
        parse(build_wav(extra_chunk=huge))
    with pytest.raises(err, match="Not PCM"):
        parse(build_wav(audio_format=3))
    with pytest.raises(err, match="Not 16-bit"):
        parse(build_wav(bits=8))


@pytest.mark.parametrize("name,parse,err", PARSERS)
def test_wav_samples_constrained_to_chunk_size(name, parse, err):     # testSamplesConstrainedToChunkSize
    s, _ = parse(build_wav(override_data_size=4))
    assert s.size == 2
    # a well-formed extra chunk before the data chunk is skipped
    s, _ = parse(build_wav(extra_chunk=b"LIST" + struct.pack("<I", 6) + b"abcdef"))
    assert s.size == 8


def test_wav_c_parser_matches_oracle_on_random_files(tmp_path):
    rng = np.random.default_rng(7)
    for ch in (1, 2, 3):
        pcm = rng.integers(-32768, 32768, size=ch * 1000).tolist()
        blob = build_wav(sample_rate=22050, channels=ch, samples=pcm)
        a, ra = q3asr.AudioFileLoader.parse_wav(blob)
        b, rb = oio.wav_parse(blob)
        assert ra == rb == 22050 and np.array_equal(a, b)
    path = tmp_path / "x.wav"
    path.write_bytes(blob)
    c, rc = q3asr.AudioFileLoader.load_wav(str(path))
    assert rc == 22050 and np.array_equal(c, b)
    with pytest.raises(OSError):
        q3asr.AudioFileLoader.load_wav(str(tmp_path / "missing.wav"))


# ---- sample-rate converter ----
@pytest.mark.parametrize("rates", [(24000, 16000), (48000, 16000), (44100, 16000), (8000, 16000), (22050, 16000), (16000, 24000)])
def test_resample_design_matches_oracle(rates):
    L, M, K, taps = q3asr.AudioFileLoader.resample_design(*rates)
    Lo, Mo, Ko, to = oio.resample_design(*rates)
    assert (L, M, K) == (Lo, Mo, Ko) and taps.shape == to.shape
    assert np.abs(taps - to).max() < 1e-7
    assert np.allclose(taps.sum(1), 1.0, atol=1e-5)  # unit DC gain in every phase


def test_resample_length_follows_the_reference():
    # AudioFileLoader.swift:190-191: AVAudioFrameCount(Double(samples.count) * ratio)
    for n, a, b in [(480000, 24000, 16000), (100001, 44100, 16000), (7, 48000, 16000), (12345, 8000, 16000), (999, 16000, 16000)]:
        assert q3asr.AudioFileLoader.resample_len(n, a, b) == oio.resample_len(n, a, b) == (n if a == b else int(n * (b / a)))


def test_oracle_resampler_against_scipy():
    """Independent check of the restatement: a tone below the new Nyquist survives with its amplitude (<= 0.1 %), a tone above it
    is rejected by >= 80 dB, and the result follows scipy.signal.resample_poly away from the edges."""
    from scipy.signal import resample_poly
    t = np.arange(48000) / 24000.0
    for f, keep in [(1000.0, True), (5000.0, True), (10000.0, False)]:
        y = oio.resample(np.sin(2 * np.pi * f * t), 24000, 16000)
        mid = y[2000:-2000]
        if keep:
            ref = np.sin(2 * np.pi * f * np.arange(y.size) / 16000.0)[2000:-2000]
            assert np.abs(mid - ref).max() < 1e-3
        else:
            assert np.abs(mid).max() < 1e-4
    x = synth.clip(3, 24000)
    x24 = np.interp(np.arange(36000) / 24000.0, np.arange(24000) / 16000.0, x).astype(np.float32)
    a = oio.resample(x24, 24000, 16000)
    b = resample_poly(x24.astype(np.float64), 2, 3, window=("kaiser", 10.0))
    assert a.size == 24000 and np.abs(a[500:-500] - b[500:-500]).max() < 2e-2


def test_longform_plan():
    assert q3asr.longform_plan(16000 * 95, 480000) == oio.longform_plan(16000 * 95, 480000) == \
        [(0, 480000), (480000, 480000), (960000, 480000), (1440000, 80000)]
    assert q3asr.longform_plan(480100, 480000) == oio.longform_plan(480100, 480000) == [(0, 480100)]  # 100-sample tail merged
    assert q3asr.longform_plan(0, 480000) == []
    rng = np.random.default_rng(0)
    for _ in range(50):
        n, w, mt = int(rng.integers(0, 10 ** 7)), int(rng.integers(1, 10 ** 6)), int(rng.integers(0, 2000))
        wins = q3asr.longform_plan(n, w, mt)
        assert wins == oio.longform_plan(n, w, mt)
        assert sum(ln for _, ln in wins) == n and all(s == sum(l for _, l in wins[:i]) for i, (s, _) in enumerate(wins))
    # BASELINE config 5: 60 minutes -> 120 windows of 30 s
    assert len(q3asr.longform_plan(3600 * 16000, 480000)) == 120
    # absurd arguments are refused instead of looping for 2^60 windows
    for n, w in [(1 << 60, 1), (1 << 40, 16), (100, 0)]:
        with pytest.raises(q3asr.Q3Error):
            q3asr.longform_plan(n, w)


# ---- GPU ----
@pytest.mark.gpu
@pytest.mark.parametrize("rates,n", [((24000, 16000), 72001), ((48000, 16000), 50000), ((44100, 16000), 44100), ((8000, 16000), 9000),
                                     ((16000, 24000), 16000), ((24000, 16000), 5)])
def test_gpu_resample_matches_oracle(tiny_model, rates, n):
    x = synth.clip(1, n)
    got = tiny_model.resample(x, *rates)
    ref = oio.resample(x, *rates)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-5


@pytest.mark.gpu
def test_gpu_transcribe_resamples_on_device(tiny_model):
    """transcribe(audio, sampleRate: 24000) == transcribe(resample(audio), 16000): the device-side conversion writes the very samples
    the separate call returns (AudioPreprocessing.swift:323-337)."""
    x24 = [synth.clip(i, 24000 * 2 + 77 * i) for i in range(3)]
    x16 = [tiny_model.resample(x, 24000, 16000) for x in x24]
    a = tiny_model.transcribe_ids(x24 + [x16[0]], max_tokens=6, stop_on_eos=False, sample_rates=[24000, 24000, 24000, 16000])
    b = tiny_model.transcribe_ids(x16 + [x16[0]], max_tokens=6, stop_on_eos=False)
    assert [t.tolist() for t in a] == [t.tolist() for t in b]
    feats = tiny_model.extract_features(x16[1])
    assert feats.shape == (128, x16[1].size // 160)
    assert tiny_model.transcribe(x24[0], sample_rate=24000, max_tokens=4) == tiny_model.transcribe(x16[0], max_tokens=4)


@pytest.mark.gpu
def test_gpu_transcribe_long_windows(tiny_model):
    x = synth.clip(5, 16000 * 7 + 100)
    segs = tiny_model.transcribe_long(x, window_seconds=2.0, max_tokens=5, batch=3)
    assert [s["segment_index"] for s in segs] == [0, 1, 2, 3] and segs[-1]["end_time"] == x.size / 16000
    wins = q3asr.longform_plan(x.size, 32000)
    solo = tiny_model.transcribe_ids([x[s:s + n] for s, n in wins], max_tokens=5)
    assert [s["ids"].tolist() for s in segs] == [t.tolist() for t in solo]


# ---- WAVWriter (Sources/AudioCommon/WAVWriter.swift:11-47): the reference's WAVWriterTests.swift ported, and the write -> load round trip ----
def test_wav_write_and_read_back(tmp_path):                                   # testWriteAndReadBack
    p = tmp_path / "a.wav"
    q3asr.AudioFileLoader.write_wav(p, [0.0, 0.5, -0.5, 1.0, -1.0], 16000)
    data = p.read_bytes()
    assert len(data) > 44 and data[:4] == b"RIFF" and data[8:12] == b"WAVE"
    x, rate = q3asr.AudioFileLoader.load_wav(p)
    assert rate == 16000 and x.tolist() == [0.0, 16383 / 32768, -16383 / 32768, 32767 / 32768, -32767 / 32768]   # Int16(x * 32767) truncates


def test_wav_write_empty_samples(tmp_path):                                   # testWriteEmptySamples
    p = tmp_path / "empty.wav"
    q3asr.AudioFileLoader.write_wav(p, [], 16000)
    assert len(p.read_bytes()) == 44


def test_wav_round_trip_preserves_length_and_values(tmp_path):                # testRoundTripPreservesLength (+ values)
    x = np.sin(np.arange(1000) * 0.1).astype(np.float32)
    p = tmp_path / "rt.wav"
    q3asr.AudioFileLoader.write_wav(p, x, 24000)
    y, rate = q3asr.AudioFileLoader.load_wav(p)
    assert rate == 24000 and y.size == 1000
    assert np.array_equal(y, np.trunc(x * np.float32(32767.0)).astype(np.float32) / np.float32(32768.0))
    assert np.abs(y - x).max() <= 2.0 / 32768
    # the oracle's parser reads the same file the same way
    yo, ro = oio.wav_parse(p.read_bytes())
    assert ro == 24000 and np.array_equal(np.asarray(yo, dtype=np.float32), y)


@pytest.mark.parametrize("rate", [8000, 16000, 22050, 24000, 44100, 48000])  # testDifferentSampleRates
def test_wav_write_sample_rates(tmp_path, rate):
    p = tmp_path / f"r{rate}.wav"
    q3asr.AudioFileLoader.write_wav(p, [0.1, 0.2, 0.3], rate)
    data = p.read_bytes()
    assert len(data) == 50 and struct.unpack("<I", data[24:28])[0] == rate and struct.unpack("<I", data[28:32])[0] == 2 * rate
    assert q3asr.AudioFileLoader.load_wav(p)[1] == rate


def test_wav_write_clamps_and_refuses_bad_arguments(tmp_path):
    p = tmp_path / "c.wav"
    q3asr.AudioFileLoader.write_wav(p, [2.0, -3.0, float("inf"), float("-inf"), 1e-9], 16000)
    assert q3asr.AudioFileLoader.load_wav(p)[0].tolist() == [32767 / 32768, -32767 / 32768, 32767 / 32768, -32767 / 32768, 0.0]
    with pytest.raises(q3asr.AudioLoadError):
        q3asr.AudioFileLoader.write_wav(p, [0.0], 0)
    with pytest.raises(q3asr.AudioLoadError):
        q3asr.AudioFileLoader.write_wav(tmp_path / "no" / "such" / "dir.wav", [0.0], 16000)
