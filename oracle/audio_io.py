"""CPU restatement (TEST INFRASTRUCTURE ONLY — never imported by the product path) of the front door of the batched path:

* ``wav_parse``    — AudioFileLoader.loadWAV, /root/reference/Sources/AudioCommon/AudioFileLoader.swift:70-157, check by check.
  Pinned by the reference's own unit tests (Tests/Qwen3ASRTests/SecurityHardeningTests.swift:83-190, ported in tests/test_audio_io.py).
* ``resample``     — the polyphase Kaiser-windowed-sinc converter csrc/audio_io.cu states (the reference calls AVAudioConverter,
  AudioFileLoader.swift:159-213: an Apple framework whose filter is not in the reference, so only the output LENGTH,
  floor(n * out / in) at :190-191, is a parity target; "parity unpinned" for the sample values).  float64 arithmetic.
* ``longform_plan`` — fixed windows, short tail merged into the previous window.
"""
import math
import struct

import numpy as np


class WavError(ValueError):
    pass


def wav_parse(data: bytes):
    """-> (float32 samples of the first channel, sample rate).  AudioFileLoader.swift:70-157."""
    if len(data) <= 44:                                   # :74-76
        raise WavError("Invalid WAV file format")
    if data[0:4] != b"RIFF" or data[8:12] != b"WAVE":     # :79-88
        raise WavError("Invalid WAV file format")
    fmt, ch = struct.unpack_from("<HH", data, 20)         # :91-92
    rate, = struct.unpack_from("<I", data, 24)
    bits, = struct.unpack_from("<H", data, 34)
    if fmt != 1:
        raise WavError("Unsupported audio format: Not PCM format")
    if ch == 0:
        raise WavError("Invalid WAV file format")
    if bits != 16:
        raise WavError("Unsupported audio format: Not 16-bit")
    off, size = 36, None                                  # :109-127
    while off < len(data) - 8:
        cid = data[off:off + 4]
        csz, = struct.unpack_from("<I", data, off + 4)
        if cid == b"data":
            off += 8
            size = csz
            break
        nxt = off + 8 + csz
        if nxt > len(data):
            raise WavError("Invalid WAV file format")
        off = nxt
    if size is None or off > len(data) or off + size > len(data):   # :130-136
        raise WavError("Invalid WAV file format")
    frames = size // (2 * ch)
    pcm = np.frombuffer(data, dtype="<i2", count=frames * ch, offset=off)
    return (pcm[::ch].astype(np.float32) / np.float32(32768.0)), int(rate)


def resample_len(n, in_rate, out_rate):
    return n if in_rate == out_rate else int(float(n) * (float(out_rate) / float(in_rate)))


def _i0(x):
    return np.i0(x)


def resample_design(in_rate, out_rate):
    """(L, M, K, taps[L, 2K+2] float64) — see the design comment in csrc/audio_io.cu."""
    g = math.gcd(in_rate, out_rate)
    L, M = out_rate // g, in_rate // g
    fc = 0.945 * min(1.0, L / M)
    half = 24.0 / fc
    K = int(math.ceil(half))
    k = np.arange(-K, K + 2, dtype=np.float64)
    taps = np.zeros((L, k.size))
    for p in range(L):
        t = p / L - k
        r = np.clip(1.0 - (t / half) ** 2, 0.0, None)
        h = fc * np.sinc(fc * t) * _i0(10.0 * np.sqrt(r)) / _i0(10.0)
        h[np.abs(t) >= half] = 0.0
        taps[p] = h / h.sum()
    return L, M, K, taps


def resample(x, in_rate, out_rate):
    x = np.asarray(x, dtype=np.float64)
    if in_rate == out_rate or x.size == 0:
        return x.astype(np.float32)
    L, M, K, taps = resample_design(in_rate, out_rate)
    taps = taps.astype(np.float32).astype(np.float64)      # the library stores the taps as float
    n_out = resample_len(x.size, in_rate, out_rate)
    nt = 2 * K + 2
    xp = np.concatenate([np.zeros(K), x, np.zeros(nt + M)])
    j = np.arange(n_out, dtype=np.int64)
    i0 = (j * M) // L
    p = (j * M) % L
    y = np.zeros(n_out)
    for t in range(nt):
        y += taps[p, t] * xp[i0 + t]                        # xp index (i0 - K + t) + K
    return y.astype(np.float32)


def longform_plan(n, window, min_tail=160):
    out, pos = [], 0
    while pos < n:
        ln = min(window, n - pos)
        rest = n - pos - ln
        if 0 < rest < min_tail:
            ln += rest
        out.append((pos, ln))
        pos += ln
    return out
