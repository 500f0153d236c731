"""ctypes front to mel_oracle.c plus an independent NumPy twin (see mel_oracle.c header).

Reference: /root/reference/Sources/Qwen3ASR/AudioPreprocessing.swift:39-53, 61-164, 169-317.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MelOpts(ctypes.Structure):
    _fields_ = [("fft_size", ctypes.c_int), ("vdsp_scale2", ctypes.c_int),
                ("max_before_trim", ctypes.c_int), ("precise", ctypes.c_int)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libq3oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.q3o_mel.restype = ctypes.c_int
        _LIB.q3o_mel.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.q3o_mel_frames.restype = ctypes.c_int
        _LIB.q3o_mel_frames.argtypes = [ctypes.c_long]
        _LIB.q3o_mel_filterbank.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _LIB.q3o_hann.argtypes = [ctypes.c_void_p]
    return _LIB


def mel_frames(n):
    return min(n // 160, 120000)


def mel(x, fft_size=512, vdsp_scale2=True, max_before_trim=True, precise=False):
    """x: float32 [n] at 16 kHz -> float32 [128, n//160] (reference layout, mel-major)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    T = mel_frames(x.size)
    out = np.empty((128, T), dtype=np.float32)
    o = MelOpts(fft_size, int(vdsp_scale2), int(max_before_trim), int(precise))
    r = lib().q3o_mel(x.ctypes.data, x.size, ctypes.byref(o), out.ctypes.data)
    if r != T:
        raise RuntimeError(f"q3o_mel failed: {r}")
    return out


def filterbank(fft_size=512):
    fb = np.empty((128, fft_size // 2 + 1), dtype=np.float32)
    lib().q3o_mel_filterbank(fft_size, fb.ctypes.data)
    return fb


# ---------------------------------------------------------------------------------------
# Independent NumPy twin (same algorithm card; different code path: rfft)
# ---------------------------------------------------------------------------------------
def filterbank_numpy(fft_size=512, dtype=np.float32):
    """dtype=np.float32 is the reference's Float arithmetic; np.float64 is used only by the HF cross-check to
    separate "same algorithm" from "same rounding" (the fp32 weights differ by up to 0.5 % near triangle edges)."""
    f32 = dtype
    nb = fft_size // 2 + 1
    min_log_hz, min_log_mel = f32(1000.0), f32(15.0)
    h2m = f32(27.0) / np.log(f32(6.4))
    m2h = np.log(f32(6.4)) / f32(27.0)
    mel_max = min_log_mel + np.log(f32(8000.0) / min_log_hz) * h2m
    pts = (f32(0.0) + np.arange(130, dtype=f32) * (mel_max - f32(0.0)) / f32(129)).astype(f32)
    hz = np.where(pts < min_log_mel, f32(200.0) * pts / f32(3.0),
                  min_log_hz * np.exp((pts - min_log_mel) * m2h)).astype(f32)
    diff = (hz[1:] - hz[:-1]).astype(f32)
    freqs = (np.arange(nb, dtype=f32) * f32(16000.0) / f32(fft_size)).astype(f32)
    down = (freqs[None, :] - hz[:-2, None]) / diff[:-1, None]
    up = (hz[2:, None] - freqs[None, :]) / diff[1:, None]
    fb = np.maximum(f32(0.0), np.minimum(down, up)).astype(f32)
    return (fb * (f32(2.0) / (hz[2:] - hz[:-2]))[:, None]).astype(f32)


def mel_numpy(x, fft_size=512, vdsp_scale2=True, max_before_trim=True, fb_dtype=np.float32):
    x = np.asarray(x, dtype=np.float32)
    n = x.size
    left = x[np.clip(200 - np.arange(200), 0, n - 1)]
    right = x[np.clip(n - 2 - np.arange(200), 0, None)]
    pad = np.concatenate([left, x, right])
    nF = (pad.size - 400) // 160 + 1
    hann = (0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(400) / 400.0))).astype(np.float32)
    idx = np.arange(nF)[:, None] * 160 + np.arange(400)[None, :]
    frames = pad[idx] * hann[None, :]
    spec = np.fft.rfft(frames.astype(np.float64), n=fft_size, axis=1)
    if vdsp_scale2:
        spec = spec * 2.0
    power = (spec.real ** 2 + spec.imag ** 2).astype(np.float32)
    if fb_dtype == np.float64:
        power = spec.real ** 2 + spec.imag ** 2
    m = power.astype(np.float64) @ filterbank_numpy(fft_size, fb_dtype).astype(np.float64).T
    lm = np.log10(np.maximum(m, 1e-10)).astype(np.float32)
    g = lm.max() if max_before_trim else lm[:-1].max()
    lm = np.maximum(lm, g - np.float32(8.0)) * np.float32(0.25) + np.float32(1.0)
    T = min(nF - 1, 120000)
    return np.ascontiguousarray(lm[:T].T.astype(np.float32))
