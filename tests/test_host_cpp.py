"""The C++ host mirror of the reference's Swift API (qwen3-asr-swift_b200/host/qwen3_asr.hpp) compiles against the
C ABI, links libq3asr.so, and behaves like the reference surface (defaults, size detection, error behaviour)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "qwen3-asr-swift_b200")
EXE = os.path.join(PKG, "build", "host_demo")


@pytest.fixture(scope="module")
def host_demo(built_lib):
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, os.path.join(PKG, "host", "host_demo.cpp"),
                    "-L" + os.path.join(PKG, "lib"), "-lq3asr", "-Wl,-rpath," + os.path.join(PKG, "lib")], check=True)
    return EXE


def test_host_mirror_compiles_links_and_reports(host_demo, tmp_path):
    r = subprocess.run([host_demo, "symbols", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sm_100a" in r.stdout and "tokenizer ok" in r.stdout and "checkpoint detection ok" in r.stdout and "wav round trip ok" in r.stdout
    # without a GPU the load must fail loudly (no CPU fallback), with one it must succeed
    assert ("load error" in r.stdout) or ("created on GPU" in r.stdout)


@pytest.mark.gpu
def test_host_mirror_transcribes(host_demo):
    r = subprocess.run([host_demo, "run", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mel 128 x 300" in r.stdout


# ---- the transcribe-batch front door (host/transcribe_batch.cpp; TranscribeBatchCommand.swift:45-139) ----
BATCH_EXE = os.path.join(PKG, "build", "transcribe_batch")


@pytest.fixture(scope="module")
def transcribe_batch(built_lib):
    os.makedirs(os.path.dirname(BATCH_EXE), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", BATCH_EXE, os.path.join(PKG, "host", "transcribe_batch.cpp"),
                    "-L" + os.path.join(PKG, "lib"), "-lq3asr", "-Wl,-rpath," + os.path.join(PKG, "lib")], check=True)
    return BATCH_EXE


def _write_wav(path, x, rate, channels=1):
    pcm = np.clip(np.round(np.asarray(x) * 32767.0), -32768, 32767).astype("<i2")
    if channels > 1:
        pcm = np.repeat(pcm[:, None], channels, axis=1).reshape(-1)
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + pcm.nbytes, b"WAVE", b"fmt ", 16, 1, channels, rate, rate * 2 * channels,
                      2 * channels, 16, b"data", pcm.nbytes)
    with open(path, "wb") as f:
        f.write(hdr + pcm.tobytes())


def test_transcribe_batch_file_discovery(transcribe_batch, tmp_path):
    """TranscribeBatchCommand.swift:217-234: listed extensions only (case-insensitive), hidden entries skipped, sorted by file name."""
    t = np.arange(1600) / 16000.0
    for name in ["b.wav", "a.WAV", "sub/c.wav", ".hidden.wav", ".cache/d.wav", "notes.txt", "e.flac"]:
        path = tmp_path / name
        path.parent.mkdir(parents=True, exist_ok=True)
        _write_wav(path, 0.1 * np.sin(2 * np.pi * 440 * t), 16000)
    r = subprocess.run([transcribe_batch, str(tmp_path), "--list"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split() == ["Found", "4", "audio", "files", "a.WAV", "b.wav", "c.wav", "e.flac"]
    r = subprocess.run([transcribe_batch, str(tmp_path), "--list", "--extensions", "txt"], capture_output=True, text=True, timeout=60)
    assert r.stdout.split()[-1] == "notes.txt"
    empty = tmp_path / "empty"
    empty.mkdir()
    r = subprocess.run([transcribe_batch, str(empty)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "No audio files found" in r.stdout


@pytest.mark.gpu
def test_transcribe_batch_jsonl(transcribe_batch, tmp_path):
    from oracle import synth
    _write_wav(tmp_path / "a16.wav", synth.clip(0, 16000 * 2), 16000)
    _write_wav(tmp_path / "b24.wav", synth.clip(1, 24000 * 2), 24000)
    _write_wav(tmp_path / "c8_stereo.wav", synth.clip(2, 8000 * 3), 8000, channels=2)
    _write_wav(tmp_path / "d_long.wav", synth.clip(3, 16000 * 5), 16000)
    (tmp_path / "e_bad.wav").write_bytes(b"RIFF" + bytes(100))
    out = tmp_path / "txt"
    # two pool workers (both on device 0 when the box has one GPU), groups of 2 x 3 utterances, file loading overlapped with the GPU work
    r = subprocess.run([transcribe_batch, str(tmp_path), "--jsonl", "--max-tokens", "6", "--window-seconds", "2", "--batch", "3",
                        "--devices", "0,0", "--output-dir", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    recs = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    assert [x["file"] for x in recs] == ["a16", "b24", "c8_stereo", "d_long", "e_bad"]
    assert "error" in recs[4] and "Invalid WAV" in recs[4]["error"]
    for x, dur in zip(recs[:4], [2.0, 2.0, 3.0, 5.0]):
        assert abs(x["duration"] - dur) < 0.01 and x["time"] > 0 and x["rtf"] > 0 and x["text"]
    # 5 s at 2 s windows -> 3 windows of up to 6 ids each, joined by spaces (id-string fallback without a tokenizer)
    assert 3 <= len(recs[3]["text"].split()) <= 18 and 1 <= len(recs[0]["text"].split()) <= 6
    assert "Batch complete: 5 files, 12.0s audio" in r.stdout and "Aggregate RTF" in r.stdout
    assert sorted(os.listdir(out)) == ["a16.txt", "b24.txt", "c8_stereo.txt", "d_long.txt"]
    assert (out / "a16.txt").read_text() == recs[0]["text"]
