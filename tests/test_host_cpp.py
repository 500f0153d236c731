"""The C++ host mirror of the reference's Swift API (qwen3-asr-swift_b200/host/qwen3_asr.hpp) compiles against the
C ABI, links libq3asr.so, and behaves like the reference surface (defaults, size detection, error behaviour)."""
import os
import subprocess

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
    assert "sm_100a" in r.stdout and "tokenizer ok" in r.stdout
    # without a GPU the load must fail loudly (no CPU fallback), with one it must succeed
    assert ("load error" in r.stdout) or ("created on GPU" in r.stdout)


@pytest.mark.gpu
def test_host_mirror_transcribes(host_demo):
    r = subprocess.run([host_demo, "run", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mel 128 x 300" in r.stdout
