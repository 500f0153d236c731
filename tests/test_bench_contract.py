"""bench.py's command-line contract on the host: the reference arm (`--impl reference`, the CPU restatement timed on the host cores)
prints exactly ONE JSON line on stdout with the keys the driver reads, rank > 0 prints nothing, and the product arm refuses to run
without a GPU instead of falling back.  A short clip and a handful of tokens keep this to seconds; the real sizes run on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

try:
    import torch
    HAS_GPU = torch.cuda.is_available()
except Exception:  # pragma: no cover
    HAS_GPU = False


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--clip-seconds", "2", "--max-tokens", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"].startswith("RTFx") and d["unit"] == "audio-seconds/second"
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - 2.0 / (d["ms_per_step"] / 1000.0)) < 1e-6 * d["value"] + 1e-9
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_rank0_only():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--clip-seconds", "2", "--max-tokens", "4"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_product_arm_has_no_cpu_fallback():
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in r.stderr
