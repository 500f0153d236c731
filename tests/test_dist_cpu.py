"""The N > 1 path on CPU: two ranks over gloo (tests/dist_worker.py).  The data path has no collective; what is covered is the
sharding, the scheduler's determinism across ranks, the host-side gather and the max-over-ranks timing reduction."""
import os
import subprocess
import sys

from conftest import ROOT


def test_two_rank_gloo_host_logic(built_lib):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "DIST_OK 2" in r.stdout
