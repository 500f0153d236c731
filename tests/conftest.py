import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _cuda_ok()


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built in-tree (never from a cache outside the repo)."""
    import q3asr
    if not os.path.exists(q3asr.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return q3asr


@pytest.fixture(scope="session")
def tiny_model(built_lib):
    m = built_lib.Qwen3ASRModel.random_init("tiny", seed=20260418)
    yield m
    m.close()


@pytest.fixture(scope="session")
def tiny_oracle():
    from oracle import model, weights
    cfg = weights.preset("tiny")
    return model.Oracle(cfg, weights.random_state_dict(cfg, 20260418))
