import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: full-size sweeps (still part of -m gpu; tens of seconds each)")


@pytest.fixture(scope="session")
def models():
    """Oracle (predictor, segpp) with deterministic synthetic weights (seed 0)."""
    import torch
    from oracle.model import build_models
    torch.manual_seed(0)
    return build_models(0)


@pytest.fixture(scope="session")
def golden_nms():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "nms_golden.pt"), weights_only=False)
