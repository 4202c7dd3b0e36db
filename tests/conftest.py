import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) the shared library; CPU-only tests that inspect the ABI need it too."""
    from gmlm_b200 import build as _b
    return _b.build()


@pytest.fixture(scope="session")
def cuda_dev(lib_built):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(1, max|b|)-style relative error used for the 1e-5 / 2e-2 gates:
    norm-wise relative error, the measure that is meaningful for sums with cancellation."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().clamp(min=1e-30)
    return float((a - b).abs().max() / denom)
