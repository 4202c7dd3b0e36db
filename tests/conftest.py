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


def rel_err(a: torch.Tensor, b: torch.Tensor, floor: float = 0.0) -> float:
    """max |a-b| / max|b|: norm-wise relative error, the measure that is meaningful for sums with
    cancellation.  ``floor`` bounds the denominator from below for quantities whose exact value is zero
    (e.g. the gradient of an RGCN bias that feeds GraphNorm with mean_scale = 1: the mean subtraction
    cancels it, so both sides hold only rounding noise and a pure ratio is meaningless)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().clamp(min=max(floor, 1e-30))
    return float((a - b).abs().max() / denom)


def elementwise_err(a: torch.Tensor, b: torch.Tensor, atol_frac: float = 1e-2) -> float:
    """Elementwise relative error  max_i |a_i - b_i| / (|b_i| + atol),  atol = atol_frac * rms(b):
    the stricter companion of ``rel_err`` (north_star's "1e-5 relative" read per element).  The absolute
    term keeps elements that are tiny only through cancellation from dominating."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    atol = atol_frac * float(b.pow(2).mean().sqrt().clamp(min=1e-30))
    return float(((a - b).abs() / (b.abs() + atol)).max())
