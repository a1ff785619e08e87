import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(GOLDEN / f"{name}.npz")
    return load


def kw_of(npz, name):
    return json.loads(bytes(npz[f"{name}_kw"]).decode())


def assert_close(actual, expected, rtol=1e-3, atol_rms=1e-3, what=""):
    """|a-b| <= rtol*|b| + atol with atol = atol_rms * rms(b): the tolerance BASELINE.json states for
    fp32 cost volumes and flows (rtol 1e-3), made robust to values that cancel to ~0."""
    a = np.asarray(actual, np.float64)
    b = np.asarray(expected, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.isfinite(a).all(), f"{what}: non-finite values"
    atol = atol_rms * float(np.sqrt(np.mean(b * b)) + 1e-30)
    err = np.abs(a - b) - (rtol * np.abs(b) + atol)
    if (err > 0).any():
        i = np.unravel_index(np.argmax(err), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]:.6g} want {b[i]:.6g} "
                             f"(rms {np.sqrt(np.mean(b*b)):.3g}, {int((err > 0).sum())} of {a.size} out of tol)")
