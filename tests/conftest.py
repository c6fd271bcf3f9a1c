import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
TOY = os.path.join(GOLDEN_DIR, "toy")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def unhex(v):
    if v is None:
        return None
    if v == "nan":
        return float("nan")
    return float.fromhex(v)


def unhex_arr(a, shape=None):
    out = np.array([unhex(v) for v in a], dtype=np.float64)
    return out.reshape(shape) if shape is not None else out


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def toy():
    return lambda name: os.path.join(TOY, name)


def same_float(a, b):
    """Bit-level agreement with NaN == NaN."""
    a, b = float(a), float(b)
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    return a == b


def close(a, b, rtol):
    a, b = float(a), float(b)
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    if np.isinf(a) or np.isinf(b):
        return a == b
    return abs(a - b) <= rtol * max(abs(a), abs(b), 1e-300)
