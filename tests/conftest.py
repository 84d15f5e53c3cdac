import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, e.g. `pytest tests` on the CPU box."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Access to tests/golden/<name>.npz with 'case/key' addressing."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)

    def case(self, prefix):
        p = prefix + "/"
        return {k[len(p):]: self.z[k] for k in self.z.files if k.startswith(p)}

    def cases(self):
        return sorted({k.split("/")[0] for k in self.z.files if "/" in k})

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z.files


@pytest.fixture(scope="session")
def golden_scan():
    return Golden("scan")


@pytest.fixture(scope="session")
def golden_csm():
    return Golden("csm")


@pytest.fixture(scope="session")
def golden_bayes():
    return Golden("bayes")


@pytest.fixture(scope="session")
def golden_models():
    return Golden("models")


@pytest.fixture(scope="session")
def golden_select():
    return Golden("select")


def nmax_err(a, b):
    """normalised max error max|a-b| / max|b| (element-wise relative error is ill-posed at zero crossings)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
