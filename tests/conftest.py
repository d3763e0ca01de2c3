import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    from detex_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def ds_golden():
    return np.load(os.path.join(GOLDEN, "ds_golden.npz"))


@pytest.fixture(scope="session")
def trig_golden():
    return np.load(os.path.join(GOLDEN, "trigger_golden.npz"))


@pytest.fixture(scope="session")
def ccx_golden():
    return np.load(os.path.join(GOLDEN, "ccx_golden.npz"))


@pytest.fixture(scope="session")
def align_golden():
    return np.load(os.path.join(GOLDEN, "align_golden.npz"))


@pytest.fixture(scope="session")
def gap_golden():
    return np.load(os.path.join(GOLDEN, "gap_golden.npz"))
