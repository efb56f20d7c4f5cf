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


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The product has no CPU fallback: make sure the C-ABI library exists before any test imports it."""
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session")
def golden_kmers():
    return np.load(os.path.join(GOLDEN, "kmers_golden.npz"))


@pytest.fixture(scope="session")
def golden_pca():
    return np.load(os.path.join(GOLDEN, "pca_golden.npz"))


@pytest.fixture(scope="session")
def golden_dbscan():
    return np.load(os.path.join(GOLDEN, "dbscan_golden.npz"))


@pytest.fixture(scope="session")
def golden_binning():
    return np.load(os.path.join(GOLDEN, "binning_golden.npz"))


@pytest.fixture(scope="session")
def small_fna():
    return os.path.join(GOLDEN, "small.fna")
