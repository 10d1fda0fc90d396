import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs an NVIDIA B200 (run with -m gpu)")


def _gpu_available():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked ``gpu`` are skipped on a box without CUDA, so that a plain ``pytest tests``
    stays green there; a GPU box selects them with ``-m gpu``."""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs an NVIDIA B200 (torch.cuda.is_available() is False)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """Build libqnmfit.so and the lane-emulation harness if they are stale/missing."""
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def qf():
    """The package with the synthetic Kerr tables installed."""
    import qnmfits_b200
    from qnmfits_b200 import workloads
    workloads.use_synthetic_tables()
    return qnmfits_b200


@pytest.fixture(scope="session")
def oracle_tables():
    from oracle import qnmfits_oracle as orc
    from qnmfits_b200 import synthetic
    return orc.OracleTables(synthetic.modes_cache)


@pytest.fixture(scope="session")
def reference():
    """The live reference module, or None when /root/reference is absent (GPU box)."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        return None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_loader.load_reference()


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.abs(b)))
