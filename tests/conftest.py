import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))

from oracle import oracle as O  # noqa: E402  (test infrastructure: the CPU checker)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running statistical test")


@pytest.fixture(scope="session")
def golden():
    arrays = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    with open(os.path.join(GOLDEN, "reference_vectors.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference (oracle/_ref/libmcref.so); ~5.6 GB, one per session."""
    if not O.have_reference_lib():
        pytest.skip("oracle/_ref/libmcref.so not built (needs /root/reference: make -C oracle ref)")
    import psutil
    if psutil.virtual_memory().available < 8 * 2 ** 30:
        pytest.skip("not enough free memory for the reference's 5.6 GB arrays")
    r = O.RefOracle()
    yield r
    r.close()


def system_from_meta(d):
    return O.make_system(**d)


@pytest.fixture(scope="session")
def mcpm():
    import mcpm_b200
    return mcpm_b200


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / den) if a.size else 0.0
