import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# A peer-flag wait that can never be satisfied (a bug in the exchange protocol) must fail the test in
# seconds, not after the library's 10-minute default (csrc/p2p.cu, k_wait).
os.environ.setdefault("NB_P2P_TIMEOUT_MS", "30000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes binding of libnbody_b200.so).  Fails, never skips, when the
    library is missing: there is no fallback path to test instead."""
    mod = importlib.import_module("procedural-universe_b200")
    mod.load()
    return mod


@pytest.fixture(scope="session")
def particle_dtype(pkg):
    return pkg.PARTICLE_DTYPE


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def as_particles(u8, dtype):
    return np.ascontiguousarray(u8).view(dtype).reshape(-1)


def rel_err(a, b):
    """Per-body relative error of 3-vectors: |a - b| / |b|."""
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def same_particles(a, b):
    """Field-by-field bit equality (the 4 padding bytes at offset 44 are not data)."""
    return all(np.array_equal(a[f], b[f]) for f in a.dtype.names)
