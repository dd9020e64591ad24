import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; oracle/mrlda_oracle.c)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def capi():
    from mrlda_b200 import capi
    capi.load()
    return capi


def small_corpus(D=60, V=300, K=8, mean_len=40, seed=1):
    from mrlda_b200 import corpus
    return corpus.generate(D, V, K, mean_len, seed=seed)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
