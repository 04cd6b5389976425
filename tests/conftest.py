import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: a -m gpu test that also runs ~1 minute of CPU oracle at a BASELINE config size")


@pytest.fixture(scope='session')
def golden():
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz')))


@pytest.fixture(scope='session')
def gweights():
    from gen_common import golden_weights
    return golden_weights(7)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
