import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


@pytest.fixture(scope='session')
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _default_env():
    import mdpy_b200 as md
    md.env.set_default()
    yield
    md.env.set_default()


def rel_rms(a, b):
    """sqrt(sum |a-b|^2 / sum |b|^2) — the north star's per-atom force error measure."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))
