import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'small_cases.npz'))


@pytest.fixture(scope='session')
def digests():
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'full_size_digests.json')) as f:
        return json.load(f)['digests']


@pytest.fixture
def libenv(monkeypatch):
    """The library reads its VFA_* debug switches once per process; this sets / clears one and makes it re-read them
    (and restores the environment and the library's view of it afterwards)."""
    import vfa_b200

    class _Env:
        def setenv(self, key, value):
            monkeypatch.setenv(key, str(value))
            vfa_b200.reload_env()

        def delenv(self, key, raising=False):
            monkeypatch.delenv(key, raising=raising)
            vfa_b200.reload_env()

    yield _Env()
    monkeypatch.undo()
    vfa_b200.reload_env()
