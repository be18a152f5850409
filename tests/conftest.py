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
