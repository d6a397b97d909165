import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session')
def oracle():
    from oracle import reference_path as rp
    rp.build()
    return rp


@pytest.fixture(scope='session')
def cuda_lib():
    """The product library must be present on a GPU box: fail loudly, never skip."""
    import torch
    assert torch.cuda.is_available(), 'GPU test run without a CUDA device'
    from synchrad_b200 import _lib
    return _lib.load()


def rel_errors(got, ref):
    import numpy as np
    m = np.abs(ref).max()
    n2 = np.linalg.norm(ref)
    if m == 0:
        return float(np.abs(got).max()), float(np.linalg.norm(got))
    return float(np.abs(got - ref).max() / m), float(np.linalg.norm(got - ref) / n2)
