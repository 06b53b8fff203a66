import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def solver():
    """One CUDA solver handle for the whole GPU session; fails loudly when the library or GPU is missing."""
    from er3t_b200.solver import Solver
    s = Solver(device=0)
    yield s
    s.close()
