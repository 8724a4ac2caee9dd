import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(scope='session')
def oracle_c():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope='session')
def cuda_lib():
    """Build (if stale) and load the CUDA library; GPU tests must fail -- not skip -- if it cannot be loaded."""
    from viterbi_spl_b200 import _lib, build
    build.build()
    return _lib.load()
