import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def built():
    """Builds the oracle and the product library (make is incremental: a stale binary after a source edit would let
    the tests pass against old code).  On a box without nvcc the prebuilt library that travelled with the snapshot is
    used as it is."""
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
    lib = os.path.join(ROOT, 'unicycler_b200', 'libunicycler_b200.so')
    import shutil
    if shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc') or not os.path.isfile(lib):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'unicycler_b200', 'csrc')])
    return True


@pytest.fixture(scope='session')
def oracle(built):
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope='session')
def ub(built):
    import unicycler_b200
    unicycler_b200.load_library()
    return unicycler_b200
