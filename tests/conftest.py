import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def mcbb():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def gpu(mcbb):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    mcbb.check(mcbb.lib().mcbb_init(0))
    return mcbb
