import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def mvr():
    import mvr_b200
    return mvr_b200


@pytest.fixture(scope="session")
def synth():
    import mvr_b200
    import mvr_b200.synth as s
    return s


@pytest.fixture(scope="session")
def ctx(mvr):
    c = mvr.Context(0)   # raises without a CUDA device: GPU tests never fall back
    yield c
    c.close()
