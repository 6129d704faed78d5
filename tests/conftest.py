import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc.Oracle("canonical")


@pytest.fixture(scope="session")
def oracle_shipped():
    import oracle as orc
    orc.build()
    return orc.Oracle("shipped")


@pytest.fixture(scope="session")
def ctx():
    """jpezy_b200 context on cuda:0 -- fails loudly when the CUDA library is not usable."""
    import torch
    import jpezy_b200 as J
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.init()
    c = J.Context(0)
    yield c
    c.close()
