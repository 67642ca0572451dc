import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (C restatement of the reference, oracle/ssq_oracle.c)."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def sq():
    """The product package; GPU tests fail loudly if the CUDA library cannot be used."""
    import torch
    import shortseq_b200
    from shortseq_b200 import _lib
    _lib.lib()
    assert torch.cuda.is_available(), "GPU test started without a CUDA device"
    return shortseq_b200
