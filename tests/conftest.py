import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """TEST INFRASTRUCTURE: the CPU restatement (oracle/vlq_oracle.c) through ctypes."""
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def small_model(oracle):
    """A small VLQ model trained by the oracle (C=256, E=32, M=16, d=128) + encoded base + queries."""
    from vector_line_quantization_b200 import data

    xt = data.sift_like(16384, kc=512, seed=1)
    xb = data.sift_like(20000, kc=512, seed=2)
    xq = data.sift_like(64, kc=512, seed=3)
    m = oracle.train_all(xt, nlist=256, E=32, M=16, nL=256, niter=6, pq_niter=8)
    m.update(xt=xt, xb=xb, xq=xq, d=128, C=256, E=32, M=16)
    return m


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vector_line_quantization_b200 import _abi

    _abi.lib()  # fail loudly if the extension is missing
    return torch.device("cuda:0")
