import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def O():
    """the parity oracle (test infrastructure)"""
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def gcnb():
    """ctypes binding of libgcn_b200.so; GPU tests fail loudly if it is missing or the device is not sm_100."""
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    b = importlib.import_module("parallel_gcn_b200.binding")
    return b


@pytest.fixture(scope="session")
def dev(gcnb):
    import torch
    assert torch.cuda.is_available(), "GPU test without a CUDA device"
    gcnb.device_check()
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def datasets(O):
    out = {}
    for name in ("cora", "citeseer"):
        out[name] = O.parse_dataset(os.path.join(ROOT, "data", name))
    return out
