import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def lo():
    import linearoperators_jl_b200 as lo
    return lo


@pytest.fixture(scope="session")
def ctx(lo):
    if not _have_cuda():
        pytest.skip("no CUDA device")
    return lo.default_context()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    oracle.set_mode(True, 1)
    return oracle
