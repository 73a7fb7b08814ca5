import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "point-cloud-interpolation-_b200")
for p in (ROOT, PKG_DIR, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the product has no fallback: make sure the native library is what runs
    from b200pc import _lib
    _lib.load()
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    from oracle import strict
    strict.build()


@pytest.fixture(autouse=True)
def _fresh_tuning():
    """the library caches its B200PC_* tuning variables; re-read them once a test that changed the environment
    (monkeypatch.setenv + ops.reload_tuning()) is over, so knobs never leak into the next test"""
    yield
    from b200pc import _lib
    if _lib._lib is not None:
        _lib._lib.b200pc_tuning_reload()
