import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    import sylver_b200
    from sylver_b200 import build
    build.build()
    return sylver_b200.lib()


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/liboracle.so not built (needs /root/reference; run `make -C oracle`)")
    ref.lib()
    return ref
