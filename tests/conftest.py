import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import orc

    orc.build()
    return orc.lib()


@pytest.fixture(scope="session")
def product_lib():
    from pota_b200 import build, camera

    if not os.path.exists(build.LIB):
        build.build()
    return camera.lib()


@pytest.fixture(params=["unrolled", "table"])
def kernel_kind(request, monkeypatch):
    """Both evaluator kinds behind the same C ABI: per-lens unrolled kernels and the table-driven ones."""
    monkeypatch.setenv("LB_FORCE_TABLE", "1" if request.param == "table" else "0")
    return request.param
