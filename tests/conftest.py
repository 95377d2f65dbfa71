import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    import xinvert_b200
    if xinvert_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return xinvert_b200.default_context(0)
