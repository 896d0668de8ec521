import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu on the GPU box)")
    # a fresh checkout has no libegobox_gpu.so (built artefacts are git-ignored): build it once (nvcc cross-compiles
    # sm_100a without a GPU); the product package itself never builds or falls back silently
    from egobox_b200 import _build
    if not os.path.exists(_build.LIB):
        _build.build_library()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
