import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) device; run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """Builds (if stale) and loads the C-ABI library. No GPU needed to load it."""
    from cerberus_b200 import _lib, build
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def six_head_sd():
    from cerberus_b200 import synth
    return synth.make_state_dict(seed=0)
