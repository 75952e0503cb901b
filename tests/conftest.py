import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: long CPU test, opt-in with GM_SLOW=1")


@pytest.fixture(scope="session")
def gm():
    from _pkg import gm as pkg
    return pkg


@pytest.fixture(scope="session")
def gpu_lib(gm):
    """Builds nothing on the GPU box: the .so travels with the snapshot.  Fails loudly if it is
    missing -- there is no CPU path."""
    import torch  # noqa: F401  (device plumbing only)
    assert os.path.exists(gm.capi.LIB_PATH), "libgemotion_b200.so missing: run __graft_entry__.build()"
    return gm.capi.lib()
