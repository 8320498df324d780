import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_visible():
    try:
        from qmps_b200 import _lib
        return _lib.load().qmps_device_count() > 0
    except Exception:  # noqa: BLE001 - library not built yet: the gpu tests cannot run either way
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a CPU-only box a plain `pytest tests` skips them instead of failing
    (the product itself still raises QmpsError without a device -- there is no CPU fallback to test)."""
    if not any("gpu" in it.keywords for it in items) or _gpu_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (gpu tests run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the in-tree libraries once per session."""
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")
    return {name[:-4]: np.load(os.path.join(gdir, name)) for name in os.listdir(gdir) if name.endswith(".npz")}
