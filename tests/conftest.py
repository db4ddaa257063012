import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    from kinetica_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "libcrn_oracle.so")):
        g.build()
    return True


def _have_gpu():
    try:
        from kinetica_b200 import _lib
        h = _lib.Handle(0)
        h.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device: skip them (instead of failing in kb2_create) where there is none."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)
