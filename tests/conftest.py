import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_b200():
    """False only when the built library itself reports that there is no sm_100 device (SVB_ERR_NO_DEVICE). A missing or
    broken library is NOT a reason to skip: the GPU tests then fail loudly (there is no CPU fallback to hide behind)."""
    if os.environ.get("SEEKSV_B200_REQUIRE_GPU"):
        return True
    try:
        import ctypes
        import seeksv_b200
        L = seeksv_b200.load()
        h = ctypes.c_void_p()
        rc = L.svb_ctx_create(0, ctypes.byref(h))
        if rc == 0:
            L.svb_ctx_destroy(h)
        return rc != -1          # SVB_ERR_NO_DEVICE
    except Exception:
        return True


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _have_b200():
        return
    skip = pytest.mark.skip(reason="no B200 (sm_100) device: the CUDA path has no CPU fallback")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def read_text(path):
    with open(path, "r", encoding="latin-1", newline="") as f:
        return f.read()
