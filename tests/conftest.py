import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL(str(ROOT / "ampliconreconstructorom_b200" / "lib" / "libsegaligner_b200.so"))
        lib.sa_device_count.restype = ctypes.c_int
        return lib.sa_device_count() > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C ABI.  No fallback: a gpu-marked test on a box without the built library or
    without a B200 fails loudly."""
    from ampliconreconstructorom_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.Oracle()


@pytest.fixture(scope="session")
def ref():
    import helpers
    r = helpers.RefShim.try_load()
    if r is None:
        pytest.skip("oracle/_ref/libsa_ref.so not built (reference sources absent)")
    return r
