import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count() -> int:
    try:
        from _dip_pkg import load
        import ctypes as C
        n = C.c_int(0)
        rc = load().lib().dip_device_count(C.byref(n))
        return n.value if rc == 0 else 0
    except Exception:  # noqa: BLE001
        return 0


@pytest.fixture(scope="session")
def dip():
    from _dip_pkg import load
    return load()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx(dip):
    if _cuda_device_count() == 0:
        pytest.fail("a test marked gpu ran without a CUDA device; the product path has no CPU fallback")
    c = dip.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "small_cases.npz"))


@pytest.fixture(scope="session")
def golden_generic():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "generic_cases.npz"))


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)
