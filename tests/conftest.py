import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle.binding import REF_SO, RefLib
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libbliss_ref.so not built (needs /root/reference at build time)")
    return RefLib()


@pytest.fixture(scope="session")
def engine():
    import bliss_b200
    eng = bliss_b200.Engine(0)
    yield eng
    eng.close()


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
