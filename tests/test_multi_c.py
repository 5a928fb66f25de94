"""GPU: the multi-device C-ABI (blx_multi_*, include/blx.h) driven by a C programme (tests/c/test_multi.c), no Python in
the data path: sharded analysis byte-identical to one device, ncclAllGather + nearest neighbours equal to a brute-force
scan over bl_distance. Uses every visible GPU (one on the driver's test box; 2 / 4 / 8 under `gpurun --gpus N`)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_multi_device_c_programme():
    exe = os.path.join(ROOT, "tests", "c", "bin", "test_multi")
    if not os.path.exists(exe):
        pytest.skip("tests/c/bin/test_multi not built (make -C bliss_b200)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "PASS" in r.stdout and "devices=" in r.stdout
    print(r.stdout)
