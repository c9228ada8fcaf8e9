"""The "VlWide in lanes of a warp" experiment (tests/widelanes/widelanes.cu): both forms of
wide_bus.v's per-cycle work must agree bit for bit; the timings it prints are recorded in
profiles/README.md."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "widelanes", "widelanes")


@pytest.mark.parametrize("n,cycles", [(4096, 512), (520, 300)])
def test_thread_per_instance_and_lanes_per_limb_agree(n, cycles):
    if not os.path.exists(EXE):
        pytest.skip("tests/widelanes/widelanes not built")
    res = subprocess.run([EXE, str(n), str(cycles)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    got = json.loads(res.stdout)
    assert got["mismatches"] == 0 and got["distinct_neighbours"] > n * 0.99
