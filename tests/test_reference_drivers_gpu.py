"""Drop-in proof: the REFERENCE'S OWN test programs (src/tests/ctest.c,
testSparseVector.c), compiled unmodified against include/ and linked to OUR
libspgpu.so by oracle/Makefile (target ref-tests), must pass on the GPU."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name):
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout + p.stderr


@pytest.mark.parametrize("name", ["testSpVec_s_ours", "testSpVec_d_ours"])
def test_reference_sparse_vector_program(name):
    rc, out = _run(name)
    assert rc == 0, out
    assert "Test Passed (Scatter operation)" in out and "Test Passed (Gather operation)" in out, out


def test_reference_ctest_program():
    """ctest.c prints dot(z,z) after the ELL and after the HELL SpMV; the two
    numbers must be identical (that is the reference's own acceptance check)."""
    rc, out = _run("ctest_ours")
    assert rc == 0, out
    dots = re.findall(r"dot res: (\S+)", out)
    assert len(dots) == 2 and dots[0] == dots[1] and float(dots[0]) > 0, out
    assert "Error" not in out, out
