import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ours():
    from spgpu_b200 import capi
    return capi.lib()


@pytest.fixture(scope="session")
def oracle():
    from tests import util
    return util.oracle_lib()


@pytest.fixture(scope="session")
def ref():
    """The reference library itself (oracle/_ref/libspgpu_ref.so), if it was built."""
    from tests import util
    lib = util.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libspgpu_ref.so not built (needs /root/reference at build time)")
    return lib


@pytest.fixture(scope="session")
def gpu_handle(ours):
    import ctypes
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    h = ctypes.c_void_p()
    st = ours.spgpuCreate(ctypes.byref(h), 0)
    assert st == 0, f"spgpuCreate -> {st}"
    yield h
    ours.spgpuDestroy(h)


@pytest.fixture(scope="session")
def ref_handle(ref):
    import ctypes
    import torch
    assert torch.cuda.is_available()
    h = ctypes.c_void_p()
    st = ref.spgpuCreate(ctypes.byref(h), 0)
    assert st == 0
    yield h
    ref.spgpuDestroy(h)
