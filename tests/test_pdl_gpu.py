"""Programmatic dependent launch (tuning key `pdl`, csrc/launch.cuh: spgpu_launch_dep): the kernels of the hot path are
launched so that the CTAs of kernel k+1 take the slots the last wave of kernel k leaves and wait there
(griddepcontrol.wait) until kernel k has completed and its writes are visible.  The results must not depend on it:
a chain in which every kernel consumes what the previous one wrote gives the same bits with the key on and off,
on the handle's own stream, on a caller's stream, and when the chain is captured into a CUDA graph."""
import ctypes

import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu


def _chain(ours, h, fmt, A, dA, x0, steps):
    """x <- A x / 8 + x (SpMV -> axpby -> scal of a device scalar -> dot), `steps` times; returns (x, dots)"""
    import torch
    n = A.nrows
    T = util.TYPES["D"]
    x = util.to_dev(x0.copy())
    y = torch.zeros_like(x)
    dots = torch.zeros(steps, dtype=torch.float64, device="cuda")
    for k in range(steps):
        util.dev_spmv_ptr(ours, h, fmt, A, dA, y, x, None, 0.125, 0.0)
        ours.spgpuDaxpby(h, x.data_ptr(), n, T.scalar(1.0), x.data_ptr(), T.scalar(1.0), y.data_ptr())
        ours.spgpuDdotDev(h, n, x.data_ptr(), y.data_ptr(), dots.data_ptr() + 8 * k)
        ours.spgpuDscal(h, y.data_ptr(), n, T.scalar(0.5), x.data_ptr())         # y is rewritten: the next SpMV overwrites it again
    return x, dots


@pytest.mark.parametrize("fmt", ["hell", "hdia", "ell", "dia"])
def test_dependent_chain_gives_the_same_bits_with_and_without_pdl(ours, gpu_handle, fmt):
    import torch
    coo = G.laplace3d_7pt(40)                                    # 64000 rows: 500 CTAs, kernels of a few microseconds
    A = {"hell": lambda: F.ell_to_hell(F.coo_to_ell(coo), 32), "hdia": lambda: F.coo_to_hdia(coo, 32),
         "ell": lambda: F.coo_to_ell(coo), "dia": lambda: F.coo_to_dia(coo)}[fmt]()
    dA = util.upload(A)
    x0 = G.random_vector(coo.nrows, np.float64, 3, -1, 1)
    got = {}
    try:
        for pdl in (0, 1, 0, 1):
            assert ours.spgpuSetTuning(gpu_handle, b"pdl", pdl) == 0
            x, dots = _chain(ours, gpu_handle, fmt, A, dA, x0, 25)
            torch.cuda.synchronize()
            res = (x.cpu().numpy().copy(), dots.cpu().numpy().copy())
            if pdl in got:
                assert np.array_equal(res[0], got[pdl][0]) and np.array_equal(res[1], got[pdl][1])
            got[pdl] = res
        assert np.array_equal(got[0][0], got[1][0])
        assert np.array_equal(got[0][1], got[1][1])
        assert np.isfinite(got[0][0]).all() and np.abs(got[0][1]).min() > 0
        # and the chain is what it should be: one step against the oracle
        y = util.oracle_spmv(fmt, A, x0, None, 0.125, 0.0)
        x1, _ = _chain(ours, gpu_handle, fmt, A, dA, x0, 1)
        torch.cuda.synchronize()
        np.testing.assert_allclose(x1.cpu().numpy(), x0 + y, rtol=1e-13, atol=1e-14)
    finally:
        ours.spgpuSetTuning(gpu_handle, b"pdl", 1)


def test_pdl_chain_on_a_caller_stream_and_in_a_cuda_graph(ours, gpu_handle):
    import torch
    coo = G.laplace3d_7pt(32)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    dA = util.upload(A)
    x0 = G.random_vector(coo.nrows, np.float64, 5, -1, 1)
    assert ours.spgpuGetTuning(gpu_handle, b"pdl") == 1          # the default
    ref_x, ref_d = _chain(ours, gpu_handle, "hell", A, dA, x0, 10)
    torch.cuda.synchronize()
    ref_x, ref_d = ref_x.cpu().numpy(), ref_d.cpu().numpy()
    stream = torch.cuda.Stream()
    try:
        ours.spgpuSetStream(gpu_handle, stream.cuda_stream)
        with torch.cuda.stream(stream):
            x, d = _chain(ours, gpu_handle, "hell", A, dA, x0, 10)
        stream.synchronize()
        assert np.array_equal(x.cpu().numpy(), ref_x) and np.array_equal(d.cpu().numpy(), ref_d)
        # captured: the launches become kernel nodes (with programmatic edges where the driver supports them)
        n = A.nrows
        T = util.TYPES["D"]
        xg = util.to_dev(x0.copy())
        yg = torch.zeros_like(xg)
        dg = torch.zeros(1, dtype=torch.float64, device="cuda")
        ours.spgpuReserveScratch(gpu_handle, 1 << 20)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            with torch.cuda.graph(g, stream=stream):
                util.dev_spmv_ptr(ours, gpu_handle, "hell", A, dA, yg, xg, None, 0.125, 0.0)
                ours.spgpuDaxpby(gpu_handle, xg.data_ptr(), n, T.scalar(1.0), xg.data_ptr(), T.scalar(1.0), yg.data_ptr())
                ours.spgpuDdotDev(gpu_handle, n, xg.data_ptr(), yg.data_ptr(), dg.data_ptr())
                ours.spgpuDscal(gpu_handle, yg.data_ptr(), n, T.scalar(0.5), xg.data_ptr())
            for _ in range(10):
                g.replay()
        stream.synchronize()
        assert np.array_equal(xg.cpu().numpy(), ref_x)
        assert dg.cpu().numpy()[0] == ref_d[9]
    finally:
        ours.spgpuSetStream(gpu_handle, None)
