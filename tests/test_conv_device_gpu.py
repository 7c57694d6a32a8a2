"""Device-side CSR -> HELL construction (spgpuCsrToHellLayoutDevice + spgpu?csrToHellDevice)
must produce bit for bit what the reference's host route (cooToEll + ellToHell, through our
bit-exact C port) produces from the same entries."""
import ctypes

import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("hack", [32, 64])
@pytest.mark.parametrize("base", [0, 1])
def test_csr_to_hell_device_matches_host_route(ours, gpu_handle, dtype, hack, base):
    import torch
    s = util.sym_of(dtype)
    for coo in (G.random_coo(3001, 2500, (0, 17), 3, dtype, base), G.powerlaw(4000, 6, 500, 512, 2, np.float32)):
        coo = F.Coo(coo.rows - coo.base + base, coo.cols - coo.base + base, coo.vals.astype(dtype),
                    coo.nrows, coo.ncols, base)
        host = F.ell_to_hell(F.coo_to_ell(coo, base), hack)
        # CSR of the same entries (COO is row-sorted by construction)
        counts = np.bincount(coo.rows - base, minlength=coo.nrows)
        rowptr = (np.concatenate(([0], np.cumsum(counts))) + base).astype(np.int32)
        d_rowptr, d_cols, d_vals = util.to_dev(rowptr), util.to_dev(coo.cols), util.to_dev(coo.vals)
        hacks = (coo.nrows + hack - 1) // hack
        d_rs = torch.zeros(coo.nrows, dtype=torch.int32, device="cuda")
        d_ho = torch.zeros(hacks, dtype=torch.int32, device="cuda")
        total = ctypes.c_longlong(0)
        rc = ours.spgpuCsrToHellLayoutDevice(gpu_handle, coo.nrows, d_rowptr.data_ptr(), hack, d_rs.data_ptr(),
                                             d_ho.data_ptr(), ctypes.byref(total))
        assert rc == 0
        assert total.value == host.values.shape[0] == host.height * hack
        np.testing.assert_array_equal(d_rs.cpu().numpy(), host.rs)
        np.testing.assert_array_equal(d_ho.cpu().numpy(), host.hack_offsets)
        tdt = {"S": torch.float32, "D": torch.float64, "C": torch.complex64, "Z": torch.complex128}[s]
        d_hv = torch.full((max(total.value, 1),), float("nan"), dtype=tdt, device="cuda")
        d_hi = torch.full((max(total.value, 1),), -(2 ** 30), dtype=torch.int32, device="cuda")
        getattr(ours, f"spgpu{s}csrToHellDevice")(gpu_handle, coo.nrows, d_rowptr.data_ptr(), d_cols.data_ptr(),
                                                   d_vals.data_ptr(), base, hack, d_ho.data_ptr(), base,
                                                   d_hv.data_ptr(), d_hi.data_ptr())
        torch.cuda.synchronize()
        # the host wrapper poisons padding the same way, so the whole arrays must be identical
        np.testing.assert_array_equal(d_hi.cpu().numpy()[:total.value], host.indices)
        np.testing.assert_array_equal(d_hv.cpu().numpy()[:total.value].view(np.uint8), host.values.view(np.uint8))


@pytest.mark.parametrize("dtype", [np.float32, np.complex128])
@pytest.mark.parametrize("hack", [32, 64])
def test_csr_to_ohell_device_matches_ell_to_oell(ours, gpu_handle, dtype, hack):
    """spgpuCsrToOhellLayoutDevice + spgpu?csrToOhellDevice == cooToEll + ellToOell + ellToHell (host):
    same rIdx (the reference mergesort's tie order: equal lengths by descending row), rS, hackOffsets,
    and the same slot for every entry"""
    import torch
    s = util.sym_of(dtype)
    for coo in (G.random_coo(2999, 2500, (0, 17), 3, dtype, 0), G.powerlaw(4000, 6, 500, 512, 2, np.float32),
                G.laplace2d_5pt(37, 41)):
        coo = F.Coo(coo.rows, coo.cols, coo.vals.astype(dtype), coo.nrows, coo.ncols, 0)
        oell = F.ell_to_oell(F.coo_to_ell(coo, 0))
        host = F.ell_to_hell(oell, hack)
        counts = np.bincount(coo.rows, minlength=coo.nrows)
        rowptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
        d_rowptr, d_cols, d_vals = util.to_dev(rowptr), util.to_dev(coo.cols), util.to_dev(coo.vals)
        hacks = (coo.nrows + hack - 1) // hack
        d_ridx = torch.zeros(coo.nrows, dtype=torch.int32, device="cuda")
        d_rs = torch.zeros(coo.nrows, dtype=torch.int32, device="cuda")
        d_ho = torch.zeros(hacks, dtype=torch.int32, device="cuda")
        total = ctypes.c_longlong(0)
        rc = ours.spgpuCsrToOhellLayoutDevice(gpu_handle, coo.nrows, d_rowptr.data_ptr(), hack, d_ridx.data_ptr(),
                                              d_rs.data_ptr(), d_ho.data_ptr(), ctypes.byref(total))
        assert rc == 0
        np.testing.assert_array_equal(d_ridx.cpu().numpy(), oell.ridx)
        np.testing.assert_array_equal(d_rs.cpu().numpy(), oell.rs)
        np.testing.assert_array_equal(d_ho.cpu().numpy(), host.hack_offsets)
        assert total.value == host.values.shape[0]
        tdt = {"S": torch.float32, "D": torch.float64, "C": torch.complex64, "Z": torch.complex128}[s]
        d_hv = torch.full((max(total.value, 1),), float("nan"), dtype=tdt, device="cuda")
        d_hi = torch.full((max(total.value, 1),), -(2 ** 30), dtype=torch.int32, device="cuda")
        getattr(ours, f"spgpu{s}csrToOhellDevice")(gpu_handle, coo.nrows, d_rowptr.data_ptr(), d_cols.data_ptr(),
                                                    d_vals.data_ptr(), 0, hack, d_ho.data_ptr(), d_ridx.data_ptr(), 0,
                                                    d_hv.data_ptr(), d_hi.data_ptr())
        torch.cuda.synchronize()
        # compare the live slots (the host route's padding comes from its ELL intermediate)
        live = np.zeros(total.value, dtype=bool)
        ho, rs = host.hack_offsets, host.rs
        for i in range(coo.nrows):
            at = ho[i // hack] + i % hack
            live[at + hack * np.arange(rs[i])] = True
        np.testing.assert_array_equal(d_hi.cpu().numpy()[:total.value][live], host.indices[live])
        np.testing.assert_array_equal(d_hv.cpu().numpy()[:total.value][live], host.values[live])
        assert (d_hi.cpu().numpy()[:total.value][~live] == -(2 ** 30)).all()          # padding untouched
        # and the product through the C ABI with rIdx equals the oracle on the unsorted matrix
        x = G.random_vector(coo.ncols, dtype, 1, -1, 1)
        y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
        dx, dy = util.to_dev(x), util.to_dev(y)
        dz = torch.zeros_like(dy)
        t = util.TYPES[s]
        getattr(ours, f"spgpu{s}hellspmv")(gpu_handle, dz.data_ptr(), dy.data_ptr(), t.scalar(1.5), d_hv.data_ptr(),
                                           d_hi.data_ptr(), hack, d_ho.data_ptr(), d_rs.data_ptr(), d_ridx.data_ptr(),
                                           max(1, coo.nnz // coo.nrows), coo.nrows, dx.data_ptr(), t.scalar(0.5), 0)
        torch.cuda.synchronize()
        want = util.oracle_spmv("ell", F.coo_to_ell(coo, 0), x, y, 1.5, 0.5)
        util.assert_rows_close(dz.cpu().numpy(), want, util.row_scale(coo, x, y, 1.5, 0.5), s, "ohell device")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("hack", [32, 64])
@pytest.mark.parametrize("base", [0, 1])
def test_coo_to_hdia_device_matches_host_route(ours, gpu_handle, dtype, hack, base):
    """spgpuHdiaHackOffsetsFromCooDevice + spgpu?cooToHdiaDevice == computeHdiaHackOffsetsFromCoo +
    cooToHdia (host): hackOffsets, offsets and every cell bit for bit, with the COO entries shuffled"""
    import torch
    s = util.sym_of(dtype)
    rng = np.random.default_rng(5)
    for coo in (G.stencil3d_27pt(11), G.laplace2d_5pt(45, 31), G.random_coo(700, 900, (0, 9), 4, dtype, 0),
                G.random_coo(33, 20, (0, 4), 1, dtype, 0)):
        vals = coo.vals if (np.dtype(dtype).kind == "c" or coo.vals.dtype.kind != "c") else coo.vals.real
        coo = F.Coo(coo.rows - coo.base + base, coo.cols - coo.base + base, vals.astype(dtype), coo.nrows, coo.ncols, base)
        host = F.coo_to_hdia(coo, hack)
        perm = rng.permutation(coo.nnz)
        d_rows, d_cols, d_vals = util.to_dev(coo.rows[perm]), util.to_dev(coo.cols[perm]), util.to_dev(coo.vals[perm])
        hacks = (coo.nrows + hack - 1) // hack
        d_ho = torch.full((hacks + 1,), -1, dtype=torch.int32, device="cuda")
        height = ctypes.c_int(-1)
        rc = ours.spgpuHdiaHackOffsetsFromCooDevice(gpu_handle, ctypes.byref(height), d_ho.data_ptr(), hack, coo.nrows,
                                                    coo.ncols, coo.nnz, d_rows.data_ptr(), d_cols.data_ptr(), base)
        assert rc == 0
        assert height.value == host.height
        np.testing.assert_array_equal(d_ho.cpu().numpy(), host.hack_offsets)
        tdt = {"S": torch.float32, "D": torch.float64, "C": torch.complex64, "Z": torch.complex128}[s]
        d_hv = torch.zeros(max(height.value * hack, 1), dtype=tdt, device="cuda")
        d_off = torch.full((max(height.value, 1),), 12345, dtype=torch.int32, device="cuda")
        rc = getattr(ours, f"spgpu{s}cooToHdiaDevice")(gpu_handle, d_hv.data_ptr(), d_off.data_ptr(), d_ho.data_ptr(), hack,
                                                       coo.nrows, coo.ncols, coo.nnz, d_rows.data_ptr(), d_cols.data_ptr(),
                                                       d_vals.data_ptr(), base)
        assert rc == 0
        torch.cuda.synchronize()
        np.testing.assert_array_equal(d_off.cpu().numpy()[:height.value], host.offsets)
        np.testing.assert_array_equal(d_hv.cpu().numpy()[:height.value * hack].view(np.uint8), host.values.view(np.uint8))


def test_coo_to_hdia_device_empty_matrix(ours, gpu_handle):
    import torch
    d_ho = torch.full((5,), -1, dtype=torch.int32, device="cuda")
    height = ctypes.c_int(-1)
    rc = ours.spgpuHdiaHackOffsetsFromCooDevice(gpu_handle, ctypes.byref(height), d_ho.data_ptr(), 32, 100, 100, 0, 0, 0, 0)
    assert rc == 0 and height.value == 0
    assert (d_ho.cpu().numpy() == 0).all()
