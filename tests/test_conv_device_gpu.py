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
