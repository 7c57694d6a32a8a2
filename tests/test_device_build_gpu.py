"""The on-device builders of the full-size benchmark matrices
(spgpu_b200/device_build.py) must produce, bit for bit, what the reference's
host conversion route (COO -> cooToEll -> ellToHell / cooToHdia, through our
bit-exact C port) produces -- checked at sizes where both can run."""
import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu


def _same_hell(dev, host):
    import torch
    v, i, ho, rs = (t.cpu().numpy() for t in (dev.values, dev.indices, dev.hack_offsets, dev.rs))
    np.testing.assert_array_equal(rs, host.rs)
    np.testing.assert_array_equal(ho, host.hack_offsets)
    assert v.shape == host.values.shape
    # compare only the live slots (padding is undefined in both)
    hs = host.hack_size
    for h in range(ho.shape[0]):
        rows = rs[h * hs:(h + 1) * hs]
        for k in range(int(rows.max()) if rows.size else 0):
            live = np.nonzero(rows > k)[0] + int(ho[h]) + k * hs
            np.testing.assert_array_equal(i[live], host.indices[live])
            np.testing.assert_array_equal(v[live].view(np.uint8), host.values[live].view(np.uint8))


@pytest.mark.parametrize("n", [32, 64])
def test_laplace7_hell_matches_host_route(n):
    from spgpu_b200 import device_build as DB
    host = F.ell_to_hell(F.coo_to_ell(G.laplace3d_7pt(n)), 32)
    dev = DB.hell_laplace3d_7pt(n)
    assert dev.nnz == int(host.rs.sum())
    _same_hell(dev, host)


def test_laplace7_slab_with_local_columns():
    """a z-slab with local (x_ext) column numbering == the same rows of the global
    matrix with columns shifted by (row_lo - plane)"""
    from spgpu_b200 import device_build as DB
    n, z_lo, z_hi = 32, 8, 16
    plane = n * n
    glob = DB.hell_laplace3d_7pt(n)
    slab = DB.hell_laplace3d_7pt(n, z_lo, z_hi, local_columns=True)
    assert slab.nrows == (z_hi - z_lo) * plane and slab.ncols == (z_hi - z_lo + 2) * plane
    gv, gi, gho, grs = DB.to_host_hell(glob)
    sv, si, sho, srs = DB.to_host_hell(slab)
    r0 = z_lo * plane
    np.testing.assert_array_equal(srs, grs[r0:r0 + slab.nrows])
    e0 = int(gho[r0 // 32])
    np.testing.assert_array_equal(sho, gho[r0 // 32:(r0 + slab.nrows) // 32] - e0)
    live = ~np.isnan(sv)
    np.testing.assert_array_equal(live, ~np.isnan(gv[e0:e0 + sv.shape[0]]))
    np.testing.assert_array_equal(sv[live], gv[e0:e0 + sv.shape[0]][live])
    np.testing.assert_array_equal(si[live], gi[e0:e0 + sv.shape[0]][live] - (r0 - plane))


@pytest.mark.parametrize("n", [32, 64])
def test_stencil27_hdia_matches_host_route(n):
    from spgpu_b200 import device_build as DB
    coo = G.stencil3d_27pt(n)
    host = F.coo_to_hdia(coo, 32)
    dev = DB.hdia_stencil27(n)
    np.testing.assert_array_equal(dev.hack_offsets.cpu().numpy(), host.hack_offsets)
    np.testing.assert_array_equal(dev.offsets.cpu().numpy(), host.offsets)
    np.testing.assert_array_equal(dev.values.cpu().numpy().view(np.uint8), host.values.view(np.uint8))
    assert dev.nnz == coo.nnz
    i = np.repeat(np.arange(host.nrows // 32), np.diff(host.hack_offsets))[:, None] * 32 + np.arange(32)[None, :]
    c = i + host.offsets[:, None]
    assert dev.cells_in_range == int(((c >= 0) & (c < host.ncols)).sum())


def test_survey_counts_cfg2():
    """128^3 27-point HDIA hack 32 -> 65 536 hacks, 1 751 088 hack-diagonals,
    56 034 816 stored cells, 55 742 968 non-zeros"""
    from spgpu_b200 import device_build as DB
    d = DB.hdia_stencil27(128)
    assert d.hack_offsets.numel() == 65536 + 1
    assert d.offsets.numel() == 1751088
    assert d.values.numel() == 56034816
    assert d.nnz == 55742968


@pytest.mark.parametrize("which", ["powerlaw", "banded"])
def test_entry_builders_match_host_route(which):
    import torch
    from spgpu_b200 import device_build as DB
    if which == "powerlaw":
        R = 20000
        lens, cols, vals = DB.powerlaw_entries(R, mean=8, maxlen=600, spike_every=4096)
        base = 0
    else:
        R = 20000
        lens, cols, vals = DB.banded_complex_entries(R, per_row=12, bw=300)
        base = 1
    dev = DB.hell_from_rows(lens, cols, vals, R, 32, base)
    rows = torch.repeat_interleave(torch.arange(R, device=cols.device), lens).cpu().numpy()
    c = cols.cpu().numpy()
    # columns are distinct and ascending inside each row
    same_row = rows[1:] == rows[:-1]
    assert (np.diff(c)[same_row] > 0).all() and c.min() >= 0 and c.max() < R
    coo = F.Coo((rows + base).astype(np.int32), (c + base).astype(np.int32), vals.cpu().numpy(), R, R, base)
    host = F.ell_to_hell(F.coo_to_ell(coo, base), 32)
    _same_hell(dev, host)


def test_sort_rows_by_length_matches_ell_to_oell():
    """device OHELL ordering == the reference's ellToOell permutation (ties: higher row first)"""
    import torch
    from spgpu_b200 import device_build as DB
    R = 3000
    lens, cols, vals = DB.powerlaw_entries(R, mean=5, maxlen=300, spike_every=512)
    slens, scols, svals, ridx = DB.sort_rows_by_length(lens, cols, vals)
    rows = torch.repeat_interleave(torch.arange(R, device=cols.device), lens).cpu().numpy()
    coo = F.Coo(rows.astype(np.int32), cols.cpu().numpy().astype(np.int32), vals.cpu().numpy(), R, R, 0)
    ell = F.coo_to_ell(coo)
    oell = F.ell_to_oell(ell)
    np.testing.assert_array_equal(ridx.cpu().numpy(), oell.ridx)
    np.testing.assert_array_equal(slens.cpu().numpy().astype(np.int32), oell.rs)
    dev = DB.hell_from_rows(slens, scols, svals, R, 32, 0)
    _same_hell(dev, F.ell_to_hell(oell, 32))
