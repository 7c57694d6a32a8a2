"""Parity at BASELINE.json's FULL sizes, where the CPU oracle would take minutes: size-
independent properties of the operator, checked through the C ABI on the device-built
matrices (which tests/test_device_build_gpu.py proves bit-identical to the host conversion
route at small sizes).

  * known answer: A * 1 for the stencil matrices is the row sum, known in closed form from the
    grid position (6 - #neighbours for the 7-point Laplacian, 26 - #neighbours for the
    27-point stencil) -- exact in floating point (small integers);
  * linearity: A(a*x + b*y) == a*A(x) + b*A(y) within the north_star tolerance;
  * two formats, one operator: the 27-point stencil stored as HDIA and as plain DIA must give
    the same product; HELL and its length-sorted OHELL twin (rIdx) must give the same product;
  * 8x smaller twin: the first rows of the 512^3 result equal the 512x512x64 slab result."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
T = util.TYPES


def _hell(L, h, sym, A, x, z, alpha=1.0, beta=0.0, y=None, ridx=None):
    t = T[sym]
    getattr(L, f"spgpu{sym}hellspmv")(h, z.data_ptr(), y.data_ptr() if y is not None else 0, t.scalar(alpha),
                                      A.values.data_ptr(), A.indices.data_ptr(), A.hack_size, A.hack_offsets.data_ptr(),
                                      A.rs.data_ptr(), ridx.data_ptr() if ridx is not None else 0, A.avg, A.nrows,
                                      x.data_ptr(), t.scalar(beta), A.base)


def test_cfg5_full_size_known_answer_and_linearity(ours, gpu_handle):
    """512^3 7-point Laplacian, double HELL (BASELINE configs[4])"""
    import torch
    from spgpu_b200 import device_build as DB
    n = 512
    A = DB.hell_laplace3d_7pt(n)
    assert A.nrows == 134217728 and A.nnz == 937951232            # R and nnz of cfg5
    N = A.nrows
    ones = torch.ones(N, dtype=torch.float64, device="cuda")
    z = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
    _hell(ours, gpu_handle, "D", A, ones, z)
    torch.cuda.synchronize()
    r = torch.arange(N, device="cuda")
    xg, yg, zg = r % n, (r // n) % n, r // (n * n)
    nb = ((xg > 0).to(torch.float64) + (xg < n - 1) + (yg > 0) + (yg < n - 1) + (zg > 0) + (zg < n - 1))
    assert torch.equal(z, 6.0 - nb)                               # exact: small integers
    del r, xg, yg, zg, nb, ones
    # linearity
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.rand(N, generator=g, device="cuda", dtype=torch.float64)
    y = torch.rand(N, generator=g, device="cuda", dtype=torch.float64)
    zx, zy = torch.empty_like(z), torch.empty_like(z)
    _hell(ours, gpu_handle, "D", A, x, zx)
    _hell(ours, gpu_handle, "D", A, y, zy)
    x.mul_(0.75).add_(y, alpha=-1.5)                              # x <- 0.75 x - 1.5 y
    _hell(ours, gpu_handle, "D", A, x, z)
    torch.cuda.synchronize()
    want = 0.75 * zx - 1.5 * zy
    # per-row scale: |a| sum|a_ik||x_k| <= 12 * (0.75 + 1.5)
    assert float((z - want).abs().max().item()) <= 1e-12 * 12 * 2.25
    # beta != 0 and in place: z <- 2 A x - 3 z
    zc = z.clone()
    _hell(ours, gpu_handle, "D", A, x, z, alpha=2.0, beta=-3.0, y=z)
    torch.cuda.synchronize()
    assert float((z - (2.0 * zc - 3.0 * zc)).abs().max().item()) <= 1e-12 * (2 * 27 + 3 * 27)
    # the 64-plane slab (1/8 of the rows, same structure) agrees on the rows that do not see the cut
    del zx, zy, want, zc
    S = DB.hell_laplace3d_7pt(n, 0, 64, nz=64)
    zs = torch.empty(S.nrows, dtype=torch.float64, device="cuda")
    z_full = torch.empty(N, dtype=torch.float64, device="cuda")
    _hell(ours, gpu_handle, "D", A, x, z_full)
    _hell(ours, gpu_handle, "D", S, x, zs)
    torch.cuda.synchronize()
    inner = 63 * n * n
    assert torch.equal(zs[:inner], z_full[:inner])


def test_cfg2_full_size_hdia_vs_dia_and_known_answer(ours, gpu_handle):
    """128^3 27-point stencil: HDIA (BASELINE configs[1]) and plain DIA are the same operator"""
    import torch
    from spgpu_b200 import device_build as DB
    n = 128
    H = DB.hdia_stencil27(n)
    D = DB.dia_stencil27(n)
    assert H.nnz == D.nnz == 55742968
    N = H.nrows
    t = T["D"]

    def hdia(x, z):
        ours.spgpuDhdiaspmv(gpu_handle, z.data_ptr(), 0, t.scalar(1.0), H.values.data_ptr(), H.offsets.data_ptr(), 32,
                            H.hack_offsets.data_ptr(), N, N, x.data_ptr(), t.scalar(0.0))

    def dia(x, z):
        ours.spgpuDdiaspmv(gpu_handle, z.data_ptr(), 0, t.scalar(1.0), D.values.data_ptr(), D.offsets.data_ptr(), D.pitch,
                           N, N, D.diags, x.data_ptr(), t.scalar(0.0))

    ones = torch.ones(N, dtype=torch.float64, device="cuda")
    z1 = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
    z2 = torch.full_like(z1, float("nan"))
    hdia(ones, z1); dia(ones, z2)
    torch.cuda.synchronize()
    r = torch.arange(N, device="cuda")
    cnt = torch.ones(N, dtype=torch.float64, device="cuda")
    for c in (r % n, (r // n) % n, r // (n * n)):
        cnt *= 1.0 + (c > 0).to(torch.float64) + (c < n - 1).to(torch.float64)      # cells of the 3x3x3 box inside the grid
    want = 26.0 - (cnt - 1.0)
    assert torch.equal(z1, want) and torch.equal(z2, want)
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    x = torch.rand(N, generator=g, device="cuda", dtype=torch.float64)
    hdia(x, z1); dia(x, z2)
    torch.cuda.synchronize()
    assert float((z1 - z2).abs().max().item()) <= 1e-12 * 52      # same diagonals, same order: expected equal


def test_cfg3_full_size_hell_vs_sorted_ohell(ours, gpu_handle):
    """4 M power-law rows, float HELL (BASELINE configs[2]) vs the same matrix length-sorted with rIdx"""
    import torch
    from spgpu_b200 import device_build as DB
    R = 1 << 22
    lens, cols, vals = DB.powerlaw_entries(R)
    A = DB.hell_from_rows(lens, cols, vals, R)
    assert int(A.rs.max().item()) == 4096 and 14 < A.nnz / R < 18
    sl, sc, sv, ridx = DB.sort_rows_by_length(lens, cols, vals)
    B = DB.hell_from_rows(sl, sc, sv, R)
    del lens, sl, sc, sv
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    x = torch.rand(R, generator=g, device="cuda", dtype=torch.float32)
    z1 = torch.full((R,), float("nan"), dtype=torch.float32, device="cuda")
    z2 = torch.full_like(z1, float("nan"))
    _hell(ours, gpu_handle, "S", A, x, z1)
    _hell(ours, gpu_handle, "S", B, x, z2, ridx=ridx)
    torch.cuda.synchronize()
    # scale per row: sum |a||x| computed with the same kernels on |A|, |x| (values are in (-1,1), x in (0,1))
    A.values.abs_()
    scale = torch.empty_like(z1)
    _hell(ours, gpu_handle, "S", A, x, scale)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(z1).all()) and bool(((z1 - z2).abs() <= 1e-5 * scale + 1e-30).all())
    # independent check of a sample of rows against float64 arithmetic on the raw entries
    start = torch.cumsum(A.rs.to(torch.int64), 0) - A.rs.to(torch.int64)
    xs = x.to(torch.float64)
    for row in (0, 16384, 16385, R - 1, 1234567):
        s0, ln = int(start[row].item()), int(A.rs[row].item())
        ref = float((vals[s0:s0 + ln].to(torch.float64) * xs[cols[s0:s0 + ln]]).sum().item())
        assert abs(float(z1[row].item()) - ref) <= 1e-5 * float(scale[row].item()) + 1e-30


def test_cfg4_full_size_linearity(ours, gpu_handle):
    """2 M rows banded complex-double HELL, alpha/beta != 1/0 (BASELINE configs[3])"""
    import torch
    from spgpu_b200 import device_build as DB
    R = 2_000_000
    lens, cols, vals = DB.banded_complex_entries(R)
    A = DB.hell_from_rows(lens, cols, vals, R)
    assert A.nnz == 80_000_000 - (40 * 0) or A.nnz > 79_000_000
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    mk = lambda: torch.complex(torch.rand(R, generator=g, device="cuda", dtype=torch.float64),
                               torch.rand(R, generator=g, device="cuda", dtype=torch.float64))
    x, y, w = mk(), mk(), mk()
    alpha, beta = 0.7 - 0.3j, -0.5 + 0.25j
    zx, zy, z = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    _hell(ours, gpu_handle, "Z", A, x, zx)
    _hell(ours, gpu_handle, "Z", A, y, zy)
    comb = (0.5 + 2j) * x - 1.25 * y
    _hell(ours, gpu_handle, "Z", A, comb, z, alpha=alpha, beta=beta, y=w)
    torch.cuda.synchronize()
    want = alpha * ((0.5 + 2j) * zx - 1.25 * zy) + beta * w
    # |a| <= sqrt(2), |x| <= sqrt(2), 40 entries per row, coefficients (|0.5+2j| + 1.25) * |alpha|, plus |beta||w|
    bound = 1e-12 * (40 * 2 * (2.07 + 1.25) * abs(alpha) + abs(beta) * 1.5)
    assert float((z - want).abs().max().item()) <= bound


def test_cfg2_full_size_device_coo_to_hdia_equals_the_analytic_builder(ours, gpu_handle):
    """the 55.7 M entries of the 128^3 27-point stencil as shuffled COO -> spgpuHdiaHackOffsetsFromCooDevice +
    spgpuDcooToHdiaDevice must give, bit for bit, the arrays of device_build.hdia_stencil27 (itself proven
    equal to the reference's cooToHdia at sizes the host route can run)"""
    import ctypes
    import torch
    from spgpu_b200 import device_build as DB
    n, hack = 128, 32
    A = DB.hdia_stencil27(n)
    N = n ** 3
    r = torch.arange(N, device="cuda", dtype=torch.int64)
    x, y, z = r % n, (r // n) % n, r // (n * n)
    rows, cols, vals = [], [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (z + dz >= 0) & (z + dz < n) & (y + dy >= 0) & (y + dy < n) & (x + dx >= 0) & (x + dx < n)
                rr = r[ok]
                rows.append(rr.to(torch.int32))
                cols.append((rr + (dz * n + dy) * n + dx).to(torch.int32))
                vals.append(torch.full((rr.numel(),), 26.0 if (dz, dy, dx) == (0, 0, 0) else -1.0, dtype=torch.float64, device="cuda"))
    rows, cols, vals = torch.cat(rows), torch.cat(cols), torch.cat(vals)
    assert rows.numel() == A.nnz == 55742968                     # nnz of cfg2
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    perm = torch.randperm(rows.numel(), generator=g, device="cuda")
    rows, cols, vals = rows[perm].contiguous(), cols[perm].contiguous(), vals[perm].contiguous()
    del perm, r, x, y, z
    hacks = N // hack
    hoff = torch.full((hacks + 1,), -1, dtype=torch.int32, device="cuda")
    height = ctypes.c_int(-1)
    rc = ours.spgpuHdiaHackOffsetsFromCooDevice(gpu_handle, ctypes.byref(height), hoff.data_ptr(), hack, N, N, rows.numel(),
                                                rows.data_ptr(), cols.data_ptr(), 0)
    assert rc == 0
    assert height.value == A.offsets.numel() == 1751088          # hack-diagonals of cfg2
    assert torch.equal(hoff, A.hack_offsets)
    hv = torch.zeros(height.value * hack, dtype=torch.float64, device="cuda")
    off = torch.full((height.value,), 123456789, dtype=torch.int32, device="cuda")
    rc = ours.spgpuDcooToHdiaDevice(gpu_handle, hv.data_ptr(), off.data_ptr(), hoff.data_ptr(), hack, N, N, rows.numel(),
                                    rows.data_ptr(), cols.data_ptr(), vals.data_ptr(), 0)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(off, A.offsets)
    assert torch.equal(hv, A.values)


def test_cfg5_full_size_device_csr_to_hell_roundtrip(ours, gpu_handle):
    """the 938 M entries of the 512^3 Laplacian as CSR -> spgpuCsrToHellLayoutDevice + spgpuDcsrToHellDevice must
    reproduce rS, hackOffsets and every live slot of device_build.hell_laplace3d_7pt; the OHELL layout of the same
    CSR multiplies to the same vector"""
    import ctypes
    import torch
    from spgpu_b200 import device_build as DB
    n = 512
    A = DB.hell_laplace3d_7pt(n)
    N = A.nrows
    # CSR of the same matrix, read back out of the HELL arrays: entry k of row i sits at hackOffsets[i/32] + 32k + i%32
    rs64 = A.rs.to(torch.int64)
    rowptr64 = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(rs64, 0, out=rowptr64[1:])
    nnz = int(rowptr64[-1].item())
    assert nnz == A.nnz == 937951232
    csr_cols = torch.empty(nnz, dtype=torch.int32, device="cuda")
    csr_vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    rows = torch.arange(N, device="cuda", dtype=torch.int64)
    at = A.hack_offsets.to(torch.int64)[rows // 32] + rows % 32
    for k in range(7):
        live = rs64 > k
        src = (at + 32 * k)[live]
        dst = (rowptr64[:-1] + k)[live]
        csr_cols[dst] = A.indices[src]
        csr_vals[dst] = A.values[src]
        del src, dst, live
    del rows, at
    rowptr = rowptr64.to(torch.int32)
    del rowptr64, rs64
    hacks = N // 32
    rs = torch.full((N,), -1, dtype=torch.int32, device="cuda")
    hoff = torch.full((hacks,), -1, dtype=torch.int32, device="cuda")
    total = ctypes.c_longlong(0)
    rc = ours.spgpuCsrToHellLayoutDevice(gpu_handle, N, rowptr.data_ptr(), 32, rs.data_ptr(), hoff.data_ptr(), ctypes.byref(total))
    assert rc == 0
    assert torch.equal(rs, A.rs) and torch.equal(hoff, A.hack_offsets) and total.value == A.values.numel()
    hv = torch.full((total.value,), float("nan"), dtype=torch.float64, device="cuda")
    hi = torch.full((total.value,), DB.POISON_INDEX, dtype=torch.int32, device="cuda")
    ours.spgpuDcsrToHellDevice(gpu_handle, N, rowptr.data_ptr(), csr_cols.data_ptr(), csr_vals.data_ptr(), 0, 32, hoff.data_ptr(), 0,
                               hv.data_ptr(), hi.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(hi, A.indices)                            # padding poisoned the same way by both builders
    assert torch.equal(hv.view(torch.int64), A.values.view(torch.int64))
    del hv, hi
    # OHELL of the same CSR: same product through rIdx
    ridx = torch.empty(N, dtype=torch.int32, device="cuda")
    rc = ours.spgpuCsrToOhellLayoutDevice(gpu_handle, N, rowptr.data_ptr(), 32, ridx.data_ptr(), rs.data_ptr(), hoff.data_ptr(),
                                          ctypes.byref(total))
    assert rc == 0 and int(rs[0].item()) == 7 and int(rs[-1].item()) == 4      # longest rows first, corners last
    ov = torch.full((total.value,), float("nan"), dtype=torch.float64, device="cuda")
    oi = torch.full((total.value,), DB.POISON_INDEX, dtype=torch.int32, device="cuda")
    ours.spgpuDcsrToOhellDevice(gpu_handle, N, rowptr.data_ptr(), csr_cols.data_ptr(), csr_vals.data_ptr(), 0, 32, hoff.data_ptr(),
                                ridx.data_ptr(), 0, ov.data_ptr(), oi.data_ptr())
    del csr_cols, csr_vals
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    x = torch.rand(N, generator=g, device="cuda", dtype=torch.float64)
    z0 = torch.empty(N, dtype=torch.float64, device="cuda")
    z1 = torch.full((N,), float("nan"), dtype=torch.float64, device="cuda")
    _hell(ours, gpu_handle, "D", A, x, z0)
    O = DB.DevHell(ov, oi, hoff, rs, 32, N, N, nnz, 0, 7, ridx)
    _hell(ours, gpu_handle, "D", O, x, z1, ridx=ridx)
    torch.cuda.synchronize()
    # same 7 products per row, possibly summed in another order: |dz| <= 1e-12 * sum|a||x| <= 1e-12 * 12
    assert float((z1 - z0).abs().max().item()) <= 1e-12 * 12
