"""GPU parity against the REFERENCE'S OWN KERNELS: oracle/_ref/libspgpu_ref.so is
the reference library (its unmodified sources compiled for sm_100a, see
oracle/Makefile) loaded next to ours; both are driven through the identical C
ABI on the same device buffers.  This is what pins both our kernels and the CPU
oracle to the reference."""
import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


def scalars(dtype):
    if np.dtype(dtype).kind == "c":
        return (0.7 - 0.3j), (-0.5 + 0.25j)
    return 2.0, -3.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fmt", ["ell", "hell", "dia", "hdia"])
@pytest.mark.parametrize("beta_zero", [False, True])
def test_three_way(ours, ref, gpu_handle, ref_handle, dtype, fmt, beta_zero):
    s = util.sym_of(dtype)
    # full last block only for HELL: the reference's 2-thread HELL path has a barrier
    # under a divergent guard; rows % 128 == 0 keeps it well-defined
    nrows = 2048
    coo = G.random_coo(nrows, nrows, (1, 24), 21, dtype, 0)
    ell = F.coo_to_ell(coo)
    A = {"ell": lambda: ell, "hell": lambda: F.ell_to_hell(ell, 32),
         "dia": lambda: F.coo_to_dia(coo), "hdia": lambda: F.coo_to_hdia(coo, 32)}[fmt]()
    x = G.random_vector(nrows, dtype, 1, -1, 1)
    y = G.random_vector(nrows, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    if beta_zero:
        beta = 0.0
    dA = util.upload(A)
    z_ref = util.dev_spmv(ref, ref_handle, fmt, A, dA, x, y, alpha, beta)
    z_our = util.dev_spmv(ours, gpu_handle, fmt, A, dA, x, y, alpha, beta)
    z_orc = util.oracle_spmv(fmt, A, x, y if beta != 0 else None, alpha, beta)
    scale = util.row_scale(coo, x, y, alpha, beta)
    util.assert_rows_close(z_our, z_ref, scale, s, f"ours vs reference kernels {fmt}/{s}")
    util.assert_rows_close(z_orc, z_ref, scale, s, f"oracle vs reference kernels {fmt}/{s}")


@pytest.mark.parametrize("fmt,avg", [("ell", 4), ("ell", 15), ("ell", 40), ("hell", 40)])
def test_reference_variants(ours, ref, gpu_handle, ref_handle, fmt, avg):
    """the reference picks _vanilla / _texcache / _texcache_prefetch and the
    2-threads-per-row split from avgNnzPerRow (ell_spmv_base.cuh:138-143);
    all of them must agree with us"""
    nrows = 4096
    coo = G.random_coo(nrows, nrows, (max(1, avg - 3), avg + 3), 4, np.float64, 0, empty_rows=False)
    ell = F.coo_to_ell(coo)
    A = ell if fmt == "ell" else F.ell_to_hell(ell, 32)
    x = G.random_vector(nrows, np.float64, 1, -1, 1)
    y = G.random_vector(nrows, np.float64, 2, -1, 1)
    dA = util.upload(A)
    z_ref = util.dev_spmv(ref, ref_handle, fmt, A, dA, x, y, 1.5, 0.5, avg=avg)
    z_our = util.dev_spmv(ours, gpu_handle, fmt, A, dA, x, y, 1.5, 0.5, avg=avg)
    util.assert_rows_close(z_our, z_ref, util.row_scale(coo, x, y, 1.5, 0.5), "D", f"{fmt} avg={avg}")


@pytest.mark.parametrize("name,fmt,dtype", [
    ("cfg1", "ell", np.float64), ("cfg2", "hdia", np.float64),
    ("cfg4", "hell", np.complex128), ("cfg5", "hell", np.float64)])
def test_configs_vs_reference(ours, ref, gpu_handle, ref_handle, name, fmt, dtype):
    coo = {"cfg1": lambda: G.laplace2d_5pt(256), "cfg2": lambda: G.stencil3d_27pt(32),
           "cfg4": lambda: G.banded_complex(32768, 40, 1000, 11),
           "cfg5": lambda: G.laplace3d_7pt(32)}[name]()
    ell = F.coo_to_ell(coo)
    A = {"ell": lambda: ell, "hell": lambda: F.ell_to_hell(ell, 32),
         "hdia": lambda: F.coo_to_hdia(coo, 32)}[fmt]()
    n = coo.nrows
    x = G.random_vector(n, dtype, 12345, 0, 1)
    y = G.random_vector(n, dtype, 54321, 0, 1)
    alpha, beta = ((0.7 - 0.3j), (-0.5 + 0.25j)) if name == "cfg4" else (1.0, 0.0)
    dA = util.upload(A)
    z_ref = util.dev_spmv(ref, ref_handle, fmt, A, dA, x, y, alpha, beta)
    z_our = util.dev_spmv(ours, gpu_handle, fmt, A, dA, x, y, alpha, beta)
    util.assert_rows_close(z_our, z_ref, util.row_scale(coo, x, y, alpha, beta), util.sym_of(dtype), name)


@pytest.mark.parametrize("dtype", DTYPES)
def test_blas1_vs_reference(ours, ref, gpu_handle, ref_handle, dtype):
    import torch
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    n = 100003
    x = G.random_vector(n, dtype, 1, -1, 1)
    y = G.random_vector(n, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    dx, dy = util.to_dev(x), util.to_dev(y)
    outs = []
    for L, h in ((ours, gpu_handle), (ref, ref_handle)):
        dz = torch.zeros_like(dx)
        getattr(L, f"spgpu{s}axpby")(h, dz.data_ptr(), n, t.scalar(beta), dy.data_ptr(), t.scalar(alpha), dx.data_ptr())
        dw = torch.zeros_like(dx)
        getattr(L, f"spgpu{s}scal")(h, dw.data_ptr(), n, t.scalar(alpha), dx.data_ptr())
        d = t.from_c(getattr(L, f"spgpu{s}dot")(h, n, dx.data_ptr(), dy.data_ptr()))
        nr = getattr(L, f"spgpu{s}nrm2")(h, n, dx.data_ptr())
        torch.cuda.synchronize()
        outs.append((dz.cpu().numpy(), dw.cpu().numpy(), d, nr))
    (z1, w1, d1, n1), (z2, w2, d2, n2) = outs
    tol = util.TOL[s]
    np.testing.assert_allclose(z1, z2, rtol=0, atol=tol * 4)
    np.testing.assert_allclose(w1, w2, rtol=0, atol=tol * 4)
    mag = float(np.sum(np.abs(x.astype(np.complex128)) * np.abs(y.astype(np.complex128))))
    assert abs(d1 - d2) <= (1e-4 if s in "SC" else 1e-12) * mag
    assert abs(n1 - n2) <= (1e-5 if s in "SC" else 1e-12) * n2


def test_reference_amax_defect_is_documented(ours, ref, gpu_handle, ref_handle):
    """the reference's amax only reduces the elements whose index is
    0 or 1 (mod 32) of each block partial; we return the true maximum.  A vector
    whose maximum sits at index 5 shows the difference."""
    import torch
    n = 4096
    x = np.full(n, 0.5, dtype=np.float64)
    x[5] = 9.0
    dx = util.to_dev(x)
    assert ours.spgpuDamax(gpu_handle, n, dx.data_ptr()) == 9.0
    assert ref.spgpuDamax(ref_handle, n, dx.data_ptr()) != 9.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("base", [0, 1])
def test_ellcsput_against_the_reference_kernel(ours, ref, gpu_handle, ref_handle, dtype, base):
    """spgpu?ellcsput (reference ell.h:194-302, kernels/ell_csput_base.cuh:33-75): per (aI, aJ, aVal) triple the slot of
    row aI - baseIndex whose stored column equals aJ is overwritten.  The reference's quirks are the contract and are
    pinned here against its own kernel: `alpha` is ignored (ell_csput_base.cuh:42-44 never reads it), aJ is compared
    with the STORED index as is (no base adjustment, :66), triples whose row is negative are skipped (:46-47), and a
    column that is absent from the row changes nothing.  Result must be bit-identical."""
    import torch
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    nrows = 3000
    coo = G.random_coo(nrows, nrows, (0, 20), 77, dtype, 0)             # columns ascending inside a row
    ell = F.coo_to_ell(coo)                                             # stored indices are 0-based
    rng = np.random.default_rng(base + 10)
    # half of the triples hit existing entries, a quarter name absent columns, the rest sit in row -1
    k = rng.choice(coo.nnz, size=coo.nnz // 2, replace=False)
    ai = np.concatenate([coo.rows[k] + base, rng.integers(0, nrows, 400) + base, np.full(50, base - 1)]).astype(np.int32)
    aj = np.concatenate([coo.cols[k], rng.integers(0, nrows, 400), rng.integers(0, nrows, 50)]).astype(np.int32)
    av = G.random_vector(ai.shape[0], dtype, 5, -1, 1)
    # duplicates of one (row, column) would race (last writer wins in either library): keep the first
    _, first = np.unique(ai.astype(np.int64) * nrows + aj, return_index=True)
    ai, aj, av = ai[first], aj[first], av[first]
    results = []
    for L, h in ((ours, gpu_handle), (ref, ref_handle)):
        d_vals, d_idx, d_rs = util.to_dev(ell.values), util.to_dev(ell.indices), util.to_dev(ell.rs)
        d_i, d_j, d_v = util.to_dev(ai), util.to_dev(aj), util.to_dev(av)
        getattr(L, f"spgpu{s}ellcsput")(h, t.scalar(123.0), d_vals.data_ptr(), d_idx.data_ptr(), ell.pitch, ell.pitch,
                                        d_rs.data_ptr(), ai.shape[0], d_i.data_ptr(), d_j.data_ptr(), d_v.data_ptr(), base)
        torch.cuda.synchronize()
        results.append(d_vals.cpu().numpy())
    assert np.array_equal(results[0].view(np.uint8), results[1].view(np.uint8))
    # and the host statement of the same rule
    want = ell.values.copy()
    changed = 0
    for r, c, v in zip(ai - base, aj, av):
        if r < 0:
            continue
        cols = ell.indices[r + np.arange(ell.rs[r]) * ell.pitch]
        hit = np.nonzero(cols == c)[0]
        if hit.size:
            want[r + hit[0] * ell.pitch] = v
            changed += 1
    assert changed >= coo.nnz // 2 and np.array_equal(results[0].view(np.uint8), want.view(np.uint8))
