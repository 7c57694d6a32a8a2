"""GPU parity: every spgpu?{ell,hell,dia,hdia}spmv entry point of OUR library,
called through the C ABI on device buffers, against the CPU oracle on the same
seeded inputs.  Tolerance (north_star): per row 1e-5 (S/C) / 1e-12 (D/Z) of
|alpha| sum|a||x| + |beta||y|."""
import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


def scalars(dtype, beta_zero=False):
    if np.dtype(dtype).kind == "c":
        return (0.7 - 0.3j), (0.0 if beta_zero else (-0.5 + 0.25j))
    return 2.0, (0.0 if beta_zero else -3.0)


def build(fmt, coo, base, hack):
    ell = F.coo_to_ell(coo, base)
    if fmt == "ell":
        return ell
    if fmt == "hell":
        return F.ell_to_hell(ell, hack)
    if fmt == "dia":
        return F.coo_to_dia(coo)
    return F.coo_to_hdia(coo, hack)


def check(ours, h, fmt, coo, A, x, y, alpha, beta, **kw):
    s = util.sym_of(A.values.dtype)
    dA = util.upload(A)
    got = util.dev_spmv(ours, h, fmt, A, dA, x, y, alpha, beta, **kw)
    okw = {k: v for k, v in kw.items() if k in ("base", "ridx", "rs_null")}
    want = util.oracle_spmv(fmt, A, x, y if (beta != 0) else None, alpha, beta, **okw)
    util.assert_rows_close(got, want, util.row_scale(coo, x, y, alpha, beta), s, f"{fmt}/{s}/{kw}")
    return got


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fmt", ["ell", "hell", "dia", "hdia"])
@pytest.mark.parametrize("base", [0, 1])
def test_random_ragged(ours, gpu_handle, dtype, fmt, base):
    """ragged rows incl. empty ones, rows not a multiple of 32/hack, rect matrix"""
    for nrows, ncols, hack in [(1, 7, 32), (31, 64, 32), (97, 83, 32), (1000, 1111, 64), (4099, 4099, 32)]:
        coo = G.random_coo(nrows, ncols, (0, 13), nrows, dtype, base)
        A = build(fmt, coo, base, hack)
        x = G.random_vector(ncols, dtype, 1, -1, 1)
        y = G.random_vector(nrows, dtype, 2, -1, 1)
        alpha, beta = scalars(dtype)
        check(ours, gpu_handle, fmt, coo, A, x, y, alpha, beta)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fmt", ["ell", "hell", "dia", "hdia"])
def test_beta_zero_never_reads_y(ours, gpu_handle, dtype, fmt):
    coo = G.random_coo(515, 515, (0, 9), 9, dtype, 0)
    A = build(fmt, coo, 0, 32)
    x = G.random_vector(515, dtype, 1, -1, 1)
    y = np.full(515, np.nan, dtype=dtype)
    alpha, _ = scalars(dtype)
    got = check(ours, gpu_handle, fmt, coo, A, x, y, alpha, 0.0)
    assert np.isfinite(got.view(util.real_of(dtype))).all()
    # and with y == NULL
    s = util.sym_of(dtype)
    dA = util.upload(A)
    got2 = util.dev_spmv(ours, gpu_handle, fmt, A, dA, x, None, alpha, 0.0)
    np.testing.assert_array_equal(got, got2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fmt", ["ell", "hell", "dia", "hdia"])
def test_inplace_z_aliases_y(ours, gpu_handle, dtype, fmt):
    coo = G.random_coo(777, 777, (1, 9), 11, dtype, 0)
    A = build(fmt, coo, 0, 32)
    x = G.random_vector(777, dtype, 1, -1, 1)
    y = G.random_vector(777, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    check(ours, gpu_handle, fmt, coo, A, x, y, alpha, beta, inplace=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("fmt", ["ell", "hell"])
def test_ridx_permutation(ours, gpu_handle, dtype, fmt):
    """OELL / OHELL: rows sorted by length, rIdx sends results home
    (reference hellPerf.cpp ELL == HELL == OELL)"""
    coo = G.random_coo(1500, 1500, (0, 40), 5, dtype, 0)
    ell = F.coo_to_ell(coo)
    oell = F.ell_to_oell(ell)
    A = oell if fmt == "ell" else F.ell_to_hell(oell, 32)
    x = G.random_vector(1500, dtype, 1, -1, 1)
    y = G.random_vector(1500, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    got = check(ours, gpu_handle, fmt, coo, A, x, y, alpha, beta, ridx=oell.ridx)
    plain = util.oracle_spmv("ell", ell, x, y, alpha, beta)
    util.assert_rows_close(got, plain, util.row_scale(coo, x, y, alpha, beta), util.sym_of(dtype), "ridx vs plain")


@pytest.mark.parametrize("dtype", DTYPES)
def test_ell_without_row_sizes(ours, gpu_handle, dtype):
    """rS == NULL: maxNnzPerRow slots per row, zero padding (reference hellPerf -DNO_ROW_SIZE)"""
    coo = G.random_coo(999, 999, (0, 7), 3, dtype, 0)
    A = F.coo_to_ell(coo)
    x = G.random_vector(999, dtype, 1, -1, 1)
    y = G.random_vector(999, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    check(ours, gpu_handle, "ell", coo, A, x, y, alpha, beta, rs_null=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("fmt", ["ell", "hell"])
def test_spike_rows_take_the_cooperative_path(ours, gpu_handle, dtype, fmt):
    """cfg3 in small: power-law lengths with spike rows far above the average ->
    phase 2 of spmv_slots.cuh; HELL padding is NaN / invalid indices (formats.py)"""
    coo = G.powerlaw(6000, mean=8, maxlen=1500, spike_every=512, seed=3, dtype=dtype)
    A = build(fmt, coo, 0, 32)
    x = G.random_vector(6000, dtype, 1, -1, 1)
    y = G.random_vector(6000, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    for avg in (1, 8, 4000):          # spike threshold 32 / 32 / 16000 slots: cooperative phase on .. off
        check(ours, gpu_handle, fmt, coo, A, x, y, alpha, beta, avg=avg)
    # length-sorted rows (OHELL): homogeneous long hacks must stay in the row-per-lane walk
    ell = F.coo_to_ell(coo)
    oell = F.ell_to_oell(ell)
    B = oell if fmt == "ell" else F.ell_to_hell(oell, 32)
    check(ours, gpu_handle, fmt, coo, B, x, y, alpha, beta, ridx=oell.ridx, avg=8)


@pytest.mark.parametrize("hack", [32, 64, 128])
@pytest.mark.parametrize("fmt", ["hell", "hdia"])
def test_hack_sizes(ours, gpu_handle, fmt, hack):
    coo = G.stencil3d_27pt(11)
    A = build(fmt, coo, 0, hack)
    n = coo.nrows
    x = G.random_vector(n, np.float64, 1, 0, 1)
    y = G.random_vector(n, np.float64, 2, 0, 1)
    check(ours, gpu_handle, fmt, coo, A, x, y, 1.0, 0.5)


@pytest.mark.parametrize("name,fmt,dtype", [
    ("cfg1", "ell", np.float64), ("cfg2", "hdia", np.float64), ("cfg2", "dia", np.float64),
    ("cfg3", "hell", np.float32), ("cfg4", "hell", np.complex128), ("cfg5", "hell", np.float64)])
def test_baseline_configs_scaled_down(ours, gpu_handle, name, fmt, dtype):
    """the five BASELINE.json configurations at sizes the oracle finishes in seconds"""
    coo = {"cfg1": lambda: G.laplace2d_5pt(300),
           "cfg2": lambda: G.stencil3d_27pt(40),
           "cfg3": lambda: G.powerlaw(1 << 17, 16, 4096, 32768, 7, np.float32),
           "cfg4": lambda: G.banded_complex(60000, 40, 1000, 11),
           "cfg5": lambda: G.laplace3d_7pt(48)}[name]()
    A = build(fmt, coo, 0, 32)
    n = coo.nrows
    x = G.random_vector(n, dtype, 12345, 0, 1)
    y = G.random_vector(n, dtype, 54321, 0, 1)
    if name == "cfg4":
        alpha, beta = (0.7 - 0.3j), (-0.5 + 0.25j)
    else:
        alpha, beta = 1.0, 0.0
    check(ours, gpu_handle, fmt, coo, A, x, y, alpha, beta)
    if name == "cfg4":          # baseIndex exercised at 1 too
        coo1 = G.banded_complex(60000, 40, 1000, 11, base=1)
        A1 = build(fmt, coo1, 1, 32)
        check(ours, gpu_handle, fmt, coo1, A1, x, y, alpha, beta)


def test_empty_matrix_is_a_noop(ours, gpu_handle):
    import torch
    z = torch.full((4,), 7.0, dtype=torch.float64, device="cuda")
    t = util.TYPES["D"]
    ours.spgpuDhellspmv(gpu_handle, z.data_ptr(), 0, t.scalar(1.0), 0, 0, 32, 0, 0, 0, 1, 0, 0, t.scalar(0.0), 0)
    ours.spgpuDellspmv(gpu_handle, z.data_ptr(), 0, t.scalar(1.0), 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, t.scalar(0.0), 0)
    ours.spgpuDdiaspmv(gpu_handle, z.data_ptr(), 0, t.scalar(1.0), 0, 0, 0, 0, 0, 0, 0, t.scalar(0.0))
    ours.spgpuDhdiaspmv(gpu_handle, z.data_ptr(), 0, t.scalar(1.0), 0, 0, 32, 0, 0, 0, 0, t.scalar(0.0))
    torch.cuda.synchronize()
    assert (z == 7.0).all()


def test_custom_stream(ours, gpu_handle):
    """spgpuSetStream / spgpuGetStream (reference core.c:62-78): work follows the stream"""
    import ctypes
    import torch
    st = ctypes.c_void_p()
    ours.spgpuStreamCreate(gpu_handle, ctypes.byref(st))
    default = ours.spgpuGetStream(gpu_handle)
    ours.spgpuSetStream(gpu_handle, st)
    assert ours.spgpuGetStream(gpu_handle) == st.value
    coo = G.laplace2d_5pt(64)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    x = G.random_vector(coo.nrows, np.float64, 1)
    check(ours, gpu_handle, "hell", coo, A, x, x, 1.0, 0.0)
    ours.spgpuSetStream(gpu_handle, None)
    assert ours.spgpuGetStream(gpu_handle) == default
    ours.spgpuStreamDestroy(st)


@pytest.mark.parametrize("variant,occ", [(1, 64), (2, 64), (2, 256), (3, 0)])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hack", [32, 64])
def test_hell_kernel_variants(ours, gpu_handle, variant, occ, dtype, hack):
    """every HELL code path (predicated / unpredicated slab reads / 32 vs 48 warps per SM /
    the bulk-async TMA pipeline incl. its oversize-tile and tail-tile fallbacks)
    must give the oracle's result; padding is NaN / invalid indices"""
    try:
        assert ours.spgpuSetTuning(gpu_handle, b"hellVariant", variant) == 0
        assert ours.spgpuSetTuning(gpu_handle, b"hellBlock", occ) == 0
        mats = [G.laplace3d_7pt(20), G.random_coo(5000, 5000, (0, 13), 1, dtype, 0),
                G.powerlaw(9000, mean=6, maxlen=900, spike_every=700, seed=5, dtype=dtype)]
        for coo in mats:
            coo = F.Coo(coo.rows, coo.cols, (coo.vals if np.dtype(dtype).kind == "c" or coo.vals.dtype.kind != "c" else coo.vals.real).astype(dtype), coo.nrows, coo.ncols, coo.base)
            A = build("hell", coo, 0, hack)
            x = G.random_vector(coo.ncols, dtype, 1, -1, 1)
            y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
            alpha, beta = scalars(dtype)
            check(ours, gpu_handle, "hell", coo, A, x, y, alpha, beta)
            check(ours, gpu_handle, "hell", coo, A, x, y, alpha, 0.0)
    finally:
        ours.spgpuSetTuning(gpu_handle, b"hellVariant", 0)
        ours.spgpuSetTuning(gpu_handle, b"hellBlock", 0)


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("dtype", DTYPES)
def test_ell_kernel_variants(ours, gpu_handle, variant, dtype):
    """ELL code paths: predicated / unpredicated slab reads / bulk-async pipeline (incl. ragged tail
    tile, rS == NULL and rIdx)"""
    try:
        assert ours.spgpuSetTuning(gpu_handle, b"hellVariant", variant) == 0
        for coo in (G.laplace3d_7pt(20), G.random_coo(5000, 5000, (0, 5), 1, dtype, 0), G.laplace2d_5pt(70, 53)):
            coo = F.Coo(coo.rows, coo.cols, (coo.vals if np.dtype(dtype).kind == "c" or coo.vals.dtype.kind != "c" else coo.vals.real).astype(dtype), coo.nrows, coo.ncols, coo.base)
            A = build("ell", coo, 0, 32)
            x = G.random_vector(coo.ncols, dtype, 1, -1, 1)
            y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
            alpha, beta = scalars(dtype)
            check(ours, gpu_handle, "ell", coo, A, x, y, alpha, beta)
            check(ours, gpu_handle, "ell", coo, A, x, y, alpha, 0.0, rs_null=True)
            oell = F.ell_to_oell(A)
            check(ours, gpu_handle, "ell", coo, oell, x, y, alpha, beta, ridx=oell.ridx)
    finally:
        ours.spgpuSetTuning(gpu_handle, b"hellVariant", 0)


@pytest.mark.parametrize("rows_per_lane", [-1, 0, 2])
@pytest.mark.parametrize("dtype", DTYPES)
def test_ell_short_rows_several_rows_per_lane(ours, gpu_handle, dtype, rows_per_lane):
    """ELL, short regular rows, general kernel (-1), exact-slot-count kernel with 1 (default) / 2 rows per lane (ellRows): ragged last CTA, rS == NULL, rIdx,
    in-place beta; the random matrix (avg far below max) must fall back to the general kernel"""
    try:
        assert ours.spgpuSetTuning(gpu_handle, b"ellRows", rows_per_lane) == 0
        for coo in (G.laplace2d_5pt(70, 53), G.laplace3d_7pt(17), G.random_coo(3000, 3000, (0, 5), 1, dtype, 0)):
            coo = F.Coo(coo.rows, coo.cols, (coo.vals if np.dtype(dtype).kind == "c" or coo.vals.dtype.kind != "c" else coo.vals.real).astype(dtype), coo.nrows, coo.ncols, coo.base)
            A = build("ell", coo, 0, 32)
            x = G.random_vector(coo.ncols, dtype, 1, -1, 1)
            y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
            alpha, beta = scalars(dtype)
            check(ours, gpu_handle, "ell", coo, A, x, y, alpha, beta)
            check(ours, gpu_handle, "ell", coo, A, x, y, alpha, 0.0, rs_null=True)
            oell = F.ell_to_oell(A)
            check(ours, gpu_handle, "ell", coo, oell, x, y, alpha, beta, ridx=oell.ridx)
    finally:
        ours.spgpuSetTuning(gpu_handle, b"ellRows", 0)


@pytest.mark.parametrize("variant,occ", [(0, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (6, 5), (7, 0), (7, 3), (0, 8), (0, 64), (0, 160), (0, 176), (0, 192), (0, 224), (0, 256)])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hack", [32, 64])
def test_hdia_kernel_variants(ours, gpu_handle, variant, occ, dtype, hack):
    """HDIA code paths: direct x loads at four occupancy levels, the variant that stages the
    x windows of runs of consecutive offsets in shared memory (stencils have such runs, the random
    matrix has none, the rectangular one exercises the column range test), the bulk-async pipeline,
    the persistent one, and the per-warp slab variants (6/7; occ = diagonals per slice there, so 5
    and 3 force hacks to be walked in several bulk copies)"""
    try:
        assert ours.spgpuSetTuning(gpu_handle, b"hdiaVariant", variant) == 0
        assert ours.spgpuSetTuning(gpu_handle, b"hdiaBlock", occ) == 0
        for coo in (G.stencil3d_27pt(12), G.stencil3d_27pt(20), G.laplace2d_5pt(61, 47), G.random_coo(700, 900, (0, 9), 4, dtype, 0),
                    G.banded_complex(3000, 45, 30, 5)):
            coo = F.Coo(coo.rows, coo.cols, (coo.vals if np.dtype(dtype).kind == "c" or coo.vals.dtype.kind != "c" else coo.vals.real).astype(dtype), coo.nrows, coo.ncols, coo.base)
            A = build("hdia", coo, 0, hack)
            x = G.random_vector(coo.ncols, dtype, 1, -1, 1)
            y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
            alpha, beta = scalars(dtype)
            check(ours, gpu_handle, "hdia", coo, A, x, y, alpha, beta)
            check(ours, gpu_handle, "hdia", coo, A, x, y, alpha, 0.0)
    finally:
        ours.spgpuSetTuning(gpu_handle, b"hdiaVariant", 0)
        ours.spgpuSetTuning(gpu_handle, b"hdiaBlock", 0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("split", [-1, 1, 8])
def test_hell_long_hack_split_mode(ours, gpu_handle, dtype, split):
    """split mode: slots beyond 64 of a long hack are queued as chunks for the tail kernel (atomic
    adds into z); -1 = off, 1 = on, 8 = on with a queue of 8 items (overflow: the warp that cannot
    queue walks its rows itself).  Sorted (rIdx) and unsorted matrices, beta != 0, in place."""
    try:
        assert ours.spgpuSetTuning(gpu_handle, b"hellSplit", split) == 0
        coo = G.powerlaw(6000, mean=8, maxlen=1500, spike_every=256, seed=4, dtype=np.float32)
        coo = F.Coo(coo.rows, coo.cols, coo.vals.astype(dtype), coo.nrows, coo.ncols, coo.base)
        x = G.random_vector(6000, dtype, 1, -1, 1)
        y = G.random_vector(6000, dtype, 2, -1, 1)
        alpha, beta = scalars(dtype)
        ell = F.coo_to_ell(coo)
        A = F.ell_to_hell(ell, 32)
        check(ours, gpu_handle, "hell", coo, A, x, y, alpha, beta, avg=8)
        check(ours, gpu_handle, "hell", coo, A, x, y, alpha, beta, avg=8, inplace=True)
        oell = F.ell_to_oell(ell)
        B = F.ell_to_hell(oell, 32)
        check(ours, gpu_handle, "hell", coo, B, x, y, alpha, beta, ridx=oell.ridx, avg=8)
        check(ours, gpu_handle, "hell", coo, B, x, y, alpha, 0.0, ridx=oell.ridx, avg=8)
    finally:
        ours.spgpuSetTuning(gpu_handle, b"hellSplit", 0)


@pytest.mark.parametrize("dtype", [np.float32, np.complex128])
def test_split_mode_is_bit_reproducible(ours, gpu_handle, dtype):
    """the long-hack split mode folds its partial sums in chunk order (no floating-point atomics): ten runs of the
    same product give the same bits, whichever warp took which chunk"""
    try:
        assert ours.spgpuSetTuning(gpu_handle, b"hellSplit", 1) == 0
        coo = G.powerlaw(20000, mean=10, maxlen=3000, spike_every=128, seed=8, dtype=np.float32)
        coo = F.Coo(coo.rows, coo.cols, coo.vals.astype(dtype), coo.nrows, coo.ncols, coo.base)
        ell = F.coo_to_ell(coo)
        oell = F.ell_to_oell(ell)
        A = F.ell_to_hell(oell, 32)
        x = G.random_vector(coo.ncols, dtype, 1, -1, 1)
        y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
        alpha, beta = scalars(dtype)
        dA = util.upload(A)
        first = util.dev_spmv(ours, gpu_handle, "hell", A, dA, x, y, alpha, beta, ridx=oell.ridx)
        for _ in range(9):
            again = util.dev_spmv(ours, gpu_handle, "hell", A, dA, x, y, alpha, beta, ridx=oell.ridx)
            np.testing.assert_array_equal(first.view(np.uint8), again.view(np.uint8))
        want = util.oracle_spmv("hell", A, x, y, alpha, beta, ridx=oell.ridx)
        util.assert_rows_close(first, want, util.row_scale(coo, x, y, alpha, beta), util.sym_of(dtype), "split mode")
    finally:
        ours.spgpuSetTuning(gpu_handle, b"hellSplit", 0)
