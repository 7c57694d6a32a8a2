"""The C-level multi-GPU API (include/spgpu_mg.h, csrc/mg.c): ONE process drives N ranks.

On a one-GPU box the ranks are the SAME device listed several times: the partition, the column remap, the
vectors with their halo zones and the EVENTS exchange (push kernels + CUDA events between the ranks' streams)
all run for real, only the transport is local.  With two or more GPUs the same tests also run with distinct
devices, i.e. the FUSED exchange (NVLink peer stores inside the SpMV kernel) -- tests/test_mg_multi_gpu.py.
Reference: the single-GPU product of the same library (spgpu?hellspmv on the whole matrix) and the CPU oracle."""
import ctypes

import numpy as np
import pytest

from spgpu_b200 import capi, formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


def scalars(dtype):
    if np.dtype(dtype).kind == "c":
        return (0.7 - 0.3j), (-0.5 + 0.25j)
    return 2.0, -3.0


class Mg:
    """thin ctypes convenience over the C API"""

    def __init__(self, L, devices, exchange=capi.MG_AUTO):
        self.L = L
        self.h = ctypes.c_void_p()
        devs = (ctypes.c_int * len(devices))(*devices)
        st = L.spgpuMgCreate(ctypes.byref(self.h), devs, len(devices))
        assert st == 0, f"spgpuMgCreate -> {st}"
        assert L.spgpuMgSetExchange(self.h, exchange) == 0
        self.world = L.spgpuMgWorld(self.h)

    def matrix(self, hell):
        s = util.sym_of(hell.values.dtype)
        A = ctypes.c_void_p()
        avg = max(1, int(round(hell.rs.mean()))) if hell.nrows else 1
        st = getattr(self.L, f"spgpuMg{s}hellCreate")(self.h, ctypes.byref(A), util.ptr(hell.values), util.ptr(hell.indices),
                                                      hell.hack_size, util.ptr(hell.hack_offsets), util.ptr(hell.rs), avg,
                                                      hell.nrows, hell.ncols, hell.base)
        assert st == 0, f"spgpuMg{s}hellCreate -> {st}"
        return A

    def hdia_matrix(self, hdia, expect=0):
        s = util.sym_of(hdia.values.dtype)
        A = ctypes.c_void_p()
        st = getattr(self.L, f"spgpuMg{s}hdiaCreate")(self.h, ctypes.byref(A), util.ptr(hdia.values), util.ptr(hdia.offsets),
                                                      hdia.hack_size, util.ptr(hdia.hack_offsets), hdia.nrows, hdia.ncols)
        assert st == expect, f"spgpuMg{s}hdiaCreate -> {st}"
        return A

    def vector(self, A, values=None):
        v = ctypes.c_void_p()
        assert self.L.spgpuMgVectorCreate(A, ctypes.byref(v)) == 0
        if values is not None:
            assert self.L.spgpuMgVectorSet(v, util.ptr(np.ascontiguousarray(values))) == 0
        return v

    def get(self, v, n, dtype):
        out = np.zeros(n, dtype=dtype)
        assert self.L.spgpuMgVectorGet(v, util.ptr(out)) == 0
        return out

    def close(self):
        self.L.spgpuMgDestroy(self.h)


def _device_lists():
    import torch
    lists = [([0], "1 rank"), ([0, 0], "2 ranks on one device (events)"), ([0, 0, 0], "3 ranks on one device (events)")]
    if torch.cuda.device_count() >= 2:
        lists.append(([0, 1], "2 devices (fused)"))
    return lists


@pytest.mark.parametrize("dtype", DTYPES)
def test_partitioned_spmv_equals_the_single_gpu_product(ours, gpu_handle, dtype):
    """z = alpha A x + beta y through spgpuMg?hellspmv on 1 / 2 / 3 ranks == spgpu?hellspmv on the whole matrix
    (bit for bit: the same kernels multiply the same rows in the same order) == the CPU oracle within tolerance"""
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    n = 20
    coo = G.laplace3d_7pt(n)
    vals = coo.vals.astype(dtype)
    if t.is_complex:
        vals = (vals + 0.25j * np.random.default_rng(1).standard_normal(vals.shape[0])).astype(dtype)
    coo = F.Coo(coo.rows, coo.cols, vals, coo.nrows, coo.ncols, coo.base)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    x = G.random_vector(coo.nrows, dtype, 1, -1, 1)
    y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    single = util.dev_spmv(ours, gpu_handle, "hell", hell, util.upload(hell), x, y, alpha, beta, avg=7)
    want = util.oracle_spmv("hell", hell, x, y, alpha, beta)
    util.assert_rows_close(single, want, util.row_scale(coo, x, y, alpha, beta), s, "single-GPU product")
    for devices, what in _device_lists():
        mg = Mg(ours, devices)
        try:
            A = mg.matrix(hell)
            halo = ours.spgpuMgMatrixHalo(A)
            assert halo == (0 if len(devices) == 1 else -(-(n * n) // 32) * 32), (what, halo)
            vx, vy, vz = mg.vector(A, x), mg.vector(A, y), mg.vector(A)
            for _ in range(3):                               # three exchanges: both zone pairs of the fused protocol
                st = getattr(ours, f"spgpuMg{s}hellspmv")(mg.h, vz, vy, t.scalar(alpha), A, vx, t.scalar(beta))
                assert st == 0
            assert ours.spgpuMgSynchronize(mg.h) == 0
            got = mg.get(vz, coo.nrows, dtype)
            assert np.array_equal(got.view(np.uint8), single.view(np.uint8)), what
            # in place on y, beta != 0
            st = getattr(ours, f"spgpuMg{s}hellspmv")(mg.h, vy, vy, t.scalar(alpha), A, vx, t.scalar(beta))
            assert st == 0 and ours.spgpuMgSynchronize(mg.h) == 0
            assert np.array_equal(mg.get(vy, coo.nrows, dtype).view(np.uint8), single.view(np.uint8)), what
            for v in (vx, vy, vz):
                ours.spgpuMgVectorDestroy(v)
            ours.spgpuMgMatrixDestroy(A)
        finally:
            mg.close()


@pytest.mark.parametrize("dtype", [np.float64, np.complex64])
def test_unstructured_matrix_goes_through_the_all_gather_mode(ours, gpu_handle, dtype):
    """random columns reach anywhere: the matrix is kept with global columns and every rank gathers the whole x"""
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    nrows = 1000
    coo = G.random_coo(nrows, nrows, (0, 12), 3, dtype, 1)             # base 1, some empty rows
    hell = F.ell_to_hell(F.coo_to_ell(coo, ell_base=1), 32)
    assert hell.base == 1
    x = G.random_vector(nrows, dtype, 1, -1, 1)
    single = util.dev_spmv(ours, gpu_handle, "hell", hell, util.upload(hell), x, None, 1.0, 0.0)
    for devices, what in _device_lists()[1:]:
        mg = Mg(ours, devices)
        try:
            A = mg.matrix(hell)
            assert ours.spgpuMgMatrixHalo(A) == -1
            vx, vz = mg.vector(A, x), mg.vector(A)
            for _ in range(2):
                assert getattr(ours, f"spgpuMg{s}hellspmv")(mg.h, vz, None, t.scalar(1.0), A, vx, t.scalar(0.0)) == 0
            assert ours.spgpuMgSynchronize(mg.h) == 0
            assert np.array_equal(mg.get(vz, nrows, dtype).view(np.uint8), single.view(np.uint8)), what
            ours.spgpuMgVectorDestroy(vx); ours.spgpuMgVectorDestroy(vz)
            ours.spgpuMgMatrixDestroy(A)
        finally:
            mg.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_partitioned_blas1(ours, dtype):
    """spgpuMg?dot (unconjugated, like spgpu?dot), nrm2, axpby over a 3-rank partition"""
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    coo = G.laplace3d_7pt(12)
    hell = F.ell_to_hell(F.coo_to_ell(F.Coo(coo.rows, coo.cols, coo.vals.astype(dtype), coo.nrows, coo.ncols, coo.base)), 32)
    n = coo.nrows
    a = G.random_vector(n, dtype, 5, -1, 1)
    b = G.random_vector(n, dtype, 6, -1, 1)
    alpha, beta = scalars(dtype)
    tol = util.TOL[s] * 8
    mg = Mg(ours, [0, 0, 0])
    try:
        A = mg.matrix(hell)
        va, vb, vz = mg.vector(A, a), mg.vector(A, b), mg.vector(A)
        res = t.ctype()
        assert getattr(ours, f"spgpuMg{s}dot")(mg.h, ctypes.byref(res), va, vb) == 0
        ref = np.sum(a.astype(np.complex128) * b.astype(np.complex128))
        got = complex(res.x, res.y) if t.is_complex else complex(res.value)
        assert abs(got - ref) <= tol * float(np.sum(np.abs(a) * np.abs(b)))
        nr = t.rtype()
        assert getattr(ours, f"spgpuMg{s}nrm2")(mg.h, ctypes.byref(nr), va) == 0
        assert abs(nr.value - np.linalg.norm(a.astype(np.complex128))) <= tol * np.linalg.norm(a.astype(np.complex128))
        assert getattr(ours, f"spgpuMg{s}axpby")(mg.h, vz, t.scalar(beta), vb, t.scalar(alpha), va) == 0
        assert ours.spgpuMgSynchronize(mg.h) == 0
        np.testing.assert_allclose(mg.get(vz, n, dtype), beta * b + alpha * a, atol=tol * 4, rtol=0)
        for v in (va, vb, vz):
            ours.spgpuMgVectorDestroy(v)
        ours.spgpuMgMatrixDestroy(A)
    finally:
        mg.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_partitioned_hdia_spmv_equals_the_single_gpu_product(ours, gpu_handle, dtype):
    """the same for a matrix handed over in HDIA (spgpuMg?hdiaCreate: hacks re-based, diagonal offsets raised by the
    halo width, cells outside the matrix stored as 0): spgpuMg?hdiaspmv and the format-agnostic spgpuMg?spmv on
    1 / 2 / 3 ranks == spgpu?hdiaspmv on the whole matrix bit for bit == the CPU oracle within tolerance"""
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    n = 16
    coo = G.stencil3d_27pt(n)
    vals = coo.vals.astype(dtype)
    if t.is_complex:
        vals = (vals + 0.25j * np.random.default_rng(1).standard_normal(vals.shape[0])).astype(dtype)
    coo = F.Coo(coo.rows, coo.cols, vals, coo.nrows, coo.ncols, coo.base)
    hdia = F.coo_to_hdia(coo, 32)
    x = G.random_vector(coo.nrows, dtype, 1, -1, 1)
    y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    single = util.dev_spmv(ours, gpu_handle, "hdia", hdia, util.upload(hdia), x, y, alpha, beta)
    want = util.oracle_spmv("hdia", hdia, x, y, alpha, beta)
    util.assert_rows_close(single, want, util.row_scale(coo, x, y, alpha, beta), s, "single-GPU product")
    for devices, what in _device_lists():
        mg = Mg(ours, devices)
        try:
            A = mg.hdia_matrix(hdia)
            halo = ours.spgpuMgMatrixHalo(A)
            # the furthest cell OUTSIDE a block: at most the widest diagonal n*n + n + 1, less where the block boundary
            # is a plane boundary (the rows next to it have no x+1 / y+1 neighbours)
            assert (halo == 0) if len(devices) == 1 else (halo % 32 == 0 and n * n <= halo <= -(-(n * n + n + 1) // 32) * 32), (what, halo)
            vx, vy, vz = mg.vector(A, x), mg.vector(A, y), mg.vector(A)
            for k in range(3):
                op = "hdiaspmv" if k < 2 else "spmv"
                st = getattr(ours, f"spgpuMg{s}{op}")(mg.h, vz, vy, t.scalar(alpha), A, vx, t.scalar(beta))
                assert st == 0
            assert ours.spgpuMgSynchronize(mg.h) == 0
            got = mg.get(vz, coo.nrows, dtype)
            assert np.array_equal(got.view(np.uint8), single.view(np.uint8)), what
            # the HELL entry point refuses an HDIA matrix
            st = getattr(ours, f"spgpuMg{s}hellspmv")(mg.h, vz, vy, t.scalar(alpha), A, vx, t.scalar(beta))
            assert st == capi.SPGPU_UNSUPPORTED
            for v in (vx, vy, vz):
                ours.spgpuMgVectorDestroy(v)
            ours.spgpuMgMatrixDestroy(A)
        finally:
            mg.close()


def test_hdia_that_does_not_fit_a_neighbouring_block_is_refused(ours):
    """16 blocks of 32 rows cannot feed a halo of a whole 64-row plane (27-point stencil on 8^3): SPGPU_UNSUPPORTED, no
    matrix; 8 blocks of one plane each are fine (their rows reach exactly one block away)"""
    hdia = F.coo_to_hdia(G.stencil3d_27pt(8), 32)
    mg = Mg(ours, [0] * 8)
    try:
        A = mg.hdia_matrix(hdia)
        assert ours.spgpuMgMatrixHalo(A) == 64
        ours.spgpuMgMatrixDestroy(A)
    finally:
        mg.close()
    mg = Mg(ours, [0] * 16)
    try:
        A = mg.hdia_matrix(hdia, expect=capi.SPGPU_UNSUPPORTED)
        assert not A.value
    finally:
        mg.close()


@pytest.mark.parametrize("fmt", ["hell", "hdia"])
def test_partitioned_cg(ours, oracle, fmt):
    """spgpuMgDcgStart / spgpuMgDcgStep on 1 rank (device flavour: fused SpMV + dot, device scalars) and on 3 ranks
    of one device (the blocking recurrence of the EVENTS mode): same residual history as the CPU recurrence"""
    coo = G.laplace3d_7pt(14)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    hdia = F.coo_to_hdia(coo, 32)
    n = coo.nrows
    b = G.random_vector(n, np.float64, 9)
    x = np.zeros(n); r = b.copy(); p = b.copy(); rr = float(r @ r)
    ref = []
    for _ in range(30):
        ap = util.oracle_spmv("hell", hell, p, None, 1.0, 0.0)
        a = rr / float(p @ ap)
        x += a * p; r -= a * ap
        rrn = float(r @ r)
        p = r + (rrn / rr) * p
        rr = rrn
        ref.append(rr)
    for devices, what in _device_lists():
        mg = Mg(ours, devices)
        try:
            A = mg.matrix(hell) if fmt == "hell" else mg.hdia_matrix(hdia)
            vb = mg.vector(A, b)
            cg = ctypes.c_void_p()
            assert ours.spgpuMgDcgCreate(A, ctypes.byref(cg)) == 0
            rr0 = ctypes.c_double()
            assert ours.spgpuMgDcgStart(cg, vb, ctypes.byref(rr0)) == 0
            assert abs(rr0.value - float(b @ b)) <= 1e-12 * float(b @ b)
            hist = []
            for _ in range(15):                                # one iteration per call, residual read back
                out = ctypes.c_double()
                assert ours.spgpuMgDcgStep(cg, 1, ctypes.byref(out)) == 0
                hist.append(out.value)
            assert ours.spgpuMgDcgStep(cg, 14, None) == 0      # 14 iterations without touching the host
            out = ctypes.c_double()
            assert ours.spgpuMgDcgStep(cg, 1, ctypes.byref(out)) == 0
            np.testing.assert_allclose(hist, ref[:15], rtol=1e-8, err_msg=what)
            assert abs(out.value - ref[29]) <= 1e-7 * ref[29], what
            sol = mg.get(ours.spgpuMgDcgSolution(cg), n, np.float64)
            np.testing.assert_allclose(sol, x, rtol=1e-8, atol=1e-11, err_msg=what)
            ours.spgpuMgDcgDestroy(cg)
            ours.spgpuMgVectorDestroy(vb)
            ours.spgpuMgMatrixDestroy(A)
        finally:
            mg.close()


def test_blocks_handed_over_pre_partitioned(ours, gpu_handle):
    """spgpuMgHellCreateFromBlocks: per-rank blocks with LOCAL columns (what a caller that assembles its own slab
    hands over; BASELINE configs[4] is built this way), host arrays and device arrays"""
    import torch
    from spgpu_b200 import mg as pymg
    n, world = 16, 2
    coo = G.laplace3d_7pt(n)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    halo = n * n
    x = G.random_vector(coo.nrows, np.float64, 4, -1, 1)
    single = util.dev_spmv(ours, gpu_handle, "hell", hell, util.upload(hell), x, None, 1.0, 0.0, avg=7)
    blocks = [pymg.split_hell(hell, world, r, halo) for r in range(world)]
    T = util.TYPES["D"]
    for on_device in (0, 1):
        mg = Mg(ours, [0] * world)
        try:
            keep = []
            def arr(a):
                if not on_device:
                    return util.ptr(a)
                t = util.to_dev(a); keep.append(t)
                return t.data_ptr()
            PP = ctypes.c_void_p * world
            rows = (ctypes.c_int * world)(*[b.nrows for b in blocks])
            elements = (ctypes.c_longlong * world)(*[b.values.shape[0] for b in blocks])
            A = ctypes.c_void_p()
            st = ours.spgpuMgHellCreateFromBlocks(mg.h, ctypes.byref(A), capi.SPGPU_TYPE_DOUBLE, 32, halo, 0, 7, rows,
                                                  PP(*[arr(b.values) for b in blocks]), PP(*[arr(b.indices) for b in blocks]),
                                                  PP(*[arr(b.hack_offsets) for b in blocks]), PP(*[arr(b.rs) for b in blocks]),
                                                  elements, on_device)
            assert st == 0
            lo, hi = ctypes.c_int(), ctypes.c_int()
            ours.spgpuMgMatrixRowBlock(A, 1, ctypes.byref(lo), ctypes.byref(hi))
            assert (lo.value, hi.value) == (blocks[1].lo, blocks[1].hi) and ours.spgpuMgMatrixRows(A) == coo.nrows
            vx, vz = mg.vector(A, x), mg.vector(A)
            assert ours.spgpuMgDhellspmv(mg.h, vz, None, T.scalar(1.0), A, vx, T.scalar(0.0)) == 0
            assert ours.spgpuMgSynchronize(mg.h) == 0
            assert np.array_equal(mg.get(vz, coo.nrows, np.float64), single)
            ours.spgpuMgVectorDestroy(vx); ours.spgpuMgVectorDestroy(vz)
            ours.spgpuMgMatrixDestroy(A)
            torch.cuda.synchronize()
        finally:
            mg.close()


@pytest.mark.parametrize("nrows", [0, 40, 64])
def test_fewer_hacks_than_ranks_and_the_empty_matrix(ours, gpu_handle, nrows):
    """3 ranks, a matrix of 0 / 2 / 2 hacks: some ranks own no rows at all (a zone cannot be fed from an empty
    block, so the plan is all-gather); every call still succeeds and the owners' rows are right"""
    coo = G.laplace2d_5pt(8, nrows // 8) if nrows else F.Coo(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), 0, 0, 0)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    x = G.random_vector(nrows, np.float64, 1, -1, 1)
    T = util.TYPES["D"]
    single = util.dev_spmv(ours, gpu_handle, "hell", hell, util.upload(hell), x, None, 1.0, 0.0) if nrows else np.zeros(0)
    mg = Mg(ours, [0, 0, 0])
    try:
        A = mg.matrix(hell)
        assert ours.spgpuMgMatrixRows(A) == nrows
        assert ours.spgpuMgMatrixHalo(A) == -1
        lo, hi = ctypes.c_int(), ctypes.c_int()
        owned = []
        for r in range(3):
            ours.spgpuMgMatrixRowBlock(A, r, ctypes.byref(lo), ctypes.byref(hi))
            owned.append(hi.value - lo.value)
        assert sum(owned) == nrows and min(owned) == 0
        vx, vz = mg.vector(A, x), mg.vector(A)
        for _ in range(2):
            assert ours.spgpuMgDhellspmv(mg.h, vz, None, T.scalar(1.0), A, vx, T.scalar(0.0)) == 0
        assert ours.spgpuMgSynchronize(mg.h) == 0
        assert np.array_equal(mg.get(vz, nrows, np.float64), single)
        res = ctypes.c_double(-1.0)
        assert ours.spgpuMgDdot(mg.h, ctypes.byref(res), vx, vx) == 0
        assert abs(res.value - float(x @ x)) <= 1e-12 * max(1.0, float(x @ x))
        ours.spgpuMgVectorDestroy(vx); ours.spgpuMgVectorDestroy(vz)
        ours.spgpuMgMatrixDestroy(A)
    finally:
        mg.close()


def test_argument_checks(ours):
    mg = Mg(ours, [0, 0])
    try:
        coo = G.laplace3d_7pt(8)
        hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
        A = mg.matrix(hell)
        vx = mg.vector(A)
        T = util.TYPES["D"]
        assert ours.spgpuMgDhellspmv(mg.h, vx, None, T.scalar(1.0), A, vx, T.scalar(0.0)) == capi.SPGPU_UNSUPPORTED   # z == x
        assert ours.spgpuMgShellspmv(mg.h, vx, None, 1.0, A, vx, 0.0) == capi.SPGPU_UNSUPPORTED                       # wrong type
        assert ours.spgpuMgSetExchange(mg.h, capi.MG_FUSED) == capi.SPGPU_UNSUPPORTED      # the same device twice cannot spin on itself
        assert ours.spgpuMgExchange(mg.h) == capi.MG_EVENTS
        bad = ctypes.c_void_p()
        st = ours.spgpuMgDhellCreate(mg.h, ctypes.byref(bad), util.ptr(hell.values), util.ptr(hell.indices), 32,
                                     util.ptr(hell.hack_offsets), util.ptr(hell.rs), 7, hell.nrows, hell.nrows + 1, 0)
        assert st == capi.SPGPU_UNSUPPORTED                                                # not square
        ours.spgpuMgVectorDestroy(vx)
        ours.spgpuMgMatrixDestroy(A)
    finally:
        mg.close()
    h = ctypes.c_void_p()
    assert ours.spgpuMgCreate(ctypes.byref(h), (ctypes.c_int * 1)(0), 0) == capi.SPGPU_UNSUPPORTED


@pytest.mark.parametrize("hack,base,world", [(64, 1, 2), (32, 1, 3), (96, 0, 2)])
def test_partitioned_spmv_generic_hack_sizes_and_index_base(ours, gpu_handle, hack, base, world):
    """a random banded matrix (rows reach at most 100 columns away; ragged rows, some empty; the row count is not a
    multiple of 128 or of the hack size) with 1-based indices and hack sizes other than 32: the split keeps hack
    boundaries, the remap keeps the base, the halo width is the measured reach rounded up to 32"""
    rng = np.random.default_rng(hack + base)
    n = 5000 + 37
    lens = rng.integers(0, 12, n)
    rows = np.repeat(np.arange(n), lens)
    cols = np.clip(rows + rng.integers(-100, 101, rows.shape[0]), 0, n - 1)
    key = np.unique(rows.astype(np.int64) * n + cols)
    rows, cols = (key // n).astype(np.int32), (key % n).astype(np.int32)
    vals = rng.uniform(-1, 1, rows.shape[0])
    coo = F.Coo(rows + base, cols + base, vals, n, n, base)
    hell = F.ell_to_hell(F.coo_to_ell(coo, ell_base=base), hack)
    assert hell.base == base
    x = G.random_vector(n, np.float64, 1, -1, 1)
    y = G.random_vector(n, np.float64, 2, -1, 1)
    T = util.TYPES["D"]
    single = util.dev_spmv(ours, gpu_handle, "hell", hell, util.upload(hell), x, y, 1.5, -0.25)
    want = util.oracle_spmv("hell", hell, x, y, 1.5, -0.25)
    util.assert_rows_close(single, want, util.row_scale(coo, x, y, 1.5, -0.25), "D", "single-GPU product")
    import torch
    lists = [[0] * world]
    if torch.cuda.device_count() >= world:
        lists.append(list(range(world)))
    for devices in lists:
        mg = Mg(ours, devices)
        try:
            A = mg.matrix(hell)
            halo = ours.spgpuMgMatrixHalo(A)
            assert 0 < halo <= 128 and halo % 32 == 0, halo
            lo, hi = ctypes.c_int(), ctypes.c_int()
            for r in range(world):
                ours.spgpuMgMatrixRowBlock(A, r, ctypes.byref(lo), ctypes.byref(hi))
                assert lo.value % hack == 0 and (hi.value % hack == 0 or hi.value == n)
            vx, vy, vz = mg.vector(A, x), mg.vector(A, y), mg.vector(A)
            for _ in range(3):
                assert ours.spgpuMgDhellspmv(mg.h, vz, vy, T.scalar(1.5), A, vx, T.scalar(-0.25)) == 0
            assert ours.spgpuMgSynchronize(mg.h) == 0
            got = mg.get(vz, n, np.float64)
            assert np.array_equal(got.view(np.uint8), single.view(np.uint8)), devices
            for v in (vx, vy, vz):
                ours.spgpuMgVectorDestroy(v)
            ours.spgpuMgMatrixDestroy(A)
        finally:
            mg.close()
