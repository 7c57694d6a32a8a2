"""GPU tests of the additive Krylov entry points (include/spgpu_ext.h) and of the CG
iteration built on the C ABI (BASELINE configs[4]: SpMV + dot + axpby)."""
import ctypes

import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu


def test_device_result_reductions(ours, oracle, gpu_handle):
    import torch
    for s, dtype in (("S", np.float32), ("D", np.float64), ("C", np.complex64), ("Z", np.complex128)):
        n = 300007
        x = G.random_vector(n, dtype, 1, -1, 1)
        y = G.random_vector(n, dtype, 2, -1, 1)
        dx, dy = util.to_dev(x), util.to_dev(y)
        out = torch.zeros(2, dtype=dx.dtype, device="cuda")
        getattr(ours, f"spgpu{s}dotDev")(gpu_handle, n, dx.data_ptr(), dy.data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy()[0]
        blocking = util.TYPES[s].from_c(getattr(ours, f"spgpu{s}dot")(gpu_handle, n, dx.data_ptr(), dy.data_ptr()))
        assert got == np.asarray(blocking, dtype=dtype)            # same kernel, same order -> same bits
        want = oracle.dot(s, x, y)
        mag = float(np.sum(np.abs(x.astype(np.complex128)) * np.abs(y.astype(np.complex128))))
        assert abs(got - want) <= (2e-7 if s in "SC" else 4e-16) * 64 * mag
        rout = torch.zeros(2, dtype=torch.float32 if s in "SC" else torch.float64, device="cuda")
        getattr(ours, f"spgpu{s}nrm2sqDev")(gpu_handle, n, dx.data_ptr(), rout.data_ptr())
        torch.cuda.synchronize()
        nrm = getattr(oracle, f"{s}nrm2")(n, util.ptr(x))
        assert abs(np.sqrt(float(rout[0].item())) - nrm) <= (1e-5 if s in "SC" else 1e-13) * nrm


def test_axpby_with_device_scalars(ours, gpu_handle):
    import torch
    for n in (1, 1001, 1 << 20):
        rng = np.random.default_rng(n)
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        sc = np.array([3.0, 7.0, -2.0, 5.0])
        dx, dy, ds = util.to_dev(x), util.to_dev(y), util.to_dev(sc)
        dz = torch.zeros_like(dx)
        p = ds.data_ptr()
        # z = (sc0/sc1) * y - (sc2/sc3) * x
        ours.spgpuDaxpbyDev(gpu_handle, dz.data_ptr(), n, p, p + 8, 1.0, dy.data_ptr(), p + 16, p + 24, -1.0, dx.data_ptr())
        torch.cuda.synchronize()
        np.testing.assert_allclose(dz.cpu().numpy(), (3.0 / 7.0) * y - (-2.0 / 5.0) * x, rtol=1e-14, atol=1e-15)
        # NULL pairs mean 1; in place on y
        ours.spgpuDaxpbyDev(gpu_handle, dy.data_ptr(), n, 0, 0, 1.0, dy.data_ptr(), p, 0, 1.0, dx.data_ptr())
        torch.cuda.synchronize()
        np.testing.assert_allclose(dy.cpu().numpy(), y + 3.0 * x, rtol=1e-14, atol=1e-15)


def test_fused_spmv_dot(ours, gpu_handle):
    import torch
    coo = G.laplace3d_7pt(24)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    dA = util.upload(A)
    n = coo.nrows
    x = G.random_vector(n, np.float64, 3)
    dx = util.to_dev(x)
    dz = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    dres = torch.full((1,), float("nan"), dtype=torch.float64, device="cuda")
    ours.spgpuDhellspmvDot(gpu_handle, dz.data_ptr(), dA["values"].data_ptr(), dA["indices"].data_ptr(), 32,
                           dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), n, dx.data_ptr(), 0, 0, dres.data_ptr())
    torch.cuda.synchronize()
    want = util.oracle_spmv("hell", A, x, None, 1.0, 0.0)
    util.assert_rows_close(dz.cpu().numpy(), want, util.row_scale(coo, x, None, 1.0, 0.0), "D", "fused spmv")
    scale = float(np.sum(np.abs(x) * np.abs(want)))
    assert abs(float(dres.item()) - float(np.dot(x, want))) <= 1e-12 * scale


@pytest.mark.parametrize("flavour", ["blocking", "device", "device-pingpong"])
def test_cg_converges_like_the_cpu_recurrence(ours, gpu_handle, flavour):
    """40 CG iterations on a 3-D Laplacian: same residual history as the identical
    recurrence run with numpy + the CPU oracle, and the final x solves the system
    (device-pingpong: r.r alternates between two scalar slots instead of being copied, 4 launches per iteration)"""
    import torch
    from spgpu_b200 import krylov
    coo = G.laplace3d_7pt(16)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    dA = util.upload(A)
    n = coo.nrows
    T = util.TYPES["D"]
    b = G.random_vector(n, np.float64, 9)

    def apply_A(z, x_ext):
        ours.spgpuDhellspmv(gpu_handle, z.data_ptr(), 0, T.scalar(1.0), dA["values"].data_ptr(), dA["indices"].data_ptr(), 32,
                            dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), 0, 7, n, x_ext.data_ptr(), T.scalar(0.0), 0)

    def apply_A_dot(z, x_ext, dres, ar=None):
        ours.spgpuDhellspmvDot(gpu_handle, z.data_ptr(), dA["values"].data_ptr(), dA["indices"].data_ptr(), 32,
                               dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), n, x_ext.data_ptr(), 0, 0, dres)

    stream = torch.cuda.ExternalStream(ours.spgpuGetStream(gpu_handle))
    with torch.cuda.stream(stream):
        st = krylov.CgState(n, 0, "cuda")
        cg = krylov.Cg(ours, gpu_handle, st, apply_A, apply_A_dot, pingpong=flavour == "device-pingpong")
        rr0 = cg.start(util.to_dev(b))
        hist = []
        launches0 = ours.spgpuGetLaunchCount(gpu_handle)
        for _ in range(40):
            if flavour == "blocking":
                hist.append(cg.step_blocking())
            else:
                cg.step_device()
                hist.append(cg.residual_norm2())
        torch.cuda.synchronize()
        per_iteration = (ours.spgpuGetLaunchCount(gpu_handle) - launches0) / 40
        assert per_iteration == {"blocking": 6, "device": 5, "device-pingpong": 4}[flavour]
        x_gpu = st.x.cpu().numpy()
    # CPU recurrence with the oracle SpMV
    x = np.zeros(n); r = b.copy(); p = b.copy(); rr = float(r @ r)
    assert abs(rr0 - rr) <= 1e-12 * rr
    ref = []
    for _ in range(40):
        ap = util.oracle_spmv("hell", A, p, None, 1.0, 0.0)
        alpha = rr / float(p @ ap)
        x += alpha * p; r -= alpha * ap
        rr_new = float(r @ r)
        p = r + (rr_new / rr) * p
        rr = rr_new
        ref.append(rr)
    np.testing.assert_allclose(hist, ref, rtol=1e-8)
    assert hist[-1] < 1e-6 * rr0
    np.testing.assert_allclose(x_gpu, x, rtol=1e-9, atol=1e-12)


def test_cg_on_hdia_with_the_fused_spmv_dot(ours, gpu_handle):
    """the same device-scalar CG with the matrix in HDIA (27-point stencil, SPD): spgpuDhdiaspmvHaloDot with no
    neighbours is the fused SpMV + p.Ap; residual history equals the CPU recurrence with the oracle"""
    import torch
    from spgpu_b200 import krylov
    coo = G.stencil3d_27pt(12)
    A = F.coo_to_hdia(coo, 32)
    dv, doff, dho = (util.to_dev(t) for t in (A.values, A.offsets, A.hack_offsets))
    n = coo.nrows
    b = G.random_vector(n, np.float64, 9)
    seq = [0]

    def apply_A_dot(z, x_ext, dres, ar=None):
        seq[0] += 1
        ours.spgpuDhdiaspmvHaloDot(gpu_handle, z.data_ptr(), dv.data_ptr(), doff.data_ptr(), 32, dho.data_ptr(), n, n,
                                   x_ext.data_ptr(), 0, None, seq[0], dres, None)

    stream = torch.cuda.ExternalStream(ours.spgpuGetStream(gpu_handle))
    with torch.cuda.stream(stream):
        st = krylov.CgState(n, 0, "cuda")
        cg = krylov.Cg(ours, gpu_handle, st, None, apply_A_dot)
        rr0 = cg.start(util.to_dev(b))
        hist = []
        for _ in range(30):
            cg.step_device()
            hist.append(cg.residual_norm2())
        torch.cuda.synchronize()
        x_gpu = st.x.cpu().numpy()
    x = np.zeros(n); r = b.copy(); p = b.copy(); rr = float(r @ r)
    ref = []
    for _ in range(30):
        ap = util.oracle_spmv("hdia", A, p, None, 1.0, 0.0)
        alpha = rr / float(p @ ap)
        x += alpha * p; r -= alpha * ap
        rr_new = float(r @ r)
        p = r + (rr_new / rr) * p
        rr = rr_new
        ref.append(rr)
    np.testing.assert_allclose(hist, ref, rtol=1e-8)
    assert hist[-1] < 1e-6 * rr0
    np.testing.assert_allclose(x_gpu, x, rtol=1e-9, atol=1e-12)


def test_fused_cg_update(ours, gpu_handle):
    """x += a p ; r -= a Ap ; rr' = r.r in one pass, a = rr/pAp from device memory"""
    import torch
    for n in (1, 777, 1 << 20, (1 << 20) + 1):
        rng = np.random.default_rng(n)
        x, r, p, ap = (rng.standard_normal(n) for _ in range(4))
        sc = np.array([3.5, 1.25, 0.0])
        dx, dr, dp, dap, ds = (util.to_dev(a) for a in (x, r, p, ap, sc))
        s = ds.data_ptr()
        ours.spgpuDcgUpdateDev(gpu_handle, dx.data_ptr(), dr.data_ptr(), dp.data_ptr(), dap.data_ptr(), n, s, s + 8, s + 16, None)
        torch.cuda.synchronize()
        a = 3.5 / 1.25
        np.testing.assert_allclose(dx.cpu().numpy(), x + a * p, rtol=1e-14, atol=1e-14)
        rn = r - a * ap
        np.testing.assert_allclose(dr.cpu().numpy(), rn, rtol=1e-14, atol=1e-14)
        assert abs(float(ds[2].item()) - float(rn @ rn)) <= 1e-13 * float(rn @ rn)


def test_cuda_graph_capture_of_a_device_cg_iteration(ours, gpu_handle):
    """the non-blocking entry points neither allocate nor synchronise, so a whole device-scalar CG
    iteration can be captured in a CUDA graph and replayed (B200 guidance: graphs for launch-bound
    inner loops); the replayed iterations must follow the eagerly launched ones bit for bit"""
    import torch
    from spgpu_b200 import krylov
    coo = G.laplace3d_7pt(12)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    dA = util.upload(A)
    n = coo.nrows
    T = util.TYPES["D"]
    b = util.to_dev(G.random_vector(n, np.float64, 21))

    def apply_A(z, x_ext):
        ours.spgpuDhellspmv(gpu_handle, z.data_ptr(), 0, T.scalar(1.0), dA["values"].data_ptr(), dA["indices"].data_ptr(), 32,
                            dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), 0, 7, n, x_ext.data_ptr(), T.scalar(0.0), 0)

    def apply_A_dot(z, x_ext, dres, ar=None):
        ours.spgpuDhellspmvDot(gpu_handle, z.data_ptr(), dA["values"].data_ptr(), dA["indices"].data_ptr(), 32,
                               dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), n, x_ext.data_ptr(), 0, 0, dres)

    def run(iterations, graph):
        st = krylov.CgState(n, 0, "cuda")
        cg = krylov.Cg(ours, gpu_handle, st, apply_A, apply_A_dot)
        cg.start(b)
        cg.step_device()                                  # warm-up: sizes the handle's scratch
        if graph:
            g = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream()
            default = ours.spgpuGetStream(gpu_handle)
            torch.cuda.synchronize()
            ours.spgpuSetStream(gpu_handle, cap.cuda_stream)
            with torch.cuda.graph(g, stream=cap):
                cg.step_device()
            ours.spgpuSetStream(gpu_handle, None)
            assert ours.spgpuGetStream(gpu_handle) == default
            torch.cuda.synchronize()
            for _ in range(iterations):
                g.replay()
        else:
            for _ in range(iterations):
                cg.step_device()
        torch.cuda.synchronize()
        return st.x.cpu().numpy(), float(st.s[0].item())

    stream = torch.cuda.ExternalStream(ours.spgpuGetStream(gpu_handle))
    with torch.cuda.stream(stream):
        x_eager, rr_eager = run(6, graph=False)
    x_graph, rr_graph = run(6, graph=True)
    np.testing.assert_array_equal(x_graph, x_eager)
    assert rr_graph == rr_eager


@pytest.mark.parametrize("sym", ["S", "C", "Z"])
def test_device_scalar_twins_for_the_other_value_types(ours, gpu_handle, sym):
    """spgpu{S,C,Z}axpbyDev / cgUpdateDev / hellspmvDot / sumDev: the same kernels for float and the complex types
    (complex products are the library's unconjugated ones, reference zdot.cu:54)"""
    import torch
    t = util.TYPES[sym]
    dt = t.np_dtype
    tol = 1e-5 if sym in "SC" else 1e-12
    rng = np.random.default_rng(17)

    def rnd(n):
        v = rng.standard_normal(n)
        if t.is_complex:
            v = v + 1j * rng.standard_normal(n)
        return v.astype(dt)

    for n in (1, 1001, (1 << 18) + 3):
        x, y, r, p, ap = (rnd(n) for _ in range(5))
        sc = rnd(4)
        dx, dy, ds = util.to_dev(x), util.to_dev(y), util.to_dev(sc)
        dz = torch.zeros_like(dx)
        sp, isz = ds.data_ptr(), dt.itemsize
        getattr(ours, f"spgpu{sym}axpbyDev")(gpu_handle, dz.data_ptr(), n, sp, sp + isz, 1.0, dy.data_ptr(), sp + 2 * isz,
                                             sp + 3 * isz, -1.0, dx.data_ptr())
        torch.cuda.synchronize()
        want = (sc[0] / sc[1]) * y - (sc[2] / sc[3]) * x
        np.testing.assert_allclose(dz.cpu().numpy(), want, rtol=50 * tol, atol=50 * tol)
        # fused CG update
        sc3 = np.array([sc[0], sc[1], 0], dtype=dt)
        dxx, dr, dp, dap, ds3 = (util.to_dev(a) for a in (x, r, p, ap, sc3))
        s3 = ds3.data_ptr()
        getattr(ours, f"spgpu{sym}cgUpdateDev")(gpu_handle, dxx.data_ptr(), dr.data_ptr(), dp.data_ptr(), dap.data_ptr(), n,
                                                s3, s3 + isz, s3 + 2 * isz, None)
        torch.cuda.synchronize()
        a = sc[0] / sc[1]
        rn = r - a * ap
        np.testing.assert_allclose(dxx.cpu().numpy(), x + a * p, rtol=50 * tol, atol=50 * tol)
        np.testing.assert_allclose(dr.cpu().numpy(), rn, rtol=50 * tol, atol=50 * tol)
        rr = np.sum(rn.astype(np.complex128) * rn.astype(np.complex128))
        got = complex(ds3.cpu().numpy()[2])
        assert abs(got - rr) <= 20 * tol * float(np.sum(np.abs(rn) ** 2)), (got, rr)
        # sum
        dres = torch.zeros(1, dtype=dx.dtype, device="cuda")
        getattr(ours, f"spgpu{sym}sumDev")(gpu_handle, n, dx.data_ptr(), dres.data_ptr())
        torch.cuda.synchronize()
        assert abs(complex(dres.cpu().numpy()[0]) - np.sum(x.astype(np.complex128))) <= 20 * tol * float(np.sum(np.abs(x)))

    # fused SpMV + dot
    coo = G.laplace3d_7pt(16)
    vals = coo.vals.astype(dt)
    if t.is_complex:
        vals = (vals + 0.5j * rng.standard_normal(vals.shape[0])).astype(dt)
    coo = F.Coo(coo.rows, coo.cols, vals, coo.nrows, coo.ncols, coo.base)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    dA = util.upload(A)
    n = coo.nrows
    x = rnd(n)
    dx = util.to_dev(x)
    dz = torch.zeros(n, dtype=dx.dtype, device="cuda")
    dres = torch.zeros(1, dtype=dx.dtype, device="cuda")
    getattr(ours, f"spgpu{sym}hellspmvDot")(gpu_handle, dz.data_ptr(), dA["values"].data_ptr(), dA["indices"].data_ptr(), 32,
                                            dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), n, dx.data_ptr(), 0, 0,
                                            dres.data_ptr())
    torch.cuda.synchronize()
    want = util.oracle_spmv("hell", A, x, None, 1.0, 0.0)
    util.assert_rows_close(dz.cpu().numpy(), want, util.row_scale(coo, x, None, 1.0, 0.0), sym, f"fused spmv+dot {sym}")
    ref = np.sum(x.astype(np.complex128) * want.astype(np.complex128))
    assert abs(complex(dres.cpu().numpy()[0]) - ref) <= 20 * tol * float(np.sum(np.abs(x) * np.abs(want)))
