"""GPU parity of the BLAS-1 companions (axpby, scal, dot, nrm2, amax, asum,
gather/scatter, + the element-wise helpers) against the CPU oracle, through the
C ABI.  Includes the reference's own known-answer tests
(testSparseVector.c:47-125 exact; testDenseVector.c:51-76)."""
import numpy as np
import pytest

from spgpu_b200 import generators as G
from tests import util

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64, np.complex64, np.complex128]
SIZES = [1, 3, 255, 1234, 100003, 1 << 20]


def scalars(dtype):
    if np.dtype(dtype).kind == "c":
        return (0.7 - 0.3j), (-0.5 + 0.25j)
    return 1.25, -0.75


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("offset", [0, 1])       # 1: pointers not 16-byte aligned -> scalar path
def test_axpby_scal(ours, oracle, gpu_handle, dtype, n, offset):
    import torch
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    alpha, beta = scalars(dtype)
    x = G.random_vector(n + offset, dtype, 1, -1, 1)
    y = G.random_vector(n + offset, dtype, 2, -1, 1)
    dx, dy = util.to_dev(x)[offset:], util.to_dev(y)[offset:]
    xs, ys = x[offset:].copy(), y[offset:].copy()      # keep the host copies alive across the calls
    dz = torch.zeros_like(dx)
    getattr(ours, f"spgpu{s}axpby")(gpu_handle, dz.data_ptr(), n, t.scalar(beta), dy.data_ptr(), t.scalar(alpha), dx.data_ptr())
    want = np.zeros(n, dtype=dtype)
    getattr(oracle, f"{s}axpby")(util.ptr(want), n, t.scalar(beta), util.ptr(ys), t.scalar(alpha), util.ptr(xs))
    torch.cuda.synchronize()
    np.testing.assert_allclose(dz.cpu().numpy(), want, rtol=0, atol=util.TOL[s] * 4)
    # beta == 0: y not read
    dnan = torch.full_like(dy, float("nan"))
    getattr(ours, f"spgpu{s}axpby")(gpu_handle, dz.data_ptr(), n, t.scalar(0.0), dnan.data_ptr(), t.scalar(alpha), dx.data_ptr())
    getattr(ours, f"spgpu{s}scal")(gpu_handle, dy.data_ptr(), n, t.scalar(alpha), dx.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dz.cpu().numpy(), dy.cpu().numpy())
    getattr(oracle, f"{s}scal")(util.ptr(want), n, t.scalar(alpha), util.ptr(xs))
    np.testing.assert_allclose(dz.cpu().numpy(), want, rtol=0, atol=util.TOL[s] * 4)
    # in place: z aliases x
    getattr(ours, f"spgpu{s}axpby")(gpu_handle, dx.data_ptr(), n, t.scalar(beta), dy.data_ptr(), t.scalar(alpha), dx.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(dx.cpu().numpy(), beta * dy.cpu().numpy() + alpha * x[offset:], rtol=0, atol=util.TOL[s] * 8)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES + [5_000_011])
@pytest.mark.parametrize("offset", [0, 1])
def test_reductions(ours, oracle, gpu_handle, dtype, n, offset):
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    x = G.random_vector(n + offset, dtype, 3, -1, 1)
    y = G.random_vector(n + offset, dtype, 4, -1, 1)
    xs, ys = x[offset:].copy(), y[offset:].copy()
    dx, dy = util.to_dev(x)[offset:], util.to_dev(y)[offset:]
    eps = 2e-7 if s in "SC" else 4e-16
    got = t.from_c(getattr(ours, f"spgpu{s}dot")(gpu_handle, n, dx.data_ptr(), dy.data_ptr()))
    want = oracle.dot(s, xs, ys)
    mag = float(np.sum(np.abs(xs.astype(np.complex128)) * np.abs(ys.astype(np.complex128))))
    # per-thread chains are short (n / (grid*256) terms); allow sqrt-ish growth generously
    assert abs(got - want) <= eps * 64 * mag + 1e-30, (got, want)
    got = getattr(ours, f"spgpu{s}nrm2")(gpu_handle, n, dx.data_ptr())
    want = getattr(oracle, f"{s}nrm2")(n, util.ptr(xs))
    assert abs(got - want) <= eps * 64 * want + 1e-30
    got = getattr(ours, f"spgpu{s}amax")(gpu_handle, n, dx.data_ptr())
    want = getattr(oracle, f"{s}amax")(n, util.ptr(xs))
    assert abs(got - want) <= eps * 4 * want
    got = getattr(ours, f"spgpu{s}asum")(gpu_handle, n, dx.data_ptr())
    want = getattr(oracle, f"{s}asum")(n, util.ptr(xs))
    assert abs(got - want) <= eps * 64 * want + 1e-30
    # determinism: same n, same grid -> same bits
    again = getattr(ours, f"spgpu{s}asum")(gpu_handle, n, dx.data_ptr())
    assert again == got


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reference_dense_vector_kat(ours, gpu_handle, dtype):
    """reference testDenseVector.c:51-76: x[i] = i, n = 1234"""
    s = util.sym_of(dtype)
    n = 1234
    x = np.arange(n).astype(dtype)
    dx = util.to_dev(x)
    exact = (n - 1) * n * (2 * n - 1) // 6            # 626 037 729 -> not exact in float32
    got = getattr(ours, f"spgpu{s}dot")(gpu_handle, n, dx.data_ptr(), dx.data_ptr())
    rel = 1e-6 if s == "S" else 0.0
    assert abs(got - exact) <= rel * exact
    got = getattr(ours, f"spgpu{s}nrm2")(gpu_handle, n, dx.data_ptr())
    assert abs(got - np.sqrt(float(exact))) <= (1e-6 if s == "S" else 1e-15) * np.sqrt(float(exact))


@pytest.mark.parametrize("sym,dtype", [("S", np.float32), ("D", np.float64), ("I", np.int32),
                                       ("C", np.complex64), ("Z", np.complex128)])
def test_reference_sparse_vector_kat(ours, oracle, gpu_handle, sym, dtype):
    """reference testSparseVector.c:47-125, exact equality"""
    import torch
    t = util.TYPES[sym]
    n, m = 1234, 123
    x = np.arange(n).astype(dtype)
    idx = ((np.arange(m) * 17) % n).astype(np.int32)
    if sym == "I":
        vals = (m - np.arange(m)).astype(dtype)
    else:
        vals = (np.float32(1.111) * (m - np.arange(m)).astype(np.float32)).astype(dtype)
    want = x.copy()
    getattr(oracle, f"{sym}scat")(util.ptr(want), m, util.ptr(vals), util.ptr(idx), 0, t.scalar(2))
    dx, dv, di = util.to_dev(x), util.to_dev(vals), util.to_dev(idx)
    getattr(ours, f"spgpu{sym}scat")(gpu_handle, dx.data_ptr(), m, dv.data_ptr(), di.data_ptr(), 0, t.scalar(2))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dx.cpu().numpy(), want)
    getattr(ours, f"spgpu{sym}gath")(gpu_handle, dv.data_ptr(), m, di.data_ptr(), 0, dx.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dv.cpu().numpy(), want[idx])


def test_gather_scatter_base_and_negative_positions(ours, oracle, gpu_handle):
    """entries whose position idx-base is negative are skipped (gath_base.cuh:40-43)"""
    import torch
    n, m = 5000, 2000
    rng = np.random.default_rng(0)
    y = rng.standard_normal(n)
    idx = rng.permutation(n)[:m].astype(np.int32) + 1       # base 1
    idx[::7] = 0                                             # position -1 -> skipped
    vals = rng.standard_normal(m)
    for beta in (0.0, 2.5):
        want = y.copy()
        oracle.Dscat(util.ptr(want), m, util.ptr(vals), util.ptr(idx), 1, util.TYPES["D"].scalar(beta))
        dy, dv, di = util.to_dev(y), util.to_dev(vals), util.to_dev(idx)
        ours.spgpuDscat(gpu_handle, dy.data_ptr(), m, dv.data_ptr(), di.data_ptr(), 1, util.TYPES["D"].scalar(beta))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(dy.cpu().numpy(), want)
    out = np.full(m, -7.0)
    dout = util.to_dev(out)
    ours.spgpuDgath(gpu_handle, dout.data_ptr(), m, di.data_ptr(), 1, dy.data_ptr())
    oracle.Dgath(util.ptr(out), m, util.ptr(idx), 1, util.ptr(want))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dout.cpu().numpy(), out)


@pytest.mark.parametrize("dtype", DTYPES)
def test_multivector_and_elementwise_helpers(ours, gpu_handle, dtype):
    import torch
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    n, count, pitch = 1000, 3, 1024
    alpha, beta = scalars(dtype)
    x = G.random_vector(count * pitch, dtype, 5, -1, 1)
    y = G.random_vector(count * pitch, dtype, 6, -1, 1)
    z = G.random_vector(count * pitch, dtype, 7, -1, 1)
    dx, dy, dz = util.to_dev(x), util.to_dev(y), util.to_dev(z)
    tol = util.TOL[s] * 8
    X, Y, Z = (a.reshape(count, pitch)[:, :n] for a in (x, y, z))
    # maxpby
    dw = torch.zeros_like(dx)
    getattr(ours, f"spgpu{s}maxpby")(gpu_handle, dw.data_ptr(), n, t.scalar(beta), dy.data_ptr(), t.scalar(alpha), dx.data_ptr(), count, pitch)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy().reshape(count, pitch)[:, :n], beta * Y + alpha * X, atol=tol, rtol=0)
    # mdot / mnrm2 / mamax / masum (host outputs)
    out = np.zeros(count, dtype=dtype)
    getattr(ours, f"spgpu{s}mdot")(gpu_handle, util.ptr(out), n, dx.data_ptr(), dy.data_ptr(), count, pitch)
    np.testing.assert_allclose(out, (X.astype(np.complex128) * Y).sum(1), rtol=1e-4 if s in "SC" else 1e-12)
    outr = np.zeros(count, dtype=util.real_of(dtype))
    getattr(ours, f"spgpu{s}mnrm2")(gpu_handle, util.ptr(outr), n, dx.data_ptr(), count, pitch)
    np.testing.assert_allclose(outr, np.linalg.norm(X.astype(np.complex128), axis=1), rtol=1e-5 if s in "SC" else 1e-13)
    getattr(ours, f"spgpu{s}mamax")(gpu_handle, util.ptr(outr), n, dx.data_ptr(), count, pitch)
    np.testing.assert_allclose(outr, np.abs(X).max(1), rtol=1e-6 if s in "SC" else 1e-14)
    getattr(ours, f"spgpu{s}masum")(gpu_handle, util.ptr(outr), n, dx.data_ptr(), count, pitch)
    np.testing.assert_allclose(outr, np.abs(X.astype(np.complex128)).sum(1), rtol=1e-5 if s in "SC" else 1e-13)
    # axy, axypbz, abs, setscal
    getattr(ours, f"spgpu{s}axy")(gpu_handle, dw.data_ptr(), n, t.scalar(alpha), dx.data_ptr(), dy.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy()[:n], alpha * x[:n] * y[:n], atol=tol, rtol=0)
    getattr(ours, f"spgpu{s}axypbz")(gpu_handle, dw.data_ptr(), n, t.scalar(beta), dz.data_ptr(), t.scalar(alpha), dx.data_ptr(), dy.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy()[:n], beta * z[:n] + alpha * x[:n] * y[:n], atol=tol, rtol=0)
    getattr(ours, f"spgpu{s}abs")(gpu_handle, dw.data_ptr(), n, t.scalar(alpha), dx.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy()[:n], alpha * np.abs(x[:n]), atol=tol, rtol=0)
    getattr(ours, f"spgpu{s}setscal")(gpu_handle, 11, 20, 1, t.scalar(alpha), dw.data_ptr())
    torch.cuda.synchronize()
    w = dw.cpu().numpy()
    assert (w[10:20] == np.asarray(alpha, dtype=dtype)).all() and w[9] != np.asarray(alpha, dtype=dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.complex64])
def test_multivector_reductions_are_one_launch(ours, gpu_handle, dtype):
    """spgpu?m{dot,nrm2,amax,asum}: the reference loops the blocking scalar routine `count` times (count launches +
    count host synchronisations, reference ddot.cu:152-160); here a batch is ONE launch (grid.y = vector) whatever
    `count` is -- also for odd pitches (unaligned vectors), and in slices of 65535 for huge batches -- and every
    vector's result equals the scalar routine's bit for bit."""
    s = util.sym_of(dtype)
    for n, count, pitch in ((4097, 12, 4099), (5, 70001, 7), (1 << 16, 37, 1 << 16)):
        x = G.random_vector(count * pitch, dtype, 3, -1, 1)
        y = G.random_vector(count * pitch, dtype, 4, -1, 1)
        dx, dy = util.to_dev(x), util.to_dev(y)
        expect_launches = -(-count // 65535)
        out = np.zeros(count, dtype=dtype)
        before = ours.spgpuGetLaunchCount(gpu_handle)
        getattr(ours, f"spgpu{s}mdot")(gpu_handle, util.ptr(out), n, dx.data_ptr(), dy.data_ptr(), count, pitch)
        assert ours.spgpuGetLaunchCount(gpu_handle) - before == expect_launches
        X, Y = (a.reshape(count, pitch)[:, :n] for a in (x, y))
        np.testing.assert_allclose(out, (X.astype(np.complex128) * Y).sum(1), rtol=0,
                                   atol=(1e-4 if s in "SC" else 1e-12) * np.abs(X.astype(np.complex128) * Y).sum(1).max())
        outr = np.zeros(count, dtype=util.real_of(dtype))
        for op, ref in (("nrm2", np.linalg.norm(X.astype(np.complex128), axis=1)), ("amax", np.abs(X).max(1)),
                        ("asum", np.abs(X.astype(np.complex128)).sum(1))):
            before = ours.spgpuGetLaunchCount(gpu_handle)
            getattr(ours, f"spgpu{s}m{op}")(gpu_handle, util.ptr(outr), n, dx.data_ptr(), count, pitch)
            assert ours.spgpuGetLaunchCount(gpu_handle) - before == expect_launches
            np.testing.assert_allclose(outr, ref, rtol=1e-5 if s in "SC" else 1e-13)
        # a few vectors against the scalar entry points
        isz = np.dtype(dtype).itemsize
        for v in (0, count // 2, count - 1):
            one = getattr(ours, f"spgpu{s}nrm2")(gpu_handle, n, dx.data_ptr() + v * pitch * isz)
            getattr(ours, f"spgpu{s}mnrm2")(gpu_handle, util.ptr(outr), n, dx.data_ptr(), count, pitch)
            if count <= 148 * 4:       # same grid shape per vector only when the batch does not shrink the per-vector grid
                assert abs(outr[v] - one) <= 1e-6 * abs(one)


def test_set_stream_with_work_in_flight_orders_the_handle_scratch(ours, gpu_handle):
    """The handle's reduction scratch (per-CTA partials, the ticket) is reused in stream order.  spgpuSetStream to a
    second stream while a reduction is still running on the first must not let the next reduction overtake it
    (reference core.c:62-72 only swaps the pointer; its static reduction arrays had the same hazard):
    the new stream waits for an event recorded on the old one."""
    import torch
    n = 1 << 26
    a = torch.rand(n, dtype=torch.float64, device="cuda")
    b = torch.rand(1 << 20, dtype=torch.float64, device="cuda")
    want_a, want_b = float(torch.dot(a, a).item()), float(torch.dot(b, b).item())
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    try:
        for _ in range(5):
            ours.spgpuSetStream(gpu_handle, s1.cuda_stream)
            ours.spgpuDdotDev(gpu_handle, n, a.data_ptr(), a.data_ptr(), res.data_ptr())           # ~0.17 ms on s1
            ours.spgpuSetStream(gpu_handle, s2.cuda_stream)
            ours.spgpuDdotDev(gpu_handle, b.numel(), b.data_ptr(), b.data_ptr(), res.data_ptr() + 8)   # short, on s2
            torch.cuda.synchronize()
            assert abs(float(res[0].item()) - want_a) <= 1e-11 * want_a
            assert abs(float(res[1].item()) - want_b) <= 1e-11 * want_b
    finally:
        ours.spgpuSetStream(gpu_handle, None)
