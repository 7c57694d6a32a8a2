"""CPU tests of the oracle (oracle/spmv_oracle.c): it must agree with a dense
numpy product on every format / type / option, and with the reference's own
known-answer tests that can be evaluated without a GPU."""
import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G
from tests import util

DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


def _case(dtype, base, seed, nrows=97, ncols=83):
    coo = G.random_coo(nrows, ncols, (0, 11), seed, dtype, base)
    x = G.random_vector(ncols, dtype, seed + 100, -1, 1)
    y = G.random_vector(nrows, dtype, seed + 200, -1, 1)
    return coo, x, y


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("base", [0, 1])
@pytest.mark.parametrize("fmt", ["ell", "hell", "dia", "hdia"])
def test_oracle_matches_dense(dtype, base, fmt):
    coo, x, y = _case(dtype, base, 3)
    s = util.sym_of(dtype)
    alpha, beta = (0.7 - 0.3j, -0.5 + 0.25j) if np.dtype(dtype).kind == "c" else (2.0, -3.0)
    want = alpha * (util.dense_of(coo).astype(np.complex128) @ x.astype(np.complex128)) + beta * y
    ell = F.coo_to_ell(coo, base)
    A = {"ell": lambda: ell, "hell": lambda: F.ell_to_hell(ell, 32),
         "dia": lambda: F.coo_to_dia(coo), "hdia": lambda: F.coo_to_hdia(coo, 32)}[fmt]()
    z = util.oracle_spmv(fmt, A, x, y, alpha, beta)
    util.assert_rows_close(z, want, util.row_scale(coo, x, y, alpha, beta), s, f"{fmt}/{s}")
    # beta == 0: y must not be read (NaN there must not leak)
    z0 = util.oracle_spmv(fmt, A, x, np.full_like(y, np.nan), alpha, 0.0)
    want0 = alpha * (util.dense_of(coo).astype(np.complex128) @ x.astype(np.complex128))
    util.assert_rows_close(z0, want0, util.row_scale(coo, x, None, alpha, 0), s, f"{fmt}/{s}/beta0")


def test_oracle_ridx_and_rs_null():
    coo, x, y = _case(np.float64, 0, 5)
    ell = F.coo_to_ell(coo)
    oell = F.ell_to_oell(ell)
    z_plain = util.oracle_spmv("ell", ell, x, y, 1.5, 0.5)
    z_ridx = util.oracle_spmv("ell", oell, x, y, 1.5, 0.5, ridx=oell.ridx)
    np.testing.assert_array_equal(z_plain, z_ridx)      # same slot order per row -> same bits
    z_nors = util.oracle_spmv("ell", ell, x, y, 1.5, 0.5, rs_null=True)   # zero padding
    np.testing.assert_array_equal(z_plain, z_nors)
    ohell = F.ell_to_hell(oell, 32)
    z_ohell = util.oracle_spmv("hell", ohell, x, y, 1.5, 0.5, ridx=oell.ridx)
    np.testing.assert_array_equal(z_plain, z_ohell)


def test_reference_ctest_ell_equals_hell(oracle):
    """reference src/tests/ctest.c:25-39,105,146: 100x100, 200 nnz (i%100, i%100),
    value 1, alpha=2, beta=-3: ELL and HELL must give the same dot(z,z)."""
    n, nnz = 100, 200
    rows = (np.arange(nnz) % n).astype(np.int32)
    coo = F.Coo(rows, rows.copy(), np.ones(nnz, np.float32), n, n, 0)
    rng = np.random.default_rng(0)
    x, y = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    ell = F.coo_to_ell(coo)
    assert ell.maxnnz == 2 and ell.pitch == 128
    z_ell = util.oracle_spmv("ell", ell, x, y, 2.0, -3.0)
    z_hell = util.oracle_spmv("hell", F.ell_to_hell(ell, 32), x, y, 2.0, -3.0)
    np.testing.assert_array_equal(z_ell, z_hell)
    np.testing.assert_allclose(z_ell, 2.0 * 2.0 * x - 3.0 * y, rtol=1e-5, atol=1e-6)
    assert oracle.dot("S", z_ell, z_ell) == oracle.dot("S", z_hell, z_hell)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reference_sparse_vector_kat(oracle, dtype):
    """reference src/tests/testSparseVector.c:47-125: n=1234, x[i]=i, 123 indices
    (17*i)%1234 with values 1.111f*(123-i); scatter(beta=2) then gather, exact."""
    n, m = 1234, 123
    s = util.sym_of(dtype)
    x = np.arange(n).astype(dtype)
    idx = ((np.arange(m) * 17) % n).astype(np.int32)
    vals = (np.float32(1.111) * (m - np.arange(m)).astype(np.float32)).astype(dtype)
    want = x.copy()
    for i in range(m):
        want[idx[i]] = dtype(2.0) * want[idx[i]] + vals[i]
    got = x.copy()
    getattr(oracle, f"{s}scat")(util.ptr(got), m, util.ptr(vals), util.ptr(idx), 0, util.TYPES[s].scalar(2.0))
    np.testing.assert_array_equal(got, want)
    out = np.zeros(m, dtype=dtype)
    getattr(oracle, f"{s}gath")(util.ptr(out), m, util.ptr(idx), 0, util.ptr(got))
    np.testing.assert_array_equal(out, want[idx])


def test_reference_dense_vector_kat(oracle):
    """reference src/tests/testDenseVector.c:51-76: x[i]=i, n=1234: dot and nrm2."""
    n = 1234
    x = np.arange(n, dtype=np.float64)
    exact = (n - 1) * n * (2 * n - 1) // 6
    assert oracle.dot("D", x, x) == float(exact)
    assert abs(oracle.Dnrm2(n, util.ptr(x)) - np.sqrt(float(exact))) <= 1e-12 * np.sqrt(float(exact))


@pytest.mark.parametrize("dtype", DTYPES)
def test_oracle_blas1(oracle, dtype):
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    n = 1001
    x = G.random_vector(n, dtype, 1, -1, 1)
    y = G.random_vector(n, dtype, 2, -1, 1)
    alpha, beta = (0.7 - 0.3j, -0.5 + 0.25j) if t.is_complex else (1.25, -0.75)
    z = np.zeros_like(x)
    getattr(oracle, f"{s}axpby")(util.ptr(z), n, t.scalar(beta), util.ptr(y), t.scalar(alpha), util.ptr(x))
    np.testing.assert_allclose(z, beta * y + alpha * x, rtol=util.TOL[s] * 10, atol=util.TOL[s])
    getattr(oracle, f"{s}scal")(util.ptr(z), n, t.scalar(alpha), util.ptr(x))
    np.testing.assert_allclose(z, alpha * x, rtol=util.TOL[s] * 10, atol=util.TOL[s])
    d = oracle.dot(s, x, y)
    np.testing.assert_allclose(d, np.sum(x.astype(np.complex128) * y.astype(np.complex128)), rtol=1e-6)  # unconjugated
    assert abs(getattr(oracle, f"{s}nrm2")(n, util.ptr(x)) - np.linalg.norm(x.astype(np.complex128))) < 1e-5
    assert abs(getattr(oracle, f"{s}amax")(n, util.ptr(x)) - np.abs(x).max()) < 1e-6
    assert abs(getattr(oracle, f"{s}asum")(n, util.ptr(x)) - np.abs(x.astype(np.complex128)).sum()) < 1e-3
