"""Shared helpers of the test-suite: the CPU oracle binding, device upload, and
the per-row tolerance north_star states (1e-5 S/C, 1e-12 D/Z, relative to
|alpha|*sum|a_ik||x_k| + |beta||y_i|)."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_void_p

import numpy as np

from spgpu_b200 import capi
from spgpu_b200.capi import TYPES, ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libspgpu_ref.so")

TOL = {"S": 1e-5, "C": 1e-5, "D": 1e-12, "Z": 1e-12}
P = c_void_p


class OracleLib:
    """ctypes view of oracle/liboracle.so (CPU restatement; host pointers)."""

    def __init__(self, path=ORACLE_PATH):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle liboracle.so`")
        self.dll = ctypes.CDLL(path)
        d = self.dll
        d.oracle_num_threads.restype = c_int
        d.oracle_set_num_threads.argtypes = [c_int]
        for s in "SDCZ":
            T = TYPES[s].ctype
            R = TYPES[s].rtype
            getattr(d, f"oracle_{s}ellspmv").argtypes = [P, P, T, P, P, c_int, c_int, P, P, c_int, c_int, c_int, P, T, c_int]
            getattr(d, f"oracle_{s}hellspmv").argtypes = [P, P, T, P, P, c_int, P, P, P, c_int, c_int, P, T, c_int]
            getattr(d, f"oracle_{s}diaspmv").argtypes = [P, P, T, P, P, c_int, c_int, c_int, c_int, P, T]
            getattr(d, f"oracle_{s}hdiaspmv").argtypes = [P, P, T, P, P, c_int, P, c_int, c_int, P, T]
            getattr(d, f"oracle_{s}axpby").argtypes = [P, c_int, T, P, T, P]
            getattr(d, f"oracle_{s}scal").argtypes = [P, c_int, T, P]
            getattr(d, f"oracle_{s}gath").argtypes = [P, c_int, P, c_int, P]
            getattr(d, f"oracle_{s}scat").argtypes = [P, c_int, P, P, c_int, T]
            getattr(d, f"oracle_{s}dot").argtypes = [c_int, P, P, P]
            for op in ("nrm2", "amax", "asum"):
                f = getattr(d, f"oracle_{s}{op}")
                f.argtypes = [c_int, P]
                f.restype = R
            for f in ("ellspmv", "hellspmv", "diaspmv", "hdiaspmv", "axpby", "scal", "gath", "scat", "dot"):
                getattr(d, f"oracle_{s}{f}").restype = None
        d.oracle_Igath.argtypes = [P, c_int, P, c_int, P]
        d.oracle_Iscat.argtypes = [P, c_int, P, P, c_int, c_int]

    def __getattr__(self, name):
        return getattr(self.dll, "oracle_" + name)

    def dot(self, s, a, b):
        out = np.zeros(2, dtype=np.float64)
        getattr(self.dll, f"oracle_{s}dot")(a.shape[0], ptr(a), ptr(b), ptr(out))
        return complex(out[0], out[1]) if TYPES[s].is_complex else float(out[0])


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _oracle = OracleLib()
    return _oracle


_ref = False


def ref_lib():
    global _ref
    if _ref is False:
        _ref = capi.SpgpuLib(REF_PATH, ext=False) if os.path.exists(REF_PATH) else None
    return _ref


MM_REF_PATH = os.path.join(os.path.dirname(REF_PATH), "libmm_ref.so")
_mm_ref = False


def mm_ref_lib():
    """The reference's own MatrixMarket reader behind oracle/mm_ref_shim.cpp, if it was built."""
    global _mm_ref
    if _mm_ref is False:
        _mm_ref = None
        if os.path.exists(MM_REF_PATH):
            d = ctypes.CDLL(MM_REF_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
            P, c_int, s = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
            d.ref_mm_properties.restype, d.ref_mm_properties.argtypes = c_int, [s, ctypes.POINTER(c_int)]
            for n in ("float", "double", "int"):
                f = getattr(d, f"ref_mm_load_{n}")
                f.restype, f.argtypes = c_int, [s, P, P, P]
            d.ref_mm_load_pattern.restype, d.ref_mm_load_pattern.argtypes = c_int, [s, P, P]
            for n in ("float", "double"):
                f = getattr(d, f"ref_mm_unfolded_size_{n}")
                f.restype, f.argtypes = c_int, [P, P, P, c_int]
                f = getattr(d, f"ref_mm_unfold_{n}")
                f.restype, f.argtypes = None, [P, P, P, P, P, P, c_int]
            d.ref_mm_load_vector_double.restype, d.ref_mm_load_vector_double.argtypes = c_int, [s, P, c_int]
            _mm_ref = d
    return _mm_ref


# ---------------------------------------------------------------- numpy side

def sym_of(dtype):
    dt = np.dtype(dtype)
    return {np.dtype(np.float32): "S", np.dtype(np.float64): "D",
            np.dtype(np.complex64): "C", np.dtype(np.complex128): "Z"}[dt]


def real_of(dtype):
    return {"S": np.float32, "D": np.float64, "C": np.float32, "Z": np.float64}[sym_of(dtype)]


def oracle_spmv(fmt, A, x, y, alpha, beta, base=None, ridx=None, rs_null=False):
    """Run the CPU oracle on a formats.{Ell,Hell,Dia,Hdia}; returns z (numpy)."""
    O = oracle_lib()
    s = sym_of(A.values.dtype)
    t = TYPES[s]
    n_out = A.nrows
    z = np.full(n_out, np.nan, dtype=t.np_dtype) if y is None else y.copy()
    yy = y if y is not None else None
    a, b = t.scalar(alpha), t.scalar(beta)
    if fmt == "ell":
        getattr(O, f"{s}ellspmv")(ptr(z), ptr(yy), a, ptr(A.values), ptr(A.indices), A.pitch, A.pitch,
                                  None if rs_null else ptr(A.rs), ptr(ridx), max(1, int(A.rs.mean()) if A.nrows else 1),
                                  A.maxnnz, A.nrows, ptr(x), b, A.base if base is None else base)
    elif fmt == "hell":
        getattr(O, f"{s}hellspmv")(ptr(z), ptr(yy), a, ptr(A.values), ptr(A.indices), A.hack_size,
                                   ptr(A.hack_offsets), ptr(A.rs), ptr(ridx), 1, A.nrows, ptr(x), b,
                                   A.base if base is None else base)
    elif fmt == "dia":
        getattr(O, f"{s}diaspmv")(ptr(z), ptr(yy), a, ptr(A.values), ptr(A.offsets), A.pitch, A.nrows,
                                  A.ncols, A.diags, ptr(x), b)
    elif fmt == "hdia":
        getattr(O, f"{s}hdiaspmv")(ptr(z), ptr(yy), a, ptr(A.values), ptr(A.offsets), A.hack_size,
                                   ptr(A.hack_offsets), A.nrows, A.ncols, ptr(x), b)
    else:
        raise ValueError(fmt)
    return z


def dense_of(coo):
    d = np.zeros((coo.nrows, coo.ncols), dtype=coo.vals.dtype)
    np.add.at(d, (coo.rows - coo.base, coo.cols - coo.base), coo.vals)
    return d


def row_scale(coo, x, y, alpha, beta):
    """|alpha| * sum_k |a_ik||x_k| + |beta||y_i| per row (the tolerance scale)."""
    r = np.zeros(coo.nrows, dtype=np.float64)
    np.add.at(r, coo.rows - coo.base, np.abs(coo.vals).astype(np.float64) * np.abs(x[coo.cols - coo.base]).astype(np.float64))
    r *= abs(alpha)
    if y is not None and beta != 0:
        r += abs(beta) * np.abs(y).astype(np.float64)
    return r


def assert_rows_close(z, z_ref, scale, sym, what=""):
    err = np.abs(z.astype(np.complex128) - z_ref.astype(np.complex128))
    bound = TOL[sym] * scale + np.finfo(np.float64).tiny
    bad = np.nonzero(~(err <= bound))[0]
    assert bad.size == 0, (f"{what}: {bad.size} rows out of tolerance {TOL[sym]:g}; first {bad[:5]}, "
                           f"err {err[bad[:5]]}, bound {bound[bad[:5]]}, got {z[bad[:5]]}, want {z_ref[bad[:5]]}")


# ---------------------------------------------------------------- device side

def to_dev(a):
    import torch
    if a is None:
        return None
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dptr(t):
    return 0 if t is None else t.data_ptr()


def dev_spmv_ptr(L, h, fmt, A, dA, dz, dx, dy, alpha, beta, base=None, dr=None, rs_null=False, avg=None):
    """z = alpha A x + beta y on DEVICE tensors through the C ABI of library L, asynchronous on the handle's stream
    (dy / dr may be None)."""
    s = sym_of(A.values.dtype)
    t = TYPES[s]
    a, b = t.scalar(alpha), t.scalar(beta)
    if avg is None:
        avg = max(1, int(np.ceil(A.rs.mean())) if getattr(A, "rs", None) is not None and A.nrows else 1)
    if fmt == "ell":
        getattr(L, f"spgpu{s}ellspmv")(h, dptr(dz), dptr(dy), a, dptr(dA["values"]), dptr(dA["indices"]),
                                        A.pitch, A.pitch, 0 if rs_null else dptr(dA["rs"]), dptr(dr), avg,
                                        A.maxnnz, A.nrows, dptr(dx), b, A.base if base is None else base)
    elif fmt == "hell":
        getattr(L, f"spgpu{s}hellspmv")(h, dptr(dz), dptr(dy), a, dptr(dA["values"]), dptr(dA["indices"]),
                                         A.hack_size, dptr(dA["hack_offsets"]), dptr(dA["rs"]), dptr(dr), avg,
                                         A.nrows, dptr(dx), b, A.base if base is None else base)
    elif fmt == "dia":
        getattr(L, f"spgpu{s}diaspmv")(h, dptr(dz), dptr(dy), a, dptr(dA["values"]), dptr(dA["offsets"]),
                                        A.pitch, A.nrows, A.ncols, A.diags, dptr(dx), b)
    elif fmt == "hdia":
        getattr(L, f"spgpu{s}hdiaspmv")(h, dptr(dz), dptr(dy), a, dptr(dA["values"]), dptr(dA["offsets"]),
                                         A.hack_size, dptr(dA["hack_offsets"]), A.nrows, A.ncols, dptr(dx), b)
    else:
        raise ValueError(fmt)


def dev_spmv(L, h, fmt, A, dA, x, y, alpha, beta, base=None, ridx=None, rs_null=False, avg=None,
             inplace=False):
    """Run one SpMV through the C ABI of library L on device buffers.
    dA: dict of device tensors of the format arrays.  Returns z as numpy."""
    import torch
    s = sym_of(A.values.dtype)
    t = TYPES[s]
    dx = to_dev(x)
    dy = to_dev(y)
    if inplace:
        dz = dy
    else:
        dz = torch.full((A.nrows,), float("nan"), dtype=dx.dtype, device="cuda")
    dr = to_dev(ridx)
    dev_spmv_ptr(L, h, fmt, A, dA, dz, dx, dy, alpha, beta, base=base, dr=dr, rs_null=rs_null, avg=avg)
    torch.cuda.synchronize()
    return dz.cpu().numpy()


def upload(A):
    import dataclasses
    out = {}
    for f in dataclasses.fields(A):
        v = getattr(A, f.name)
        if isinstance(v, np.ndarray):
            out[f.name] = to_dev(v)
    return out
