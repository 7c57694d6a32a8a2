"""MatrixMarket input for the SpMV harness: numpy wrappers of csrc/mmread.c (include/spgpu_mm.h),
mirroring what the reference's perf drivers do before they convert and multiply
(reference src/tests/hellPerf.cpp:60-117: loadMmProperties, loadMmMatrixToCoo, symmetric unfolding).
`write_coo` exists for tests and for exporting the synthetic matrices; it is not on any hot path."""
from __future__ import annotations

import ctypes

import numpy as np

from . import capi
from .formats import Coo

STORAGE = {0: "integer", 1: "real", 2: "complex", 3: "pattern"}
SYMMETRY = {0: "general", 1: "symmetric", 2: "skew-symmetric", 3: "hermitian"}
READ_SUCCESS, READ_UNSUPPORTED, READ_INVALID_INPUT = 0, 1, 2


class MmProperties(ctypes.Structure):
    _fields_ = [("rowsCount", ctypes.c_int), ("columnsCount", ctypes.c_int), ("nonZerosCount", ctypes.c_int),
                ("isStoredSparse", ctypes.c_int), ("matrixStorage", ctypes.c_int), ("matrixType", ctypes.c_int)]


def _bind(L):
    d = L.dll
    P, c_int = ctypes.c_void_p, ctypes.c_int
    d.spgpuMmLoadProperties.restype = c_int
    d.spgpuMmLoadProperties.argtypes = [ctypes.c_char_p, ctypes.POINTER(MmProperties)]
    d.spgpuMmLoadMatrixToCoo.restype = c_int
    d.spgpuMmLoadMatrixToCoo.argtypes = [ctypes.c_char_p, P, P, P, c_int]
    d.spgpuMmUnfoldedSymmetricSize.restype = c_int
    d.spgpuMmUnfoldedSymmetricSize.argtypes = [P, P, P, c_int, c_int]
    d.spgpuMmUnfoldSymmetric.restype = None
    d.spgpuMmUnfoldSymmetric.argtypes = [P, P, P, P, P, P, c_int, c_int]
    d.spgpuMmLoadDenseVector.restype = c_int
    d.spgpuMmLoadDenseVector.argtypes = [ctypes.c_char_p, P, c_int, c_int]
    return d


_CODES = {np.dtype(np.int32): capi.SPGPU_TYPE_INT, np.dtype(np.float32): capi.SPGPU_TYPE_FLOAT,
          np.dtype(np.float64): capi.SPGPU_TYPE_DOUBLE}


def properties(path: str, L=None) -> MmProperties:
    d = _bind(L or capi.lib())
    pr = MmProperties()
    if not d.spgpuMmLoadProperties(str(path).encode(), ctypes.byref(pr)):
        raise ValueError(f"{path}: not a valid MatrixMarket matrix file")
    return pr


def read_entries(path: str, dtype=np.float64, L=None):
    """(properties, rows, cols, vals) exactly as stored (0-based, file order); vals is None for
    a pattern file.  Raises ValueError with the reference's status code on failure."""
    d = _bind(L or capi.lib())
    pr = properties(path, L)
    n = pr.nonZerosCount
    rows, cols = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    pattern = pr.matrixStorage == 3
    vals = None if pattern else np.zeros(n, dtype=dtype)
    code = capi.SPGPU_TYPE_INT if pattern else _CODES[np.dtype(dtype)]
    rc = d.spgpuMmLoadMatrixToCoo(str(path).encode(), None if pattern else vals.ctypes.data, rows.ctypes.data,
                                  cols.ctypes.data, code)
    if rc != READ_SUCCESS:
        raise ValueError(f"{path}: spgpuMmLoadMatrixToCoo -> {rc} "
                         f"({'unsupported storage for this value type' if rc == READ_UNSUPPORTED else 'invalid input'})")
    return pr, rows, cols, vals


def read_coo(path: str, dtype=np.float64, L=None) -> Coo:
    """A .mtx coordinate file as a 0-based Coo of `dtype`, symmetric files unfolded the way the
    reference's drivers do (zero entries dropped, off-diagonal entries followed by their
    transpose); pattern files get the value 1."""
    d = _bind(L or capi.lib())
    pr, rows, cols, vals = read_entries(path, dtype, L)
    if vals is None:
        vals = np.ones(rows.shape[0], dtype=dtype)
    if pr.matrixType == 1:
        code = _CODES[np.dtype(dtype)]
        n = d.spgpuMmUnfoldedSymmetricSize(vals.ctypes.data, rows.ctypes.data, cols.ctypes.data, rows.shape[0], code)
        ur, uc, uv = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=dtype)
        d.spgpuMmUnfoldSymmetric(ur.ctypes.data, uc.ctypes.data, uv.ctypes.data, rows.ctypes.data, cols.ctypes.data,
                                 vals.ctypes.data, rows.shape[0], code)
        rows, cols, vals = ur, uc, uv
    elif pr.matrixType != 0:
        raise ValueError(f"{path}: {SYMMETRY[pr.matrixType]} matrices are not unfolded by the reference's drivers")
    return Coo(rows, cols, vals, pr.rowsCount, pr.columnsCount, 0)


def read_vector(path: str, n: int, dtype=np.float64, L=None) -> np.ndarray:
    d = _bind(L or capi.lib())
    v = np.zeros(n, dtype=dtype)
    rc = d.spgpuMmLoadDenseVector(str(path).encode(), v.ctypes.data, n, _CODES[np.dtype(dtype)])
    if rc != READ_SUCCESS:
        raise ValueError(f"{path}: spgpuMmLoadDenseVector -> {rc}")
    return v


def write_coo(path: str, coo: Coo, field="real", symmetry="general", comment="written by spgpu_b200.mmio"):
    """Write a Coo as a coordinate file (1-based).  With symmetry='symmetric' only the lower
    triangle (row >= col) is written."""
    rows, cols = coo.rows - coo.base, coo.cols - coo.base
    vals = coo.vals
    if symmetry == "symmetric":
        keep = rows >= cols
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {symmetry}\n% {comment}\n")
        f.write(f"{coo.nrows} {coo.ncols} {rows.shape[0]}\n")
        # %.17g round-trips every double (and every float) exactly
        r1, c1 = rows.astype(np.int64) + 1, cols.astype(np.int64) + 1
        if field == "pattern":
            np.savetxt(f, np.column_stack((r1, c1)), fmt="%d %d")
        elif field == "integer":
            np.savetxt(f, np.column_stack((r1, c1, np.asarray(vals).astype(np.int64))), fmt="%d %d %d")
        else:
            v = np.asarray(vals, dtype=np.float64)
            try:
                import pandas as pd                    # an order of magnitude faster than savetxt on 10^6+ lines
                pd.DataFrame({"r": r1, "c": c1, "v": v}).to_csv(f, sep=" ", header=False, index=False, float_format="%.17g")
            except ImportError:
                table = np.empty(rows.shape[0], dtype=[("r", np.int64), ("c", np.int64), ("v", np.float64)])
                table["r"], table["c"], table["v"] = r1, c1, v
                np.savetxt(f, table, fmt="%d %d %.17g")
