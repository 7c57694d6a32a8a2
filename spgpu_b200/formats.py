"""Host-side matrix formats: thin numpy wrappers over the C conversion entry points.

Every function calls the C ABI of a loaded spGPU library (ours by default; the
parity tests pass the reference library to obtain the expected arrays) exactly
the way the reference's drivers do (reference src/tests/hellPerf.cpp:120-160,
diaPerf.cpp:150-300): query a size, allocate, zero-fill, convert.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from .capi import TYPES, SpgpuLib, lib as _default_lib, ptr


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def type_of(values: np.ndarray):
    for t in TYPES.values():
        if t.np_dtype == values.dtype and t.sym != "I":
            return t
    raise TypeError(f"unsupported value dtype {values.dtype}")


@dataclass
class Coo:
    rows: np.ndarray      # int32, base `base`
    cols: np.ndarray
    vals: np.ndarray
    nrows: int
    ncols: int
    base: int = 0

    @property
    def nnz(self):
        return int(self.rows.shape[0])


@dataclass
class Ell:
    values: np.ndarray    # maxnnz * pitch, column-major slots
    indices: np.ndarray   # int32, same shape
    rs: np.ndarray        # int32 row sizes
    pitch: int
    maxnnz: int
    nrows: int
    ncols: int
    base: int
    ridx: np.ndarray | None = None


@dataclass
class Hell:
    values: np.ndarray    # height*hack_size elements
    indices: np.ndarray
    hack_offsets: np.ndarray   # int32, one per hack, ELEMENT offsets
    rs: np.ndarray
    hack_size: int
    height: int
    nrows: int
    ncols: int
    base: int
    ridx: np.ndarray | None = None


@dataclass
class Dia:
    values: np.ndarray    # diags * pitch
    offsets: np.ndarray   # int32 ascending
    pitch: int
    diags: int
    nrows: int
    ncols: int


@dataclass
class Hdia:
    values: np.ndarray    # height*hack_size elements
    offsets: np.ndarray   # int32, height entries
    hack_offsets: np.ndarray   # int32, hacks+1 entries, DIAGONAL counts
    hack_size: int
    height: int
    nrows: int
    ncols: int


def coo_to_ell(coo: Coo, ell_base: int = 0, L: SpgpuLib | None = None) -> Ell:
    L = L or _default_lib()
    t = type_of(coo.vals)
    rows, cols = _i32(coo.rows), _i32(coo.cols)
    vals = np.ascontiguousarray(coo.vals)
    rs = np.zeros(max(coo.nrows, 1), dtype=np.int32)
    maxnnz = ctypes.c_int(0)
    L.computeEllRowLenghts(ptr(rs), ctypes.byref(maxnnz), coo.nrows, coo.nnz, ptr(rows), coo.base)
    pitch = L.computeEllAllocPitch(coo.nrows)
    m = maxnnz.value
    values = np.zeros(max(m * pitch, 1), dtype=t.np_dtype)
    indices = np.zeros(max(m * pitch, 1), dtype=np.int32)
    L.cooToEll(ptr(values), ptr(indices), pitch, pitch, m, ell_base, coo.nrows, coo.nnz,
               ptr(rows), ptr(cols), ptr(vals), coo.base, t.code)
    return Ell(values[: m * pitch], indices[: m * pitch], rs[: coo.nrows], pitch, m,
               coo.nrows, coo.ncols, ell_base)


def ell_to_oell(ell: Ell, L: SpgpuLib | None = None) -> Ell:
    L = L or _default_lib()
    t = type_of(ell.values)
    ridx = np.zeros(max(ell.nrows, 1), dtype=np.int32)
    dvals = np.zeros_like(ell.values)
    dind = np.zeros_like(ell.indices)
    drs = np.zeros(max(ell.nrows, 1), dtype=np.int32)
    L.ellToOell(ptr(ridx), ptr(dvals), ptr(dind), ptr(drs), ptr(ell.values), ptr(ell.indices),
                ptr(ell.rs), ell.pitch, ell.pitch, ell.nrows, t.code)
    return Ell(dvals, dind, drs[: ell.nrows], ell.pitch, ell.maxnnz, ell.nrows, ell.ncols,
               ell.base, ridx[: ell.nrows])


def ell_to_hell(ell: Ell, hack_size: int = 32, L: SpgpuLib | None = None) -> Hell:
    L = L or _default_lib()
    t = type_of(ell.values)
    height = ctypes.c_int(0)
    rs = np.ascontiguousarray(ell.rs)
    L.computeHellAllocSize(ctypes.byref(height), hack_size, ell.nrows, ptr(rs))
    n = height.value * hack_size
    hacks = (ell.nrows + hack_size - 1) // hack_size
    # padding slots are left untouched by ellToHell (reference hell.c:91-97); fill
    # them with a poison pattern so any kernel that reads them is caught by the tests
    values = np.full(max(n, 1), np.nan, dtype=t.np_dtype)
    indices = np.full(max(n, 1), -(2 ** 30), dtype=np.int32)
    hoff = np.zeros(max(hacks, 1), dtype=np.int32)
    L.ellToHell(ptr(values), ptr(indices), ptr(hoff), hack_size, ptr(ell.values), ptr(ell.indices),
                ell.pitch, ell.pitch, ptr(rs), ell.nrows, t.code)
    return Hell(values[:n], indices[:n], hoff[:hacks], rs, hack_size, height.value,
                ell.nrows, ell.ncols, ell.base, ell.ridx)


def coo_to_dia(coo: Coo, L: SpgpuLib | None = None) -> Dia:
    L = L or _default_lib()
    t = type_of(coo.vals)
    rows, cols = _i32(coo.rows), _i32(coo.cols)
    vals = np.ascontiguousarray(coo.vals)
    # computeDiaDiagonalsCount indexes with the raw indices (reference dia.c:25-33)
    diags = L.computeDiaDiagonalsCount(coo.nrows, coo.ncols, coo.nnz, ptr(rows), ptr(cols))
    pitch = L.computeDiaAllocPitch(coo.nrows)
    values = np.zeros(max(diags * pitch, 1), dtype=t.np_dtype)
    offsets = np.zeros(max(diags, 1), dtype=np.int32)
    L.coo2dia(ptr(values), ptr(offsets), pitch, diags, coo.nrows, coo.ncols, coo.nnz,
              ptr(rows), ptr(cols), ptr(vals), coo.base, t.code)
    return Dia(values[: diags * pitch], offsets[:diags], pitch, diags, coo.nrows, coo.ncols)


def coo_to_hdia(coo: Coo, hack_size: int = 32, L: SpgpuLib | None = None) -> Hdia:
    L = L or _default_lib()
    t = type_of(coo.vals)
    rows, cols = _i32(coo.rows), _i32(coo.cols)
    vals = np.ascontiguousarray(coo.vals)
    hacks = L.getHdiaHacksCount(hack_size, coo.nrows)
    hoff = np.zeros(hacks + 1, dtype=np.int32)
    height = ctypes.c_int(0)
    L.computeHdiaHackOffsetsFromCoo(ctypes.byref(height), ptr(hoff), hack_size, coo.nrows,
                                    coo.ncols, coo.nnz, ptr(rows), ptr(cols), coo.base)
    n = height.value * hack_size
    values = np.zeros(max(n, 1), dtype=t.np_dtype)
    offsets = np.zeros(max(height.value, 1), dtype=np.int32)
    L.cooToHdia(ptr(values), ptr(offsets), ptr(hoff), hack_size, coo.nrows, coo.ncols, coo.nnz,
                ptr(rows), ptr(cols), ptr(vals), coo.base, t.code)
    return Hdia(values[:n], offsets[: height.value], hoff, hack_size, height.value,
                coo.nrows, coo.ncols)


def dia_to_hdia(dia: Dia, hack_size: int = 32, L: SpgpuLib | None = None) -> Hdia:
    L = L or _default_lib()
    t = type_of(dia.values)
    hacks = L.getHdiaHacksCount(hack_size, dia.nrows)
    hoff = np.zeros(hacks + 1, dtype=np.int32)
    height = ctypes.c_int(0)
    L.computeHdiaHackOffsets(ctypes.byref(height), ptr(hoff), hack_size, ptr(dia.values),
                             dia.pitch, dia.diags, dia.nrows, t.code)
    n = height.value * hack_size
    values = np.zeros(max(n, 1), dtype=t.np_dtype)
    offsets = np.zeros(max(height.value, 1), dtype=np.int32)
    L.diaToHdia(ptr(values), ptr(offsets), ptr(hoff), hack_size, ptr(dia.values), ptr(dia.offsets),
                dia.pitch, dia.diags, dia.nrows, t.code)
    return Hdia(values[:n], offsets[: height.value], hoff, hack_size, height.value,
                dia.nrows, dia.ncols)
