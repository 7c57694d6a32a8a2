"""Synthetic matrices of BASELINE.json's five configurations (numpy, host side).

All generators are seeded and return a `Coo` whose entries are emitted row by
row with ascending columns, which is what fixes the slot order of the ELL/HELL
conversions (reference ell.c:65-78).  Sizes are parameters so the tests can run
scaled-down instances of the same structure; bench.py builds the full-size
instances directly on the device (spgpu_b200/device_build.py) and the tests
prove both routes produce identical arrays.

RNG: numpy's PCG64 via default_rng(seed) -- deterministic across platforms.
"""
from __future__ import annotations

import numpy as np

from .formats import Coo


def _stencil(dims, offsets_values, dtype):
    """COO of a stencil on a regular grid; offsets_values = [((dz,dy,dx), v), ...]
    listed in ascending linear-offset order; x is the fastest index."""
    nz, ny, nx = dims
    n = nz * ny * nx
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    z, y, x = z.ravel(), y.ravel(), x.ravel()
    row = np.arange(n, dtype=np.int64)
    per_dir_rows, per_dir_cols, per_dir_vals, order = [], [], [], []
    for slot, ((dz, dy, dx), v) in enumerate(offsets_values):
        ok = ((z + dz >= 0) & (z + dz < nz) & (y + dy >= 0) & (y + dy < ny)
              & (x + dx >= 0) & (x + dx < nx))
        r = row[ok]
        per_dir_rows.append(r)
        per_dir_cols.append(r + (dz * ny + dy) * nx + dx)
        per_dir_vals.append(np.full(r.shape[0], v, dtype=dtype))
        order.append(np.full(r.shape[0], slot, dtype=np.int64))
    rows = np.concatenate(per_dir_rows)
    cols = np.concatenate(per_dir_cols)
    vals = np.concatenate(per_dir_vals)
    key = rows * len(offsets_values) + np.concatenate(order)
    perm = np.argsort(key, kind="stable")
    return Coo(rows[perm].astype(np.int32), cols[perm].astype(np.int32), vals[perm], n, n, 0)


def laplace2d_5pt(nx, ny=None, dtype=np.float64):
    """cfg1: 2-D 5-point Laplacian (4 / -1)."""
    ny = ny or nx
    st = [((0, -1, 0), -1), ((0, 0, -1), -1), ((0, 0, 0), 4), ((0, 0, 1), -1), ((0, 1, 0), -1)]
    return _stencil((1, ny, nx), st, dtype)


def stencil3d_27pt(n, dtype=np.float64):
    """cfg2: 3-D 27-point stencil (26 / -1), in-grid neighbours only."""
    st = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                st.append(((dz, dy, dx), 26 if (dz, dy, dx) == (0, 0, 0) else -1))
    return _stencil((n, n, n), st, dtype)


def laplace3d_7pt(n, dtype=np.float64, nz=None):
    """cfg5: 3-D 7-point Laplacian (6 / -1) on an nz x n x n grid."""
    nz = nz or n
    st = [((-1, 0, 0), -1), ((0, -1, 0), -1), ((0, 0, -1), -1), ((0, 0, 0), 6),
          ((0, 0, 1), -1), ((0, 1, 0), -1), ((1, 0, 0), -1)]
    return _stencil((nz, n, n), st, dtype)


def powerlaw_lengths(nrows, mean=16, maxlen=4096, spike_every=32768, seed=7):
    """cfg3 row lengths: Pareto(alpha=2) scaled to the requested mean, clipped to
    [1, maxlen], with one row per `spike_every` forced to maxlen."""
    rng = np.random.default_rng(seed)
    # Pareto with shape 2 and scale m has mean 2m -> m = mean/2
    raw = (mean / 2.0) * (1.0 + rng.pareto(2.0, size=nrows))
    lens = np.clip(np.floor(raw), 1, maxlen).astype(np.int64)
    if spike_every and nrows >= 1:
        spikes = np.arange(spike_every // 2 if nrows > spike_every // 2 else 0, nrows, spike_every)
        lens[spikes] = min(maxlen, nrows)
    return np.minimum(lens, nrows)


def strided_columns(lens, lo, hi, rng):
    """For each row r, lens[r] DISTINCT ascending columns in [lo[r], hi[r]]: the
    range is cut into lens[r] equal integer strata and one column is drawn in each."""
    nrows = lens.shape[0]
    start = np.concatenate(([0], np.cumsum(lens)))
    nnz = int(start[-1])
    rows = np.repeat(np.arange(nrows, dtype=np.int64), lens)
    k = np.arange(nnz, dtype=np.int64) - start[rows]
    width = (hi - lo + 1) // lens                 # >= 1 by construction
    w = width[rows]
    u = (rng.random(nnz) * w).astype(np.int64)
    u = np.minimum(u, w - 1)
    cols = lo[rows] + k * w + u
    return rows, cols


def powerlaw(nrows, mean=16, maxlen=4096, spike_every=32768, seed=7, dtype=np.float32):
    """cfg3: power-law row lengths, uniform distinct sorted columns, values U(-1,1)."""
    rng = np.random.default_rng(seed + 1)
    lens = powerlaw_lengths(nrows, mean, maxlen, spike_every, seed)
    lo = np.zeros(nrows, dtype=np.int64)
    hi = np.full(nrows, nrows - 1, dtype=np.int64)
    rows, cols = strided_columns(lens, lo, hi, rng)
    vals = (rng.random(rows.shape[0]) * 2.0 - 1.0).astype(dtype)
    return Coo(rows.astype(np.int32), cols.astype(np.int32), vals, nrows, nrows, 0)


def banded_complex(nrows, per_row=40, bw=1000, seed=11, dtype=np.complex128, base=0):
    """cfg4: ~per_row distinct sorted columns uniform in [i-bw, i+bw] (clipped),
    complex values U(-1,1)^2."""
    rng = np.random.default_rng(seed)
    i = np.arange(nrows, dtype=np.int64)
    lo = np.maximum(i - bw, 0)
    hi = np.minimum(i + bw, nrows - 1)
    lens = np.minimum(np.full(nrows, per_row, dtype=np.int64), hi - lo + 1)
    rows, cols = strided_columns(lens, lo, hi, rng)
    re = rng.random(rows.shape[0]) * 2.0 - 1.0
    im = rng.random(rows.shape[0]) * 2.0 - 1.0
    vals = (re + 1j * im).astype(dtype)
    return Coo((rows + base).astype(np.int32), (cols + base).astype(np.int32), vals,
               nrows, nrows, base)


def random_coo(nrows, ncols, density_rows=(0, 9), seed=0, dtype=np.float64, base=0,
               sort_cols=True, empty_rows=True):
    """Small irregular test matrix: ragged rows (some empty), distinct columns."""
    rng = np.random.default_rng(seed)
    lo_n, hi_n = density_rows
    lens = rng.integers(lo_n, min(hi_n, ncols) + 1, size=nrows)
    if not empty_rows:
        lens = np.maximum(lens, 1)
    rows, cols, = [], []
    for r in range(nrows):
        c = rng.choice(ncols, size=int(lens[r]), replace=False)
        if sort_cols:
            c.sort()
        rows.append(np.full(c.shape[0], r, dtype=np.int64))
        cols.append(c.astype(np.int64))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    nnz = rows.shape[0]
    dt = np.dtype(dtype)
    if dt.kind == "c":
        vals = (rng.standard_normal(nnz) + 1j * rng.standard_normal(nnz)).astype(dt)
    else:
        vals = rng.standard_normal(nnz).astype(dt)
    return Coo((rows + base).astype(np.int32), (cols + base).astype(np.int32), vals,
               nrows, ncols, base)


def random_vector(n, dtype, seed, lo=0.0, hi=1.0):
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dt.kind == "c":
        return ((lo + (hi - lo) * rng.random(n)) + 1j * (lo + (hi - lo) * rng.random(n))).astype(dt)
    return (lo + (hi - lo) * rng.random(n)).astype(dt)
