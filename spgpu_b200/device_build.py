"""Direct on-device construction of the full-size benchmark matrices (torch).

BASELINE.json's large configurations cannot take the host route COO -> ELL ->
HELL: cfg5 has 938 M non-zeros (15 GB of COO and minutes of serial host
conversion) and cfg3 would need a 137 GB ELL intermediate.  These builders write
the HELL / HDIA arrays straight into device memory with the SAME layout rules
as the host conversions (reference hell.c:46-104, hdia.cpp:230-349), and
tests/test_device_build_gpu.py proves them bit-identical to the C path at sizes
where both can run.

torch is used here as the device allocator / array language only; nothing in
this file is on the timed path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

INT_MAX = 2 ** 31 - 1
POISON_INDEX = -(2 ** 30)


@dataclass
class DevHell:
    values: torch.Tensor        # height*hack elements (padding = NaN)
    indices: torch.Tensor       # int32 (padding = POISON_INDEX)
    hack_offsets: torch.Tensor  # int32, one per hack (element offsets)
    rs: torch.Tensor            # int32 row sizes
    hack_size: int
    nrows: int
    ncols: int                  # length of the x vector the indices refer to
    nnz: int
    base: int = 0
    avg: int = 1
    ridx: torch.Tensor | None = None   # int32 output-row permutation (OHELL), or None

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.values, self.indices, self.hack_offsets, self.rs))


@dataclass
class DevHdia:
    values: torch.Tensor        # height*hack elements
    offsets: torch.Tensor       # int32, height entries
    hack_offsets: torch.Tensor  # int32, hacks+1 entries
    hack_size: int
    nrows: int
    ncols: int
    nnz: int                    # structural non-zeros of the matrix
    cells_in_range: int         # stored cells whose column is inside [0, cols)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.values, self.offsets, self.hack_offsets))


def _hack_layout(rs: torch.Tensor, hack: int):
    """rs (int32, R) -> (hack_offsets int32 [hacks], total elements)."""
    R = rs.numel()
    hacks = (R + hack - 1) // hack
    padded = torch.zeros(hacks * hack, dtype=torch.int32, device=rs.device)
    padded[:R] = rs
    longest = padded.view(hacks, hack).max(dim=1).values.to(torch.int64)
    sizes = longest * hack
    ends = torch.cumsum(sizes, 0)
    total = int(ends[-1].item()) if hacks else 0
    assert total <= INT_MAX, "HELL element offsets must fit int32 (reference ABI)"
    return (ends - sizes).to(torch.int32), total


def _alloc_hell(total, dtype, device):
    values = torch.full((max(total, 1),), float("nan"), dtype=dtype, device=device)[:total]
    indices = torch.full((max(total, 1),), POISON_INDEX, dtype=torch.int32, device=device)[:total]
    return values, indices


def hell_stencil(dims, directions, row_lo, row_hi, col_shift, ncols, dtype=torch.float64,
                 hack=32, device="cuda", chunk=1 << 24) -> DevHell:
    """HELL of rows [row_lo, row_hi) of a stencil matrix on an (nz, ny, nx) grid.

    directions: [((dz,dy,dx), value), ...] in ascending linear-offset order (this
    is the slot order cooToEll produces from a row-major, ascending-column COO).
    Stored column = global column - col_shift (local numbering of a partition);
    ncols = length of the local x vector.
    """
    nz, ny, nx = dims
    R = row_hi - row_lo
    assert row_lo % hack == 0, "partitions start on a hack boundary"
    dev = torch.device(device)

    def geometry(lo, hi):
        r = torch.arange(lo, hi, dtype=torch.int64, device=dev)
        x = r % nx
        y = (r // nx) % ny
        z = r // (nx * ny)
        return r, x, y, z

    def present(x, y, z, d):
        dz, dy, dx = d
        ok = torch.ones_like(x, dtype=torch.bool)
        if dz: ok &= (z + dz >= 0) & (z + dz < nz)
        if dy: ok &= (y + dy >= 0) & (y + dy < ny)
        if dx: ok &= (x + dx >= 0) & (x + dx < nx)
        return ok

    rs = torch.empty(R, dtype=torch.int32, device=dev)
    for lo in range(row_lo, row_hi, chunk):
        hi = min(lo + chunk, row_hi)
        _, x, y, z = geometry(lo, hi)
        n = torch.zeros(hi - lo, dtype=torch.int32, device=dev)
        for d, _v in directions:
            n += present(x, y, z, d).to(torch.int32)
        rs[lo - row_lo:hi - row_lo] = n
    hoff, total = _hack_layout(rs, hack)
    values, indices = _alloc_hell(total, dtype, dev)
    nnz = int(rs.sum(dtype=torch.int64).item())

    for lo in range(row_lo, row_hi, chunk):
        hi = min(lo + chunk, row_hi)
        r, x, y, z = geometry(lo, hi)
        local = r - row_lo
        base = hoff[(local // hack)].to(torch.int64) + (local % hack)
        slot = torch.zeros(hi - lo, dtype=torch.int64, device=dev)
        for (dz, dy, dx), v in directions:
            ok = present(x, y, z, (dz, dy, dx))
            dest = (base + slot * hack)[ok]
            col = (r + (dz * ny + dy) * nx + dx - col_shift)[ok]
            values[dest] = v
            indices[dest] = col.to(torch.int32)
            slot += ok.to(torch.int64)
            del dest, col, ok
    avg = max(1, int(round(nnz / max(R, 1))))
    return DevHell(values, indices, hoff, rs, hack, R, ncols, nnz, 0, avg)


LAPLACE7 = [((-1, 0, 0), -1.0), ((0, -1, 0), -1.0), ((0, 0, -1), -1.0), ((0, 0, 0), 6.0),
            ((0, 0, 1), -1.0), ((0, 1, 0), -1.0), ((1, 0, 0), -1.0)]


def hell_laplace3d_7pt(n, z_lo=0, z_hi=None, local_columns=False, dtype=torch.float64, hack=32,
                       device="cuda", nz=None) -> DevHell:
    """cfg5: rows of the z-slab [z_lo, z_hi) of the 7-point Laplacian on nz x n x n.

    local_columns=False: columns are global (x has nz*n*n entries).
    local_columns=True : columns index the partition's x_ext = [lower halo plane |
    owned planes | upper halo plane] (halo zones exist on every rank so the layout
    is uniform; the outermost ranks simply never reference theirs)."""
    nz = nz or n
    z_hi = nz if z_hi is None else z_hi
    plane = n * n
    row_lo, row_hi = z_lo * plane, z_hi * plane
    if local_columns:
        shift = row_lo - plane
        ncols = (z_hi - z_lo + 2) * plane
    else:
        shift, ncols = 0, nz * plane
    return hell_stencil((nz, n, n), LAPLACE7, row_lo, row_hi, shift, ncols, dtype, hack, device)


def hell_from_rows(lens: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, ncols: int,
                   hack=32, base=0) -> DevHell:
    """HELL from row-sorted entries: row r owns the next lens[r] entries of cols/vals
    (ascending columns inside a row).  Same placement as cooToEll + ellToHell."""
    dev = cols.device
    R = lens.numel()
    rs = lens.to(torch.int32)
    hoff, total = _hack_layout(rs, hack)
    values, indices = _alloc_hell(total, vals.dtype, dev)
    start = torch.cumsum(lens.to(torch.int64), 0) - lens.to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(R, dtype=torch.int64, device=dev), lens.to(torch.int64))
    k = torch.arange(cols.numel(), dtype=torch.int64, device=dev) - start[rows]
    dest = hoff[rows // hack].to(torch.int64) + k * hack + rows % hack
    del rows, k, start
    values[dest] = vals
    indices[dest] = (cols + base).to(torch.int32)
    nnz = int(cols.numel())
    return DevHell(values, indices, hoff, rs, hack, R, ncols, nnz, base, max(1, int(round(nnz / max(R, 1)))))


def sort_rows_by_length(lens: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor):
    """Device twin of ellToOell's ordering (reference ell.c:161-202): rows by descending
    length, ties by descending row index.  Returns (lens', cols', vals', rIdx) with
    rIdx[i] = original row now stored as row i."""
    R = lens.numel()
    idx = torch.arange(R, dtype=torch.int64, device=lens.device)
    key = lens.to(torch.int64) * R + idx                 # descending key = length desc, index desc
    perm = torch.argsort(key, descending=True)
    new_lens = lens[perm]
    old_start = torch.cumsum(lens.to(torch.int64), 0) - lens.to(torch.int64)
    new_start = torch.cumsum(new_lens.to(torch.int64), 0) - new_lens.to(torch.int64)
    new_rows = torch.repeat_interleave(idx, new_lens.to(torch.int64))
    k = torch.arange(cols.numel(), dtype=torch.int64, device=lens.device) - new_start[new_rows]
    src = old_start[perm[new_rows]] + k
    return new_lens, cols[src], vals[src], perm.to(torch.int32)


def _strided_columns(lens, lo, hi, gen):
    """device twin of generators.strided_columns: lens[r] distinct ascending columns
    in [lo[r], hi[r]], one per equal-width integer stratum."""
    dev = lens.device
    R = lens.numel()
    start = torch.cumsum(lens, 0) - lens
    rows = torch.repeat_interleave(torch.arange(R, dtype=torch.int64, device=dev), lens)
    nnz = rows.numel()
    k = torch.arange(nnz, dtype=torch.int64, device=dev) - start[rows]
    w = ((hi - lo + 1) // lens)[rows]
    u = (torch.rand(nnz, device=dev, generator=gen, dtype=torch.float64) * w).to(torch.int64)
    u = torch.minimum(u, w - 1)
    return rows, lo[rows] + k * w + u


def powerlaw_entries(nrows, mean=16, maxlen=4096, spike_every=32768, seed=7, dtype=torch.float32,
                     device="cuda"):
    """cfg3 entries on the device: (lens, cols, vals).  Row lengths come from the
    seeded numpy generator (generators.powerlaw_lengths), columns/values from a
    seeded torch generator."""
    from .generators import powerlaw_lengths
    lens = torch.from_numpy(powerlaw_lengths(nrows, mean, maxlen, spike_every, seed)).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    lo = torch.zeros(nrows, dtype=torch.int64, device=device)
    hi = torch.full((nrows,), nrows - 1, dtype=torch.int64, device=device)
    _rows, cols = _strided_columns(lens, lo, hi, gen)
    vals = (torch.rand(cols.numel(), device=device, generator=gen, dtype=torch.float32) * 2 - 1).to(dtype)
    return lens, cols, vals


def banded_complex_entries(nrows, per_row=40, bw=1000, seed=11, dtype=torch.complex128, device="cuda"):
    """cfg4 entries on the device: ~per_row distinct sorted columns in [i-bw, i+bw]."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    i = torch.arange(nrows, dtype=torch.int64, device=device)
    lo = torch.clamp(i - bw, min=0)
    hi = torch.clamp(i + bw, max=nrows - 1)
    lens = torch.minimum(torch.full_like(i, per_row), hi - lo + 1)
    _rows, cols = _strided_columns(lens, lo, hi, gen)
    real_dt = torch.float64 if dtype == torch.complex128 else torch.float32
    re = torch.rand(cols.numel(), device=device, generator=gen, dtype=real_dt) * 2 - 1
    im = torch.rand(cols.numel(), device=device, generator=gen, dtype=real_dt) * 2 - 1
    return lens, cols, torch.complex(re, im)


def hdia_stencil27(n, dtype=torch.float64, hack=32, device="cuda") -> DevHdia:
    """cfg2: HDIA of the 27-point stencil (26 / -1) on n^3, n a multiple of `hack`
    (so the rows of a hack share y and z).  Same output as cooToHdia."""
    assert n % hack == 0 and n >= 3
    dev = torch.device(device)
    per_line = n // hack
    H = n * n * per_line
    h = torch.arange(H, dtype=torch.int64, device=dev)
    a = h % per_line
    y = (h // per_line) % n
    z = h // (per_line * n)
    dirs = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]

    def present(dz, dy):
        return (z + dz >= 0) & (z + dz < n) & (y + dy >= 0) & (y + dy < n)

    count = torch.zeros(H, dtype=torch.int64, device=dev)
    for dz, dy, dx in dirs:
        count += present(dz, dy).to(torch.int64)
    ends = torch.cumsum(count, 0)
    height = int(ends[-1].item())
    hoff = torch.zeros(H + 1, dtype=torch.int32, device=dev)
    hoff[1:] = ends.to(torch.int32)
    first = ends - count
    offsets = torch.empty(height, dtype=torch.int32, device=dev)
    values = torch.zeros(height * hack, dtype=dtype, device=dev).view(height, hack)
    lane = torch.arange(hack, dtype=torch.int64, device=dev)
    xs = (a * hack)[:, None] + lane[None, :]
    pos = torch.zeros(H, dtype=torch.int64, device=dev)
    rows_of = (h * hack)[:, None] + lane[None, :]
    nnz = 0
    in_range = 0
    for dz, dy, dx in dirs:
        ok = present(dz, dy)
        d = (first + pos)[ok]
        off = (dz * n + dy) * n + dx
        offsets[d] = off
        c = rows_of[ok] + off
        in_range += int(((c >= 0) & (c < n ** 3)).sum().item())
        del c
        inside = ((xs + dx >= 0) & (xs + dx < n))[ok]
        v = 26.0 if (dz, dy, dx) == (0, 0, 0) else -1.0
        values[d] = torch.where(inside, torch.full((), v, dtype=dtype, device=dev),
                                torch.zeros((), dtype=dtype, device=dev))
        nnz += int(inside.sum().item())
        pos += ok.to(torch.int64)
    return DevHdia(values.view(-1), offsets, hoff, hack, n ** 3, n ** 3, nnz, in_range)


def hdia_row_block(A: DevHdia, lo: int, hi: int, halo: int) -> DevHdia:
    """Rows [lo, hi) of A (lo a multiple of the hack size) as a self-contained HDIA block on
    x_ext = [halo | owned | halo]: torch twin of mg.split_hdia (same rules: hackOffsets re-based,
    offsets + halo, cells whose global column is outside the matrix zeroed, non-zero cells outside
    the window refused).  ncols of the result is the length of x_ext."""
    hs = A.hack_size
    assert lo % hs == 0 and 0 <= lo < hi <= A.nrows
    h0, h1 = lo // hs, (hi + hs - 1) // hs
    hoff = A.hack_offsets.to(torch.int64)
    d0, d1 = int(hoff[h0].item()), int(hoff[h1].item())
    values = A.values[d0 * hs:d1 * hs].clone().view(d1 - d0, hs)
    offs = A.offsets[d0:d1].to(torch.int64)
    local_hoff = (hoff[h0:h1 + 1] - d0).to(torch.int32)
    dev = values.device
    hack_of_diag = torch.repeat_interleave(torch.arange(h1 - h0, dtype=torch.int64, device=dev),
                                           (hoff[h0 + 1:h1 + 1] - hoff[h0:h1]))
    rows = hack_of_diag[:, None] * hs + torch.arange(hs, dtype=torch.int64, device=dev)[None, :]
    cols = lo + rows + offs[:, None]
    outside_matrix = (cols < 0) | (cols >= A.ncols) | (rows >= hi - lo)
    values[outside_matrix] = 0
    outside_window = ~outside_matrix & ((cols < lo - halo) | (cols >= hi + halo))
    if bool((values[outside_window] != 0).any().item()):
        raise ValueError("diagonal reaches outside the halo window")
    values[outside_window] = 0
    in_range = int((~outside_matrix).sum().item())
    nnz = int((values != 0).sum().item())
    return DevHdia(values.view(-1), (offs + halo).to(torch.int32), local_hoff, hs, hi - lo, hi - lo + 2 * halo,
                   nnz, in_range)


@dataclass
class DevDia:
    values: torch.Tensor        # diags * pitch
    offsets: torch.Tensor       # int32 ascending
    pitch: int
    diags: int
    nrows: int
    ncols: int
    nnz: int
    cells_in_range: int


def dia_stencil27(n, dtype=torch.float64, device="cuda") -> DevDia:
    """cfg2 as plain DIA: 27 diagonals of the 27-point stencil (26 / -1) on n^3; same output
    as coo2dia on the row-major COO (zero where the neighbour falls outside the grid)."""
    dev = torch.device(device)
    N = n ** 3
    pitch = (N + 31) // 32 * 32
    r = torch.arange(N, dtype=torch.int64, device=dev)
    x, y, z = r % n, (r // n) % n, r // (n * n)
    dirs = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    values = torch.zeros(27 * pitch, dtype=dtype, device=dev).view(27, pitch)
    offs, nnz, in_range = [], 0, 0
    for j, (dz, dy, dx) in enumerate(dirs):
        ok = ((z + dz >= 0) & (z + dz < n) & (y + dy >= 0) & (y + dy < n) & (x + dx >= 0) & (x + dx < n))
        v = 26.0 if (dz, dy, dx) == (0, 0, 0) else -1.0
        values[j, :N] = torch.where(ok, torch.full((), v, dtype=dtype, device=dev), torch.zeros((), dtype=dtype, device=dev))
        off = (dz * n + dy) * n + dx
        offs.append(off)
        nnz += int(ok.sum().item())
        c = r + off
        in_range += int(((c >= 0) & (c < N)).sum().item())
    return DevDia(values.view(-1), torch.tensor(offs, dtype=torch.int32, device=dev), pitch, 27, N, N, nnz, in_range)


def to_host_hell(d: DevHell):
    """numpy copies, for the bit-exactness tests"""
    return (d.values.cpu().numpy(), d.indices.cpu().numpy(), d.hack_offsets.cpu().numpy(), d.rs.cpu().numpy())
