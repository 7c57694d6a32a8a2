#!/usr/bin/env python
"""bench.py -- SpMV GFLOP/s and achieved HBM GB/s of the B200-native spGPU drop-in.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE SpMV z = A*x over the whole matrix through the C ABI
(spgpuDhellspmv / spgpuDhdiaspmv / ...).  Default workload: BASELINE.json's
configs[4], the configuration the metric's 1/2/4/8-GPU scaling is quoted on --
the 512^3 7-point Laplacian (134 M rows, 938 M non-zeros, 13.95 GB of
algorithmic traffic per SpMV) in double HELL, hackSize 32, row-sharded by
z-slabs with a one-plane x halo per neighbour ("strong" scaling: the matrix is
fixed, N changes).  The other configurations are selectable with --workload
(cfg1 ELL 2-D Laplacian, cfg2 HDIA 27-point, cfg3 float power-law HELL, cfg4
complex banded HELL) and run on one GPU.

Prints ONE JSON line (rank 0).  `value`: device-resident inputs, CUDA events on
the launching stream, max over ranks.  `e2e`: same call with x coming from
pinned host memory and z going back to it inside the timed region.
`--impl reference`: the CPU arm -- OpenMP port of the same format's loop
(oracle/liboracle.so) on the box's host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import atexit
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0      # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


# --------------------------------------------------------------------------- #
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------- #

class ClockSampler:
    """`nvidia-smi -lms 20` in the background.  The first query on a fresh box can take over a second, so the
    process is started well before the timed region; `mark()` is called when the warm-up begins and only the
    lines printed after it count.  If none has arrived by `stop()`, the GPU is kept busy with more (untimed) steps
    until one does -- a sample taken on an idle GPU would say nothing about the clocks under load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []                 # (monotonic time, text)
        self.t_mark = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            atexit.register(self._kill)         # a path that leaves without stop() must not leave nvidia-smi behind
        except Exception:
            self.proc = None

    def _kill(self):
        if self.proc and self.proc.poll() is None:
            self.proc.kill()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        self.t_mark = time.monotonic()

    def _parsed(self):
        out = []
        for t, ln in list(self.lines):
            f = [x.strip() for x in ln.split(",")]
            if t < self.t_mark or len(f) < 9:
                continue
            try:
                out.append((float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        return out

    def stop(self, keep_busy=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        extended = False
        t0 = time.monotonic()
        while keep_busy is not None and not self._parsed() and time.monotonic() - t0 < 5.0 and self.proc.poll() is None:
            keep_busy()                 # still the same kernels on the same stream, just not timed
            extended = True
        time.sleep(0.15 if not extended else 0.0)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self._parsed()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for _, _, flags in rows:
            for name, v in zip(names, flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm, mx = [r[0] for r in rows], [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
# workloads
# --------------------------------------------------------------------------- #

def algorithmic_bytes_hell(nnz, rows, hacks, ncols_read, sizeof_t, beta_nonzero=False):
    """nnz*(sizeof(T)+4) + 4R (rS) + 4*hacks + x once + z (+ y)."""
    return nnz * (sizeof_t + 4) + 4 * rows + 4 * hacks + ncols_read * sizeof_t + rows * sizeof_t * (2 if beta_nonzero else 1)


def hell_matrix_floor_bytes(A, sizeof_t):
    """DRAM bytes of the matrix arrays a HELL SpMV cannot avoid at a fetch granularity of G bytes: slot k of a hack is
    a run of hackSize values (and one of hackSize indices); a G-byte piece of that run has to come in as soon as ONE of
    its rows is longer than k.  G = 32 is a sector, 64 what `ld.global.L2::64B` fetches, 128 the line every other
    load flavour fetches on B200 (bench/probes/sector_probe.cu).  Returns {"32B": .., "64B": .., "128B": ..}."""
    import torch
    rs = A.rs.to(torch.int64)
    hs = A.hack_size
    pad = (-rs.numel()) % hs
    if pad:
        rs = torch.cat([rs, torch.zeros(pad, dtype=rs.dtype, device=rs.device)])
    out = {}
    for G in (32, 64, 128):
        total = 0
        for esz in (sizeof_t, 4):                      # the value run and the index run of every slot
            rows_per_piece = max(1, min(hs, G // esz))
            piece = max(G, rows_per_piece * esz) if rows_per_piece == hs else G
            total += int(rs.view(-1, rows_per_piece).max(dim=1).values.sum().item()) * piece
        out[f"{G}B"] = total
    return out


def build_workload(name, rank, world, device, n_override=None, global_columns=False):
    """Returns a dict describing this rank's share of the workload (device resident)."""
    import torch
    from spgpu_b200 import device_build as DB
    w = {"name": name}
    if name == "cfg5":
        n = n_override or 512
        assert n % world == 0
        per = n // world
        z_lo, z_hi = rank * per, (rank + 1) * per
        plane = n * n
        local = world > 1 and not global_columns
        A = DB.hell_laplace3d_7pt(n, z_lo, z_hi, local_columns=local, device=device)
        w.update(kind="hell", sym="D", A=A, rows=A.nrows, nnz=A.nnz, halo=plane if local else 0,
                 x_len=A.ncols, sizeof=8, alpha=1.0, beta=0.0, flops_per_nnz=2,
                 label=f"3-D 7-point Laplacian {n}^3, double HELL hackSize 32 (BASELINE configs[4])",
                 total_rows=n ** 3, bandwidth=plane)
        x_read = A.nrows + (2 * plane if world > 1 else 0)      # owned entries + the two halo planes
        w["allgather"] = world > 1 and global_columns
        w["bytes"] = algorithmic_bytes_hell(A.nnz, A.nrows, A.hack_offsets.numel(), x_read, 8)
    elif name == "cfg2":
        n = n_override or 128
        A = DB.hdia_stencil27(n, device=device)
        total_rows, halo = A.nrows, 0
        if world > 1:
            # row block of z-planes on x_ext = [halo | owned | halo]; the 27-point stencil reaches
            # n*n + n + 1 entries, rounded up to whole hacks (mg.split_hdia's rules, on the device)
            from spgpu_b200 import mg
            halo = -(-(n * n + n + 1) // 32) * 32
            lo, hi = mg.row_blocks(A.nrows, world, 32)[rank]
            w["full"] = A
            w["lo"], w["hi"] = lo, hi
            A = DB.hdia_row_block(A, lo, hi, halo)
        hd = int(A.offsets.numel())
        w.update(kind="hdia", sym="D", A=A, rows=A.nrows, nnz=A.nnz, halo=halo, x_len=A.ncols, sizeof=8,
                 alpha=1.0, beta=0.0, flops_per_nnz=2, total_rows=total_rows, bandwidth=n * n + n + 1,
                 label=f"3-D 27-point stencil {n}^3, double HDIA hackSize 32 (BASELINE configs[1])")
        # cells_in_range*8 + 4*#hack-diagonals + 4*(hacks+1) + x (owned + the two halo windows) + z
        w["bytes"] = A.cells_in_range * 8 + 4 * hd + 4 * int(A.hack_offsets.numel()) + 8 * A.ncols + 8 * A.nrows
    elif name == "cfg2dia":
        n = n_override or 128
        A = DB.dia_stencil27(n, device=device)
        w.update(kind="dia", sym="D", A=A, rows=A.nrows, nnz=A.nnz, halo=0, x_len=A.ncols, sizeof=8,
                 alpha=1.0, beta=0.0, flops_per_nnz=2, total_rows=A.nrows,
                 label=f"3-D 27-point stencil {n}^3, double DIA (27 diagonals; BASELINE configs[1] stored as plain DIA)")
        w["bytes"] = A.cells_in_range * 8 + 4 * 27 + 8 * A.ncols + 8 * A.nrows
    elif name == "cfg1":
        from spgpu_b200 import formats as F, generators as G
        n = n_override or 1000
        ell = F.coo_to_ell(G.laplace2d_5pt(n))
        A = {"values": torch.from_numpy(ell.values).to(device), "indices": torch.from_numpy(ell.indices).to(device),
             "rs": torch.from_numpy(ell.rs).to(device), "ell": ell}
        nnz = int(ell.rs.sum())
        w.update(kind="ell", sym="D", A=A, rows=ell.nrows, nnz=nnz, halo=0, x_len=ell.ncols, sizeof=8,
                 alpha=1.0, beta=0.0, flops_per_nnz=2, total_rows=ell.nrows,
                 label=f"2-D 5-point Laplacian {n}x{n}, double ELL with rS (BASELINE configs[0])",
                 kernel="ell_spmv_short_kernel<D, 5 slots>")
        w["bytes"] = nnz * 12 + 4 * ell.nrows + 8 * ell.ncols + 8 * ell.nrows
    elif name == "cfg3":
        R = n_override or (1 << 22)
        lens, cols, vals = DB.powerlaw_entries(R, device=device)
        A = DB.hell_from_rows(lens, cols, vals, R)
        del lens, cols, vals
        w.update(kind="hell", sym="S", A=A, rows=R, nnz=A.nnz, halo=0, x_len=R, sizeof=4, alpha=1.0, beta=0.0,
                 flops_per_nnz=2, total_rows=R,
                 label=f"power-law rows {R} avg 16 max 4096, float HELL hackSize 32 (BASELINE configs[2])")
        w["bytes"] = algorithmic_bytes_hell(A.nnz, R, A.hack_offsets.numel(), R, 4)
        w["ell_bytes_avoided"] = int(A.rs.max().item()) * R * 8
    elif name == "cfg3o":
        # cfg3's matrix stored as OHELL: rows sorted by length (the reference's ellToOell order) with
        # rIdx sending results home -- hacks become homogeneous, so HELL's padding sectors disappear
        R = n_override or (1 << 22)
        lens, cols, vals = DB.powerlaw_entries(R, device=device)
        lens, cols, vals, ridx = DB.sort_rows_by_length(lens, cols, vals)
        A = DB.hell_from_rows(lens, cols, vals, R)
        A.ridx = ridx
        del lens, cols, vals
        w.update(kind="hell", sym="S", A=A, rows=R, nnz=A.nnz, halo=0, x_len=R, sizeof=4, alpha=1.0, beta=0.0,
                 flops_per_nnz=2, total_rows=R,
                 label=f"power-law rows {R} avg 16 max 4096, float OHELL (rows sorted by length + rIdx) hackSize 32")
        w["bytes"] = algorithmic_bytes_hell(A.nnz, R, A.hack_offsets.numel(), R, 4) + 4 * R
    elif name == "cfg4":
        R = n_override or 2_000_000
        lens, cols, vals = DB.banded_complex_entries(R, device=device)
        lo, hi, halo = 0, R, 0
        if world > 1:
            # row block of this rank with columns remapped into x_ext = [halo | owned | halo]; every rank generates
            # the same matrix (same seed) and keeps its rows.  bw = 1000 -> halo 1024 (whole hacks)
            from spgpu_b200 import mg
            halo = 1024
            lo, hi = mg.row_blocks(R, world, 32)[rank]
            ends = torch.cumsum(lens, 0)
            e0, e1 = int((ends[lo] - lens[lo]).item()), int(ends[hi - 1].item())
            lens, cols, vals = lens[lo:hi].contiguous(), cols[e0:e1].contiguous(), vals[e0:e1].contiguous()
            w["A_global"] = DB.hell_from_rows(lens, cols, vals, R)            # the same rows with global columns (--verify)
            cols = cols - (lo - halo)
        A = DB.hell_from_rows(lens, cols, vals, (hi - lo) + 2 * halo if world > 1 else R)
        del lens, cols, vals
        w.update(kind="hell", sym="Z", A=A, rows=hi - lo, nnz=A.nnz, halo=halo, x_len=A.ncols, sizeof=16,
                 alpha=0.7 - 0.3j, beta=-0.5 + 0.25j, flops_per_nnz=8, total_rows=R, bandwidth=1000, lo=lo, hi=hi,
                 label=f"banded complex-double {R} rows ~40 nnz/row, HELL hackSize 32 (BASELINE configs[3])")
        w["bytes"] = algorithmic_bytes_hell(A.nnz, hi - lo, A.hack_offsets.numel(), A.ncols, 16, beta_nonzero=True)
    else:
        raise SystemExit(f"unknown workload {name}")
    return w


def build_mtx_workload(path, fmt, L, h, device, dtype_sym="D"):
    """A MatrixMarket file run through the same harness: read with the C reader
    (spgpuMm*, symmetric files unfolded like the reference's drivers), laid out ON THE DEVICE with the
    ext builders (hell / ohell: CSR -> HELL, hdia: COO -> HDIA) or on the host (ell, dia: the
    reference's own conversions).  Needs a handle: the builders are library calls."""
    import torch
    from spgpu_b200 import device_build as DB, formats as F, mmio
    np_dt, t_dt, sz = {"S": (np.float32, torch.float32, 4), "D": (np.float64, torch.float64, 8)}[dtype_sym]
    coo = mmio.read_coo(path, np_dt, L)
    order = np.lexsort((coo.cols, coo.rows))                  # row-major, ascending columns
    coo = F.Coo(coo.rows[order], coo.cols[order], coo.vals[order], coo.nrows, coo.ncols, 0)
    R, C, nnz = coo.nrows, coo.ncols, coo.nnz
    w = {"name": "mtx", "sym": dtype_sym, "rows": R, "nnz": nnz, "halo": 0, "x_len": C, "sizeof": sz, "alpha": 1.0,
         "beta": 0.0, "flops_per_nnz": 2, "total_rows": R, "coo": coo,
         "label": f"{os.path.basename(path)} ({R}x{C}, {nnz} nnz after unfolding) as {fmt.upper()} {'float' if dtype_sym == 'S' else 'double'}"}
    d_cols = torch.from_numpy(coo.cols).to(device)
    d_vals = torch.from_numpy(coo.vals).to(device)
    avg = max(1, int(round(nnz / max(R, 1))))
    if fmt in ("hell", "ohell"):
        counts = np.bincount(coo.rows, minlength=R)
        d_rowptr = torch.from_numpy(np.concatenate(([0], np.cumsum(counts))).astype(np.int32)).to(device)
        hacks = (R + 31) // 32
        rs = torch.zeros(R, dtype=torch.int32, device=device)
        hoff = torch.zeros(hacks, dtype=torch.int32, device=device)
        ridx = torch.zeros(R, dtype=torch.int32, device=device) if fmt == "ohell" else None
        total = ctypes.c_longlong(0)
        if fmt == "ohell":
            rc = L.spgpuCsrToOhellLayoutDevice(h, R, d_rowptr.data_ptr(), 32, ridx.data_ptr(), rs.data_ptr(), hoff.data_ptr(),
                                               ctypes.byref(total))
        else:
            rc = L.spgpuCsrToHellLayoutDevice(h, R, d_rowptr.data_ptr(), 32, rs.data_ptr(), hoff.data_ptr(), ctypes.byref(total))
        assert rc == 0, f"HELL layout -> {rc}"
        values = torch.full((max(total.value, 1),), float("nan"), dtype=t_dt, device=device)
        indices = torch.full((max(total.value, 1),), DB.POISON_INDEX, dtype=torch.int32, device=device)
        if fmt == "ohell":
            getattr(L, f"spgpu{dtype_sym}csrToOhellDevice")(h, R, d_rowptr.data_ptr(), d_cols.data_ptr(), d_vals.data_ptr(), 0, 32,
                                                            hoff.data_ptr(), ridx.data_ptr(), 0, values.data_ptr(), indices.data_ptr())
        else:
            getattr(L, f"spgpu{dtype_sym}csrToHellDevice")(h, R, d_rowptr.data_ptr(), d_cols.data_ptr(), d_vals.data_ptr(), 0, 32,
                                                           hoff.data_ptr(), 0, values.data_ptr(), indices.data_ptr())
        torch.cuda.synchronize()
        A = DB.DevHell(values, indices, hoff, rs, 32, R, C, nnz, 0, avg, ridx)
        w.update(kind="hell", A=A, stored_bytes=int(total.value) * (sz + 4),
                 bytes=algorithmic_bytes_hell(nnz, R, hacks, C, sz) + (4 * R if fmt == "ohell" else 0))
    elif fmt == "hdia":
        d_rows = torch.from_numpy(coo.rows).to(device)
        hacks = (R + 31) // 32
        hoff = torch.zeros(hacks + 1, dtype=torch.int32, device=device)
        height = ctypes.c_int(0)
        rc = L.spgpuHdiaHackOffsetsFromCooDevice(h, ctypes.byref(height), hoff.data_ptr(), 32, R, C, nnz, d_rows.data_ptr(),
                                                 d_cols.data_ptr(), 0)
        assert rc == 0, f"HDIA layout -> {rc}"
        values = torch.zeros(max(height.value * 32, 1), dtype=t_dt, device=device)
        offsets = torch.zeros(max(height.value, 1), dtype=torch.int32, device=device)
        rc = getattr(L, f"spgpu{dtype_sym}cooToHdiaDevice")(h, values.data_ptr(), offsets.data_ptr(), hoff.data_ptr(), 32, R, C, nnz,
                                                            d_rows.data_ptr(), d_cols.data_ptr(), d_vals.data_ptr(), 0)
        assert rc == 0
        torch.cuda.synchronize()
        # cells whose column is inside the matrix: per hack-diagonal, rows [row0, row0+32) cut by 0 <= row+off < C
        off64 = offsets[:height.value].to(torch.int64)
        row0 = torch.repeat_interleave(torch.arange(hacks, dtype=torch.int64, device=device), (hoff[1:] - hoff[:-1]).to(torch.int64)) * 32
        lo = torch.maximum(row0, -off64)
        hi = torch.minimum(torch.minimum(row0 + 32, torch.full_like(row0, R)), C - off64)
        in_range = int(torch.clamp(hi - lo, min=0).sum().item())
        A = DB.DevHdia(values, offsets, hoff, 32, R, C, nnz, in_range)
        w.update(kind="hdia", A=A, stored_bytes=height.value * 32 * sz,
                 bytes=in_range * sz + 4 * height.value + 4 * (hacks + 1) + sz * C + sz * R)
    elif fmt == "ell":
        ell = F.coo_to_ell(coo)
        A = {"values": torch.from_numpy(ell.values).to(device), "indices": torch.from_numpy(ell.indices).to(device),
             "rs": torch.from_numpy(ell.rs).to(device), "ell": ell, "avg": avg}
        w.update(kind="ell", A=A, stored_bytes=ell.values.nbytes + ell.indices.nbytes, bytes=nnz * (sz + 4) + 4 * R + sz * C + sz * R)
    elif fmt == "dia":
        dia = F.coo_to_dia(coo)
        r = np.arange(R, dtype=np.int64)[None, :] + dia.offsets.astype(np.int64)[:, None]
        in_range = int(((r >= 0) & (r < C)).sum())
        A = DB.DevDia(torch.from_numpy(dia.values).to(device), torch.from_numpy(dia.offsets).to(device), dia.pitch, dia.diags,
                      R, C, nnz, in_range)
        w.update(kind="dia", A=A, stored_bytes=dia.values.nbytes, bytes=in_range * sz + 4 * dia.diags + sz * C + sz * R)
    else:
        raise SystemExit(f"unknown format {fmt}")
    return w


def make_step(L, h, w, x_ext_ptr, z_ptr, y_ptr):
    """closure(row0, row1) launching the SpMV of rows [row0,row1) through the C ABI"""
    from spgpu_b200.capi import TYPES
    t = TYPES[w["sym"]]
    a, b = t.scalar(w["alpha"]), t.scalar(w["beta"])
    A, s, sz = w["A"], w["sym"], w["sizeof"]
    if w["kind"] == "hell":
        fn = getattr(L, f"spgpu{s}hellspmv")
        cM, rP, ho, rs = A.values.data_ptr(), A.indices.data_ptr(), A.hack_offsets.data_ptr(), A.rs.data_ptr()
        hs, avg, base = A.hack_size, A.avg, A.base
        ridx = A.ridx.data_ptr() if getattr(A, "ridx", None) is not None else 0

        def step(r0=0, r1=w["rows"]):
            if ridx:             # rIdx addresses z/y absolutely: no pointer offset for a sub-range
                fn(h, z_ptr, y_ptr or 0, a, cM, rP, hs, ho + 4 * (r0 // hs), rs + 4 * r0, ridx + 4 * r0, avg,
                   r1 - r0, x_ext_ptr, b, base)
            else:
                fn(h, z_ptr + sz * r0, (y_ptr + sz * r0) if y_ptr else 0, a, cM, rP, hs, ho + 4 * (r0 // hs),
                   rs + 4 * r0, 0, avg, r1 - r0, x_ext_ptr, b, base)
    elif w["kind"] == "hdia":
        fn = getattr(L, f"spgpu{s}hdiaspmv")
        dM, off, ho = A.values.data_ptr(), A.offsets.data_ptr(), A.hack_offsets.data_ptr()

        def step(r0=0, r1=w["rows"]):
            # rows [r0, r1): hacks from r0/hackSize on; column = row + offset, so x moves with the rows
            fn(h, z_ptr + sz * r0, (y_ptr + sz * r0) if y_ptr else 0, a, dM, off, A.hack_size,
               ho + 4 * (r0 // A.hack_size), r1 - r0, A.ncols - r0, x_ext_ptr + sz * r0, b)
    elif w["kind"] == "dia":
        fn = getattr(L, f"spgpu{s}diaspmv")

        def step(r0=0, r1=w["rows"]):
            fn(h, z_ptr, y_ptr or 0, a, A.values.data_ptr(), A.offsets.data_ptr(), A.pitch, A.nrows, A.ncols, A.diags,
               x_ext_ptr, b)
    elif w["kind"] == "ell":
        fn = getattr(L, f"spgpu{s}ellspmv")
        ell = A["ell"]

        def step(r0=0, r1=w["rows"]):
            fn(h, z_ptr, y_ptr or 0, a, A["values"].data_ptr(), A["indices"].data_ptr(), ell.pitch, ell.pitch,
               A["rs"].data_ptr(), 0, A.get("avg", 4), ell.maxnnz, ell.nrows, x_ext_ptr, b, 0)
    else:
        raise ValueError(w["kind"])
    return step


# --------------------------------------------------------------------------- #
# CPU arm (OpenMP port of the same loop) -- oracle/liboracle.so
# --------------------------------------------------------------------------- #

def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_baseline(workload_name, budget_s=20.0, repeats=5, warmup=1, size=None):
    """Times the OpenMP host SpMV (oracle port) over the SAME format, all host threads: `warmup`
    untimed calls, then up to `repeats` timed calls (stops early once budget_s is spent).  cfg5 is
    multiplied at FULL size when the host has the memory for it (14 GB of arrays), else on a
    128-plane slab of the same grid; the `sample` string says which.  Returns (cpu_baseline
    object, times) with times = the wall time of every timed call."""
    import torch
    from tests import util
    from spgpu_b200 import device_build as DB
    if not os.path.exists(util.ORACLE_PATH):       # normally prebuilt by __graft_entry__.build()
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    O = util.oracle_lib()
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    O.dll.oracle_set_num_threads(avail)
    cores = O.dll.oracle_num_threads()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    if workload_name == "cfg5":
        n = size or 512
        full = host_mem_available_gb() > 3.5 * (n ** 3) * 7 * 12 / 1e9     # arrays + the pageable staging of the copy
        nz = n if full else max(1, n // 4)
        A = DB.hell_laplace3d_7pt(n, 0, nz, local_columns=False, device=dev, nz=nz)
        sample = (f"the FULL workload ({A.nrows} rows, {A.nnz} nnz), HELL double" if full else
                  f"{nz} z-planes of the same {n}x{n}-plane 7-point Laplacian ({A.nrows} rows, {A.nnz} nnz; "
                  f"host memory too small for the full matrix), HELL double")
        vals, idx, ho, rs = (t.cpu().numpy() for t in (A.values, A.indices, A.hack_offsets, A.rs))
        nrows, ncols, nnz = A.nrows, A.ncols, A.nnz
        del A
        if dev == "cuda":
            torch.cuda.empty_cache()
        x = np.random.default_rng(12345).random(ncols)
        z = np.zeros(nrows)
        T = util.TYPES["D"]
        call = lambda: O.Dhellspmv(util.ptr(z), None, T.scalar(1.0), util.ptr(vals), util.ptr(idx), 32,
                                   util.ptr(ho), util.ptr(rs), None, 7, nrows, util.ptr(x), T.scalar(0.0), 0)
        flops = 2 * nnz
    elif workload_name == "cfg2":
        A = DB.hdia_stencil27(128, device=dev)
        sample = f"the full 128^3 27-point HDIA matrix ({A.nrows} rows, {A.nnz} nnz)"
        vals, off, ho = (t.cpu().numpy() for t in (A.values, A.offsets, A.hack_offsets))
        x = np.random.default_rng(12345).random(A.ncols)
        z = np.zeros(A.nrows)
        T = util.TYPES["D"]
        call = lambda: O.Dhdiaspmv(util.ptr(z), None, T.scalar(1.0), util.ptr(vals), util.ptr(off), 32,
                                   util.ptr(ho), A.nrows, A.ncols, util.ptr(x), T.scalar(0.0))
        flops = 2 * A.nnz
    elif workload_name == "cfg1":
        from spgpu_b200 import formats as F, generators as G
        ell = F.coo_to_ell(G.laplace2d_5pt(1000))
        nnz = int(ell.rs.sum())
        sample = f"the full 1000x1000 5-point ELL matrix ({ell.nrows} rows, {nnz} nnz)"
        x = np.random.default_rng(12345).random(ell.ncols)
        z = np.zeros(ell.nrows)
        T = util.TYPES["D"]
        call = lambda: O.Dellspmv(util.ptr(z), None, T.scalar(1.0), util.ptr(ell.values), util.ptr(ell.indices), ell.pitch,
                                  ell.pitch, util.ptr(ell.rs), None, 4, ell.maxnnz, ell.nrows, util.ptr(x), T.scalar(0.0), 0)
        flops = 2 * nnz
    elif workload_name == "cfg2dia":
        A = DB.dia_stencil27(128, device=dev)
        sample = f"the full 128^3 27-point DIA matrix ({A.nrows} rows, {A.nnz} nnz)"
        vals, off = A.values.cpu().numpy(), A.offsets.cpu().numpy()
        x = np.random.default_rng(12345).random(A.ncols)
        z = np.zeros(A.nrows)
        T = util.TYPES["D"]
        call = lambda: O.Ddiaspmv(util.ptr(z), None, T.scalar(1.0), util.ptr(vals), util.ptr(off), A.pitch, A.nrows, A.ncols,
                                  A.diags, util.ptr(x), T.scalar(0.0))
        flops = 2 * A.nnz
    elif workload_name in ("cfg3", "cfg3o", "cfg4"):
        # a quarter of the rows, generated by the same builders (device when there is one) and copied to the host
        ridx = None
        if workload_name in ("cfg3", "cfg3o"):
            R, sym, dt, fl = 1 << 20, "S", np.float32, 2
            lens, cols, vals = DB.powerlaw_entries(R, device=dev)
            if workload_name == "cfg3o":
                lens, cols, vals, ridx = DB.sort_rows_by_length(lens, cols, vals)
                ridx = ridx.cpu().numpy()
            alpha, beta = 1.0, 0.0
        else:
            R, sym, dt, fl = 500_000, "Z", np.complex128, 8
            lens, cols, vals = DB.banded_complex_entries(R, device=dev)
            alpha, beta = 0.7 - 0.3j, -0.5 + 0.25j
        A = DB.hell_from_rows(lens, cols, vals, R)
        del lens, cols, vals
        sample = f"{R} rows generated like the workload's ({A.nnz} nnz), {'OHELL' if ridx is not None else 'HELL'} {sym}"
        hv, hi, ho, rs = (t.cpu().numpy() for t in (A.values, A.indices, A.hack_offsets, A.rs))
        hi = np.where(hi < 0, 0, hi).astype(np.int32)            # builders poison the padding indices; never read
        rng = np.random.default_rng(12345)
        x = rng.random(R).astype(dt) if sym == "S" else (rng.random(R) + 1j * rng.random(R))
        y = x.copy()
        z = np.zeros(R, dtype=dt)
        T = util.TYPES[sym]
        call = lambda: getattr(O, f"{sym}hellspmv")(util.ptr(z), util.ptr(y) if beta != 0 else None, T.scalar(alpha),
                                                   util.ptr(hv), util.ptr(hi), 32, util.ptr(ho), util.ptr(rs),
                                                   util.ptr(ridx) if ridx is not None else None,
                                                   A.avg, R, util.ptr(x), T.scalar(beta), 0)
        flops = fl * A.nnz
    else:
        return None
    for _ in range(max(1, warmup)):          # untimed (page faults, thread pool)
        call()
    times, spent = [], 0.0
    for _ in range(max(1, repeats)):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        times.append(dt)
        spent += dt
        if spent > budget_s:
            break
    mean = float(np.mean(times))
    return {"value": flops / mean / 1e9, "unit": "GFLOP/s", "cores": int(cores), "kind": "port",
            "best_value": flops / min(times) / 1e9,
            "sample": sample + f"; OpenMP schedule(static) over rows, {max(1, warmup)} warm-up + mean of {len(times)} timed calls"}, times


def cpu_baseline_mtx(w, budget_s=20.0, repeats=5):
    """The same OpenMP loops over the arrays of a --matrix workload (copied back to the host)."""
    from tests import util
    from spgpu_b200 import formats as F
    O = util.oracle_lib()
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    O.dll.oracle_set_num_threads(avail)
    A, kind = w["A"], w["kind"]
    np_dt = np.float32 if w["sym"] == "S" else np.float64
    host = lambda t: t.cpu().numpy()
    ridx = None
    if kind == "hell":
        idx = host(A.indices)
        M = F.Hell(host(A.values), np.where(idx < 0, 0, idx).astype(np.int32), host(A.hack_offsets), host(A.rs), 32, 0,
                   A.nrows, A.ncols, 0)
        ridx = host(A.ridx) if A.ridx is not None else None
    elif kind == "hdia":
        M = F.Hdia(host(A.values), host(A.offsets), host(A.hack_offsets), 32, int(A.offsets.numel()), A.nrows, A.ncols)
    elif kind == "dia":
        M = F.Dia(host(A.values), host(A.offsets), A.pitch, A.diags, A.nrows, A.ncols)
    else:
        M = A["ell"]
    x = np.random.default_rng(12345).random(w["x_len"]).astype(np_dt)
    kw = {"ridx": ridx} if ridx is not None else {}
    call = lambda: util.oracle_spmv(kind, M, x, None, 1.0, 0.0, **kw)
    call()
    best, spent = float("inf"), 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        best, spent = min(best, dt), spent + dt
        if spent > budget_s:
            break
    return {"value": 2 * w["nnz"] / best / 1e9, "unit": "GFLOP/s", "cores": int(O.dll.oracle_num_threads()), "kind": "port",
            "sample": f"the whole matrix ({w['rows']} rows, {w['nnz']} nnz), same {kind.upper()} arrays; OpenMP schedule(static) over rows, "
                      f"best of {repeats}"}


def reference_kernels_leg(w, x_ptr, y, z_ours, step_ours, stream, device, dev_index, flush, reps, ours_ms):
    """Times the reference library's own kernel for this workload's entry point on the SAME device buffers:
    one CUDA-event pair per call on the reference handle's stream (its default, blocking stream), cold L2 where
    our own timing flushes it.  The texture bind the reference issues per call is, under oracle/texshim.h, an
    8-byte pointer upload; SPGPU_REF_ASYNC_BIND=1 makes it asynchronous so the pair brackets the kernel and not
    a host synchronisation.  Also checks the reference's result against ours on these buffers."""
    import torch
    os.environ["SPGPU_REF_ASYNC_BIND"] = "1"          # read by the shim at its first bind
    from tests import util
    R = util.ref_lib()
    if R is None:
        return {"unavailable": "oracle/_ref/libspgpu_ref.so not built (needs /root/reference at build time)"}
    rh = ctypes.c_void_p()
    assert R.spgpuCreate(ctypes.byref(rh), dev_index) == 0
    rstream = torch.cuda.ExternalStream(R.spgpuGetStream(rh), device=device)
    z_ref = torch.full_like(z_ours, float("nan"))
    step_ref = make_step(R, rh, w, x_ptr, z_ref.data_ptr(), y.data_ptr() if y is not None else 0)
    torch.cuda.synchronize()
    for _ in range(2):
        step_ref()
    torch.cuda.synchronize()
    pairs = []
    for _ in range(reps):
        if flush is not None:
            flush()
            rstream.wait_stream(stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(rstream)
        step_ref()
        b.record(rstream)
        pairs.append((a, b))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
    step_ours()
    torch.cuda.synchronize()
    zr, zo = torch.view_as_real(z_ref) if z_ref.is_complex() else z_ref, torch.view_as_real(z_ours) if z_ours.is_complex() else z_ours
    diff = float((zr - zo).abs().max().item())
    scale = float(zo.abs().max().item())
    R.spgpuDestroy(rh)
    s, kind = w["sym"], w["kind"]
    return {"entry": f"spgpu{s}{kind}spmv", "library": "oracle/_ref/libspgpu_ref.so (the reference's unmodified kernels, nvcc sm_100a, "
            "texture fetches mapped to plain loads by oracle/texshim.h)",
            "kernel_ms": ms, "gflops": w["flops_per_nnz"] * w["nnz"] / (ms * 1e-3) / 1e9,
            "hbm_gbs": w["bytes"] / (ms * 1e-3) / 1e9, "ours_kernel_ms": ours_ms, "speedup_ours_vs_reference_kernels": ms / ours_ms,
            "launches_timed": reps, "l2": "cold (flushed before every call)" if flush is not None else "inputs larger than L2",
            "max_abs_diff_vs_ours": diff, "max_abs_ours": scale,
            "what": "one CUDA-event pair per call on the reference handle's stream, same device buffers as our kernel"}


def workload_label(name, size=None):
    """The `config.workload` string of a synthetic workload (the same in both arms)."""
    n = size
    return {
        "cfg5": f"3-D 7-point Laplacian {n or 512}^3, double HELL hackSize 32 (BASELINE configs[4])",
        "cfg2": f"3-D 27-point stencil {n or 128}^3, double HDIA hackSize 32 (BASELINE configs[1])",
        "cfg2dia": f"3-D 27-point stencil {n or 128}^3, double DIA (27 diagonals; BASELINE configs[1] stored as plain DIA)",
        "cfg1": f"2-D 5-point Laplacian {n or 1000}x{n or 1000}, double ELL with rS (BASELINE configs[0])",
        "cfg3": f"power-law rows {n or (1 << 22)} avg 16 max 4096, float HELL hackSize 32 (BASELINE configs[2])",
        "cfg3o": f"power-law rows {n or (1 << 22)} avg 16 max 4096, float OHELL (rows sorted by length + rIdx) hackSize 32",
        "cfg4": f"banded complex-double {n or 2_000_000} rows ~40 nnz/row, HELL hackSize 32 (BASELINE configs[3])",
    }[name]


def workload_config(name, size, world, halo, overlap, rows=None, nnz=None, flush_l2=None, alpha=None, beta=None):
    """The `config` object of the JSON line.  Both arms print the same one (the reference arm fills in the
    analytic row / nnz counts of the workloads it knows)."""
    if rows is None and name == "cfg5":
        n = size or 512
        rows, nnz = n ** 3, 7 * n ** 3 - 6 * n * n
    if rows is None and name in ("cfg2", "cfg2dia"):
        n = size or 128
        rows, nnz = n ** 3, (3 * n - 2) ** 3
    if rows is None and name == "cfg1":
        n = size or 1000
        rows, nnz = n * n, 5 * n * n - 4 * n
    if rows is None and name in ("cfg3", "cfg3o", "cfg4"):
        rows = size or ((1 << 22) if name != "cfg4" else 2_000_000)      # nnz is only known once generated
    if flush_l2 is None:
        flush_l2 = name in ("cfg1", "cfg2", "cfg2dia", "cfg3", "cfg3o")
    if alpha is None:
        alpha, beta = ((0.7 - 0.3j), (-0.5 + 0.25j)) if name == "cfg4" else (1.0, 0.0)
    cfg = {"workload": workload_label(name, size), "rows": rows, "nnz": nnz,
           "parallelism": f"row-sharded z-slabs x{world}, halo={halo if world > 1 else 'none'}"
                          f"{', interior/boundary overlap' if overlap and world > 1 else ''}",
           "l2": "L2 evicted between timed steps by reading a 512 MB scratch (clean lines), outside the event pairs" if flush_l2
                 else "inputs larger than L2 (>= 8x 126 MB per GPU), no flush",
           "alpha": str(alpha), "beta": str(beta)}
    return cfg


# --------------------------------------------------------------------------- #
# main
# --------------------------------------------------------------------------- #

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg1", "cfg2", "cfg2dia", "cfg3", "cfg3o", "cfg4", "cfg5"])
    ap.add_argument("--size", type=int, default=None, help="override the grid size / row count (testing)")
    ap.add_argument("--matrix", default=None, help="a MatrixMarket coordinate file to run instead of a synthetic workload (1 GPU)")
    ap.add_argument("--format", default="hell", choices=["ell", "hell", "ohell", "dia", "hdia"], help="storage format for --matrix")
    ap.add_argument("--precision", default="D", choices=["S", "D"], help="value type for --matrix")
    ap.add_argument("--halo", default="fused", choices=["fused", "push", "nccl", "allgather"],
                    help="multi-GPU halo exchange: fused = inside the SpMV kernel over NVLink peer pointers; "
                         "push = separate NVLink push kernel + flags; nccl = grouped send/recv; "
                         "allgather = global columns + NCCL all-gather of x (the mode for unstructured matrices)")
    ap.add_argument("--overlap", action="store_true",
                    help="multiply interior rows while the halos are in flight (3 SpMV launches per step); "
                         "default: one fused exchange kernel, then one SpMV launch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cg", dest="cg", action="store_true", default=None,
                    help="also time CG iterations (SpMV + 2 dots + 3 axpby); default: on for cfg5 (BASELINE configs[4] reads 'plus CG step')")
    ap.add_argument("--no-cg", dest="cg", action="store_false")
    ap.add_argument("--no-ref-kernels", action="store_true",
                    help="skip the reference_kernels leg (the reference's own kernels, oracle/_ref, on the same buffers; N=1, cfg5/cfg2)")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="CG scalars across GPUs: peer = one-double all-reduce over NVLink peer memory; nccl")
    ap.add_argument("--verify", dest="verify", action="store_true", default=None,
                    help="N>1: check the partitioned result bit-for-bit against the same rows multiplied with global columns "
                         "(outside the timed region; default: on)")
    ap.add_argument("--no-verify", dest="verify", action="store_false")
    ap.add_argument("--tune", default="", help="key=value,... passed to spgpuSetTuning")
    ap.add_argument("--sweep", default="", help="'k=v,k=v;k=v;...': kernel-only timing per tuning set (stderr)")
    args = ap.parse_args()
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    if args.cg is None:
        args.cg = args.workload == "cfg5" and not args.matrix
    if args.verify is None:
        args.verify = True

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # ---------------- reference arm: host OpenMP on rank 0 only -----------------
    # W untimed + exactly K timed multiplications of the workload on the host cores; the line carries
    # the same `config` as our arm's (workload_config) and says in cpu_baseline.sample what was multiplied.
    if args.impl == "reference":
        if rank != 0:
            return 0
        arm = args.workload
        cb, times = cpu_baseline(arm, budget_s=120.0, repeats=K, warmup=W, size=args.size)
        mean = float(np.mean(times))
        out = {"impl": "reference", "metric": "spmv_gflops", "value": cb["value"], "unit": "GFLOP/s",
               "n_gpus": args.gpus, "steps": len(times), "warmup": W, "ms_per_step": mean * 1e3,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": {"cfg3": "f32", "cfg3o": "f32", "cfg4": "c128"}.get(arm, "f64"),
               "data": "synthetic", "config": workload_config(arm, args.size, args.gpus, args.halo, args.overlap),
               "arm": "host OpenMP SpMV (oracle port of the reference's kernel loop) over the same format on rank 0's host "
                      "cores; the reference ships no CPU SpMV of its own",
               "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out), flush=True)
        return 0

    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version
    # banner ...) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from spgpu_b200 import capi, mg

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # the first query on a fresh box can take over a second: start long before the timed region
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.workload not in ("cfg5", "cfg2", "cfg4") and world > 1:
        raise SystemExit("cfg5 (double HELL), cfg2 (double HDIA) and cfg4 (complex-double HELL) are row-sharded across GPUs; "
                         "run the other workloads with --gpus 1")
    if args.workload in ("cfg2", "cfg4") and world > 1 and args.halo == "allgather":
        args.halo = "fused"             # HDIA addresses x relative to the row; cfg4 is only built with local columns

    L = capi.lib()
    h = ctypes.c_void_p()
    st = L.spgpuCreate(ctypes.byref(h), local_rank)
    assert st == 0, f"spgpuCreate -> {st}"
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        assert L.spgpuSetTuning(h, k.encode(), int(v)) == 0, f"unknown tuning key {k}"

    # everything (torch ops, NCCL, our kernels) is ordered on ONE stream.  It is a stream torch owns, handed
    # to the handle with spgpuSetStream (reference core.c:62-72): torch's allocators remember the stream a
    # block was used on, so the stream has to outlive the handle for the interpreter to exit normally.
    stream = torch.cuda.Stream(device=device)
    L.spgpuSetStream(h, stream.cuda_stream)
    torch.cuda.set_stream(stream)

    if args.matrix:
        assert world == 1, "--matrix runs on one GPU"
        w = build_mtx_workload(args.matrix, args.format, L, h, device, args.precision)
        args.workload = "mtx"
    else:
        w = build_workload(args.workload, rank, world, device, args.size, global_columns=(args.halo == "allgather"))
    tdt = {"S": torch.float32, "D": torch.float64, "C": torch.complex64, "Z": torch.complex128}[w["sym"]]
    rows, halo = w["rows"], w["halo"]
    ext_len = w["x_len"]

    # x_ext lives in its own cudaMalloc block so it can be IPC-exported to the neighbours
    gen = torch.Generator(device=device)
    gen.manual_seed(12345 + rank)
    real = torch.float32 if w["sym"] in "SC" else torch.float64
    if w["sym"] in "CZ":
        x_init = torch.complex(torch.rand(ext_len, generator=gen, device=device, dtype=real),
                               torch.rand(ext_len, generator=gen, device=device, dtype=real))
    else:
        x_init = torch.rand(ext_len, generator=gen, device=device, dtype=real)
    if w["sym"] == "D" or world > 1:
        x_ptr, x_ext = mg.raw_device_vector(L, ext_len, tdt)
        x_ext.copy_(x_init)
    else:
        x_ext = x_init
        x_ptr = x_ext.data_ptr()
    del x_init
    z = torch.zeros(rows, dtype=tdt, device=device)
    y = None
    if w["beta"] != 0:
        y = torch.rand(rows, generator=gen, device=device, dtype=torch.float64).to(tdt)
    step = make_step(L, h, w, x_ptr, z.data_ptr(), y.data_ptr() if y is not None else 0)

    peer = None
    ex = mg.HaloExchange(rank, world, halo, "nccl")
    if world > 1 and args.halo in ("push", "fused"):
        peer = mg.PeerHalo(L, h, rank, world, x_ptr, ext_len, halo, itemsize=w["sizeof"])

    def make_fused(z_ptr_):
        if peer is None or args.halo != "fused" or w["kind"] not in ("hell", "hdia"):
            return None
        A, s = w["A"], w["sym"]
        t = capi.TYPES[s]
        one, zero = t.scalar(w["alpha"]), t.scalar(w["beta"])
        y_ptr_ = y.data_ptr() if y is not None else 0
        links = peer.links_ref()
        if w["kind"] == "hdia":
            fn = getattr(L, f"spgpu{s}hdiaspmvHalo")

            def fused_hdia(seq):
                fn(h, z_ptr_, y_ptr_, one, A.values.data_ptr(), A.offsets.data_ptr(), A.hack_size,
                   A.hack_offsets.data_ptr(), rows, A.ncols, x_ptr, zero, halo, links, seq)
            return fused_hdia
        fn = getattr(L, f"spgpu{s}hellspmvHalo")

        def fused(seq):
            fn(h, z_ptr_, y_ptr_, one, A.values.data_ptr(), A.indices.data_ptr(), A.hack_size,
               A.hack_offsets.data_ptr(), A.rs.data_ptr(), A.avg, rows, x_ptr, zero, A.base, halo, links, seq)
        return fused

    op = mg.MgHellSpmv(rank, world, rows, halo, lambda _z, _x, r0, r1: step(r0, r1), ex, peer,
                       overlap=args.overlap, fused_spmv=make_fused(z.data_ptr()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if w.get("allgather"):
        # unstructured-matrix mode: columns are global, every rank gathers the whole x per SpMV
        x_owned = x_ext[rank * rows:(rank + 1) * rows].clone()
        ag = mg.MgAllGatherSpmv(world, rows, lambda _z, _x: step(0, rows))

        def one_step():
            ag.apply(z, x_owned, x_ext)
    else:
        def one_step():
            op.apply(z, x_ext)

    # working sets that are not >> L2 (126 MB): evict it between timed steps by READING a 512 MB
    # scratch (a write-flush would leave 126 MB of dirty lines whose write-back lands inside the
    # timed kernel: +19 us on kernels that take 12-80 us)
    flush_l2 = w["bytes"] < 8 * 126e6          # w["bytes"] is what THIS rank's kernel touches per launch
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.int64, device=device) if flush_l2 else None

    def flush():
        scratch.sum()

    def timed_steps(fn, count):
        """total ms of `count` calls of fn on the handle's stream (CUDA events); with
        flush_l2 every call is bracketed by its own event pair and the flush sits outside"""
        if not flush_l2:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(count):
                fn()
            b.record(stream)
            return a, b, None
        pairs = []
        for _ in range(count):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            pairs.append((a, b))
        return None, None, pairs

    def elapsed(a, b, pairs):
        if pairs is None:
            return a.elapsed_time(b)
        per_step = [x.elapsed_time(y) for x, y in pairs]
        if os.environ.get("SPGPU_BENCH_TRACE"):              # per-step times of this rank, for diagnosing rank skew
            print(f"rank {rank} step ms: " + " ".join(f"{t:.4f}" for t in per_step), file=sys.stderr, flush=True)
        return float(sum(per_step))

    # ---------------- optional multi-GPU self-check ------------------------------
    verified = None
    if args.verify and args.workload == "cfg5" and world > 1 and not w.get("allgather"):
        from spgpu_b200 import device_build as DB
        n = args.size or 512
        per = n // world
        one_step()
        barrier()
        owned = x_ext[halo:halo + rows].contiguous()
        x_full = torch.empty(world * rows, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(x_full, owned)
        Ag = DB.hell_laplace3d_7pt(n, rank * per, (rank + 1) * per, local_columns=False, device=device)
        z_ref = torch.full((rows,), float("nan"), dtype=torch.float64, device=device)
        T = capi.TYPES["D"]
        L.spgpuDhellspmv(h, z_ref.data_ptr(), 0, T.scalar(1.0), Ag.values.data_ptr(), Ag.indices.data_ptr(), 32,
                         Ag.hack_offsets.data_ptr(), Ag.rs.data_ptr(), 0, Ag.avg, rows, x_full.data_ptr(),
                         T.scalar(0.0), 0)
        torch.cuda.synchronize()
        ok = torch.tensor([1.0 if torch.equal(z, z_ref) else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verified = bool(ok.item() == 1.0)
        del Ag, z_ref, x_full
        if rank == 0:
            print(f"verify: partitioned SpMV == global-column SpMV on every rank: {verified}", file=sys.stderr, flush=True)

    if args.verify and world > 1 and w.get("A_global") is not None:
        # each rank multiplies ITS rows with global columns by the all-gathered x and compares bit for bit
        Ag, sym_ = w["A_global"], w["sym"]
        one_step()
        barrier()
        own = x_ext[halo:halo + rows].contiguous()
        parts = [torch.empty(hi_ - lo_, dtype=tdt, device=device) for lo_, hi_ in mg.row_blocks(w["total_rows"], world, 32)]
        if tdt.is_complex:
            dist.all_gather([torch.view_as_real(p_) for p_ in parts], torch.view_as_real(own))
        else:
            dist.all_gather(parts, own)
        x_full = torch.cat(parts)
        z_ref = torch.full((rows,), float("nan"), dtype=tdt, device=device)
        T = capi.TYPES[sym_]
        getattr(L, f"spgpu{sym_}hellspmv")(h, z_ref.data_ptr(), y.data_ptr() if y is not None else 0, T.scalar(w["alpha"]),
                                          Ag.values.data_ptr(), Ag.indices.data_ptr(), 32, Ag.hack_offsets.data_ptr(),
                                          Ag.rs.data_ptr(), 0, Ag.avg, rows, x_full.data_ptr(), T.scalar(w["beta"]), 0)
        torch.cuda.synchronize()
        same = torch.equal(torch.view_as_real(z), torch.view_as_real(z_ref)) if tdt.is_complex else torch.equal(z, z_ref)
        ok = torch.tensor([1.0 if same else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verified = bool(ok.item() == 1.0)
        del z_ref, x_full, parts
        if rank == 0:
            print(f"verify: partitioned {sym_} HELL SpMV == the same rows with global columns on every rank: {verified}", file=sys.stderr, flush=True)
    w.pop("A_global", None)

    if args.verify and args.workload == "cfg2" and world > 1:
        # each rank multiplies the FULL matrix by the all-gathered x and compares its own rows
        Af = w["full"]
        one_step()
        barrier()
        own = x_ext[halo:halo + rows].contiguous()
        parts = [torch.empty(hi_ - lo_, dtype=torch.float64, device=device) for lo_, hi_ in mg.row_blocks(Af.nrows, world, 32)]
        dist.all_gather(parts, own)
        x_full = torch.cat(parts)
        z_ref = torch.full((Af.nrows,), float("nan"), dtype=torch.float64, device=device)
        T = capi.TYPES["D"]
        L.spgpuDhdiaspmv(h, z_ref.data_ptr(), 0, T.scalar(1.0), Af.values.data_ptr(), Af.offsets.data_ptr(), 32,
                         Af.hack_offsets.data_ptr(), Af.nrows, Af.ncols, x_full.data_ptr(), T.scalar(0.0))
        torch.cuda.synchronize()
        ok = torch.tensor([1.0 if torch.equal(z, z_ref[w["lo"]:w["hi"]]) else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verified = bool(ok.item() == 1.0)
        del z_ref, x_full, parts
        if rank == 0:
            print(f"verify: partitioned HDIA SpMV == full-matrix SpMV rows on every rank: {verified}", file=sys.stderr, flush=True)
    w.pop("full", None)

    # ---------------- device-resident timing ------------------------------------
    halo_trace = bool(os.environ.get("SPGPU_BENCH_TRACE")) and peer is not None and args.halo == "fused"
    if halo_trace:
        L.spgpuSetTuning(h, b"haloTrace", 1)
    if rank == 0:
        sampler.mark()                  # lines printed from here on are "during the measurement"
    for _ in range(W):
        one_step()
    barrier()
    # The host-side barrier lets the ranks go within tens of microseconds of each other; with 20 steps of
    # 0.3 ms that skew is a visible part of the first step (a rank's boundary rows wait for a late neighbour's
    # halo).  A one-double all-reduce over NVLink peer memory right behind the barrier, in stream order before
    # the first event, lines the DEVICES up to a few microseconds; the timed region is still exactly K steps
    # between two barriers.
    rendezvous = None
    if world > 1 and peer is not None:
        try:
            rendezvous = mg.PeerAllreduce(L, h, rank, world)
            sync_val = torch.zeros(1, dtype=torch.float64, device=device)
            rendezvous(sync_val)                        # first use outside the timed region
        except Exception as exc:
            print(f"device rendezvous unavailable: {exc!r}", file=sys.stderr, flush=True)
            rendezvous = None
    barrier()
    if rendezvous is not None:
        rendezvous(sync_val)
    launches0 = L.spgpuGetLaunchCount(h)
    e0, e1, pairs = timed_steps(one_step, K)
    barrier()
    ms_total = elapsed(e0, e1, pairs)
    launches = L.spgpuGetLaunchCount(h) - launches0
    if halo_trace:
        # per-exchange trace of the fused kernel for the K timed steps (include/spgpu_ext.h: spgpuHaloTraceRead)
        buf = (ctypes.c_ulonglong * (8 * K))()
        if L.spgpuHaloTraceRead(h, buf, peer.fseq - K + 1, K) == 0:
            tr = np.frombuffer(buf, dtype=np.uint64).reshape(K, 8).astype(np.int64)
            t0 = tr[0, 0]
            for k in range(K):
                r_ = tr[k]
                print(f"rank {rank} seq {peer.fseq - K + 1 + k}: start {(r_[0] - t0) / 1e3:9.1f} us  push {(r_[1] - r_[0]) / 1e3:6.1f} us  "
                      f"boundary first..last {(r_[6] - r_[0]) / 1e3:7.1f}..{(r_[7] - r_[0]) / 1e3:7.1f} us  "
                      f"waited lo {r_[2] / 1e3:7.1f} us in {r_[4]} blocks, hi {r_[3] / 1e3:7.1f} us in {r_[5]} blocks",
                      file=sys.stderr, flush=True)
        L.spgpuSetTuning(h, b"haloTrace", 0)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / K

    nnz_total = torch.tensor([float(w["nnz"])], dtype=torch.float64, device=device)
    bytes_total = torch.tensor([float(w["bytes"])], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(nnz_total)
        dist.all_reduce(bytes_total)
    nnz_total, bytes_total = float(nnz_total.item()), float(bytes_total.item())
    gflops = w["flops_per_nnz"] * nnz_total / (ms_step * 1e-3) / 1e9

    # ---------------- the dominant kernel alone (roofline) ----------------------
    # per-launch duration of the full-block SpMV kernel on this rank's stream
    # (all launches are queued before the one synchronise: a host-side delay between an event record and
    # the launch behind it would otherwise be counted as kernel time)
    step()                              # untimed: the first launch after the collectives above pays their teardown
    ker_pairs = []
    for _ in range(min(K, 10)):
        if flush_l2:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step()
        b.record(stream)
        ker_pairs.append((a, b))
    torch.cuda.synchronize()
    ker_ms = float(np.mean([a.elapsed_time(b) for a, b in ker_pairs]))
    # ... and back to back, the way the timed steps run (no event between launches): what a partitioned step has to
    # be compared with to see what the exchange costs
    ker_b2b_ms = None
    if not flush_l2:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(K):
            step()
        b.record(stream)
        torch.cuda.synchronize()
        tb = torch.tensor([a.elapsed_time(b) / K], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        ker_b2b_ms = float(tb.item())
    # the sampler has been running since before the warm-up: it covers the timed steps AND the kernel-alone launches
    # the roofline is computed from (a 40 ms timed region alone is one or two nvidia-smi samples)
    def keep_busy():
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop(keep_busy) if rank == 0 else None
    peak, peak_src = measured_peak()
    achieved = w["bytes"] / (ker_ms * 1e-3) / 1e9
    traffic, limiter = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get(f"{args.workload}_n{world}")
        limiter = (tj.get("_limiters") or {}).get(f"{args.workload}_n{world}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": w.get("kernel", f"{w['kind']}_spmv_kernel<{w['sym']}>"),
                "kernel_ms": ker_ms, "algorithmic_bytes_per_launch": w["bytes"],
                # the same kernel launched K times back to back WITHOUT the exchange (max over ranks): the partitioned
                # step minus this is what the halo exchange costs
                "kernel_back_to_back_ms": ker_b2b_ms}
    if limiter:
        roofline["limiter_ncu"] = limiter          # the kernel is not DRAM-bound: what ncu says it is bound by
    if flush_l2 and world == 1:
        # SHORT problems (cfg1: 80 MB = 12 us at the long-copy peak): launch ramp, the last partial wave and the cold L2 are a
        # visible part of the run whatever the kernel does.  The practical ceiling is a pure stream of the SAME number of
        # bytes under the SAME protocol (flush, one event pair per launch): spgpuDscal, n doubles read + n written.
        try:
            ns = max(1, int(w["bytes"] // 16))
            sx = torch.rand(ns, dtype=torch.float64, device=device)
            sz = torch.empty(ns, dtype=torch.float64, device=device)
            one_and_half = capi.TYPES["D"].scalar(1.5)
            ts = []
            for it in range(13):
                flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                L.spgpuDscal(h, sz.data_ptr(), ns, one_and_half, sx.data_ptr())
                b.record(stream)
                b.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(b))
            stream_ms = float(np.mean(ts))
            roofline["same_size_stream"] = {"op": "spgpuDscal, the same bytes, the same protocol", "bytes": 16 * ns, "ms": stream_ms,
                                            "frac_of_peak": 16 * ns / (stream_ms * 1e-3) / 1e9 / peak,
                                            "kernel_vs_stream": stream_ms * (w["bytes"] / (16 * ns)) / ker_ms}
            del sx, sz
        except Exception as exc:
            print(f"same-size stream not measured: {exc!r}", file=sys.stderr, flush=True)
    if w["kind"] == "hell":
        # what the FORMAT costs at the fetch granularities the hardware has: matrix bytes only
        try:
            roofline["matrix_floor_bytes_by_fetch_granularity"] = hell_matrix_floor_bytes(w["A"], w["sizeof"])
            roofline["matrix_algorithmic_bytes"] = int(w["nnz"]) * (w["sizeof"] + 4)
        except Exception as exc:
            print(f"matrix floor not computed: {exc!r}", file=sys.stderr, flush=True)

    # ---------------- the REFERENCE'S OWN kernels on the same buffers (N=1) ------
    # SURVEY 8(d) "reference on B200" column: oracle/_ref/libspgpu_ref.so = the reference's unmodified sources
    # compiled for sm_100a (test infrastructure; here only as a timed-beside baseline, outside the timed region,
    # protocol of reference hellPerf.cpp:236-252 with CUDA events instead of a wall clock).
    ref_kernels = None
    if world == 1 and rank == 0 and not args.no_ref_kernels and not args.matrix and w["kind"] in ("hell", "hdia", "ell", "dia"):
        try:
            ref_kernels = reference_kernels_leg(w, x_ptr, y, z, step, stream, device, local_rank, flush if flush_l2 else None,
                                                min(K, 10), ker_ms)
        except Exception as exc:
            print(f"reference_kernels leg failed: {exc!r}", file=sys.stderr, flush=True)
            ref_kernels = {"unavailable": repr(exc)}

    # ---------------- optional tuning sweep (kernel only, stderr) ---------------
    for setting in filter(None, args.sweep.split(";")):
        for kv in filter(None, setting.split(",")):
            k, v = kv.split("=")
            assert L.spgpuSetTuning(h, k.encode(), int(v)) == 0, f"unknown tuning key {k}"
        ts = []
        for _ in range(8):
            if flush_l2:
                flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); step(); b.record(stream); b.synchronize()
            ts.append(a.elapsed_time(b))
        ts = ts[2:]
        print(json.dumps({"sweep": setting, "workload": args.workload, "kernel_ms": float(np.mean(ts)),
                          "min_ms": float(np.min(ts)), "gbs": w["bytes"] / (np.mean(ts) * 1e-3) / 1e9,
                          "frac": w["bytes"] / (np.mean(ts) * 1e-3) / 1e9 / peak}), file=sys.stderr, flush=True)
    if args.sweep:
        for kv in filter(None, args.tune.split(",")):          # restore the requested tuning
            k, v = kv.split("=")
            L.spgpuSetTuning(h, k.encode(), int(v))

    # ---------------- end to end: x from pinned host, z back to pinned host -----
    # Every step copies the step's input vector from pinned host memory, multiplies through the
    # C ABI and copies the result back.  On one GPU with a banded matrix (rows r0..r1 only need
    # x[r0-bw .. r1+bw]) the three stages are pipelined over row chunks on three streams --
    # sub-range SpMV calls with offset pointers, the same device the reference's own large-vector
    # loop uses (hell_spmv_base.cuh:121-137) -- so PCIe runs in both directions at once.
    e2e = None
    if not args.no_e2e:
        own = x_owned if w.get("allgather") else (x_ext[halo:halo + rows] if halo else x_ext)
        hx = torch.empty(own.shape, dtype=own.dtype, pin_memory=True)
        hx.copy_(own)
        hz = torch.empty(z.shape, dtype=z.dtype, pin_memory=True)
        Ke = max(2, min(K, 5))
        bw = w.get("bandwidth")
        pipelined = world == 1 and w["kind"] in ("hell", "hdia") and bw is not None and rows >= (1 << 20)
        if pipelined:
            # at most 32 chunks, at least ~2 MB of x each (below that the per-chunk launches cost more than they hide)
            nchunk = int(min(32, max(4, hx.numel() * hx.element_size() // (2 << 20))))
            unit = 32 * 1024                                   # chunk boundaries on hack boundaries
            csz = -(-rows // nchunk // unit) * unit
            assert csz >= bw
            bounds = [(c * csz, min(rows, (c + 1) * csz)) for c in range(nchunk) if c * csz < rows]
            s_in, s_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
            prev_done = torch.cuda.Event()
            prev_done.record(stream)
            own_p, step_p = own, step
            if w["kind"] == "hdia":
                # HDIA addresses x relative to the row and the ABI has no row base: a sub-range call sees its rows
                # as 0.., so diagonals reaching back before the chunk would fail the kernel's own range test.  The
                # pipeline therefore runs on the row-block form of the SAME matrix (offsets + halo on
                # x = [halo zeros | x | halo zeros], what mg.split_hdia produces for one rank): same cells, same
                # order of operations, checked bit-equal to the one-shot product below.
                from spgpu_b200 import device_build as DB
                halo_p = -(-bw // 32) * 32
                Ab = DB.hdia_row_block(w["A"], 0, rows, halo_p)
                x_p = torch.zeros(rows + 2 * halo_p, dtype=own.dtype, device=device)
                own_p = x_p[halo_p:halo_p + rows]
                step_p = make_step(L, h, dict(w, A=Ab), x_p.data_ptr(), z.data_ptr(), y.data_ptr() if y is not None else 0)

            def e2e_step():
                s_in.wait_event(prev_done)                     # x may be overwritten once the last SpMV is done
                arrived = []
                with torch.cuda.stream(s_in):
                    for (c0, c1) in bounds:
                        own_p[c0:c1].copy_(hx[c0:c1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(s_in)
                        arrived.append(ev)
                for c, (r0, r1) in enumerate(bounds):
                    stream.wait_event(arrived[min(c + 1, len(bounds) - 1)])   # needs x chunks c-1, c, c+1
                    step_p(r0, r1)
                    done = torch.cuda.Event()
                    done.record(stream)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(done)
                        hz[r0:r1].copy_(z[r0:r1], non_blocking=True)
                prev_done.record(stream)
                stream.wait_stream(s_out)                      # the step ends when z is on the host
        elif world > 1 and peer is not None and w["kind"] == "hell" and bw is not None and rows >= 16 * bw:
            # the same pipeline on a partition: the two boundary chunks of x are copied first, the halo
            # planes are exchanged over NVLink as soon as they are on the device (spgpuDhaloExchange),
            # interior chunks are multiplied as they arrive, z chunks leave on a third stream
            pipelined = True
            nchunk = 8
            unit = 32 * 1024
            csz = max(-(-rows // nchunk // unit) * unit, bw)
            bounds = [(c * csz, min(rows, (c + 1) * csz)) for c in range(nchunk) if c * csz < rows]
            last = len(bounds) - 1
            order = [0, last] + list(range(1, last)) if last > 0 else [0]
            s_in, s_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
            prev_done = torch.cuda.Event()
            prev_done.record(stream)

            def e2e_step():
                s_in.wait_event(prev_done)
                arrived = {}
                with torch.cuda.stream(s_in):
                    for c in order:
                        c0, c1 = bounds[c]
                        own[c0:c1].copy_(hx[c0:c1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(s_in)
                        arrived[c] = ev
                stream.wait_event(arrived[0])
                stream.wait_event(arrived[last])
                peer.exchange_fused()                      # push my boundary planes, wait for the neighbours'
                for c in order:
                    for nb in (c - 1, c, c + 1):
                        if 0 <= nb <= last:
                            stream.wait_event(arrived[nb])
                    r0, r1 = bounds[c]
                    step(r0, r1)
                    done = torch.cuda.Event()
                    done.record(stream)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(done)
                        hz[r0:r1].copy_(z[r0:r1], non_blocking=True)
                peer.ack_fused()
                prev_done.record(stream)
                stream.wait_stream(s_out)
        else:
            def e2e_step():
                own.copy_(hx, non_blocking=True)
                one_step()
                hz.copy_(z, non_blocking=True)
        for _ in range(2):
            e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(Ke):
            e2e_step()
        b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item()) / Ke
        # the floor of this step: the same two transfers with no SpMV in between, both directions at once
        t_in, t_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        xfer_ms = None
        try:
            def xfer():
                t_in.wait_stream(stream); t_out.wait_stream(stream)
                with torch.cuda.stream(t_in):
                    own.copy_(hx, non_blocking=True)
                with torch.cuda.stream(t_out):
                    hz.copy_(z, non_blocking=True)
                stream.wait_stream(t_in); stream.wait_stream(t_out)
            xfer()
            barrier()
            a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a2.record(stream)
            for _ in range(Ke):
                xfer()
            b2.record(stream)
            barrier()
            t2 = torch.tensor([a2.elapsed_time(b2)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            xfer_ms = float(t2.item()) / Ke
        except Exception as exc:
            print(f"transfer-only leg failed: {exc!r}", file=sys.stderr, flush=True)
        e2e_ok = None
        if pipelined:                                          # the pipelined result must be the plain one
            one_step()
            torch.cuda.synchronize()
            e2e_ok = bool(torch.equal(hz.to(device), z))
        e2e = {"value": w["flops_per_nnz"] * nnz_total / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": int(hx.numel() * hx.element_size()) * world,
               "d2h_bytes_per_step": int(hz.numel() * hz.element_size()) * world,
               "ms_per_step": ms_e2e, "steps": Ke,
               "transfers_only_ms": xfer_ms,     # the same H2D + D2H copies, concurrently, without the SpMV: the PCIe floor
               "what": ("x H2D from pinned memory, SpMV through the C ABI, z D2H to pinned memory, every step; matrix "
                        "resident" + ("; the three stages pipelined over row chunks on three streams "
                                      "(banded matrix; on a partition the boundary chunks go first and the halo "
                                      "planes are exchanged over NVLink as soon as they are on the device), result "
                                      "checked equal to the one-shot SpMV" if pipelined else ""))}
        if e2e_ok is not None:
            e2e["pipelined_result_equals_one_shot"] = e2e_ok

    # ---------------- CG step: SpMV + 2 dots + 3 axpby (BASELINE configs[4]) ------
    cg_out = None
    if args.cg and w["kind"] == "hell" and w["sym"] == "D":
        from spgpu_b200 import krylov
        A = w["A"]
        st = krylov.CgState(rows, halo, device, p_ext=x_ext)
        step_cg = make_step(L, h, w, x_ptr, st.ap.data_ptr(), 0)
        op_cg = mg.MgHellSpmv(rank, world, rows, halo, lambda _z, _x, r0, r1: step_cg(r0, r1), ex, peer,
                              overlap=args.overlap)

        def apply_A(_z, _x):
            op_cg.apply(st.ap, x_ext)

        if world == 1:
            def apply_A_dot(_z, _x, dres, _ar):
                L.spgpuDhellspmvDot(h, st.ap.data_ptr(), A.values.data_ptr(), A.indices.data_ptr(), A.hack_size,
                                    A.hack_offsets.data_ptr(), A.rs.data_ptr(), rows, x_ptr, A.base, 0, dres)
        elif peer is not None and args.halo == "fused":
            links = peer.links_ref()

            def apply_A_dot(_z, _x, dres, ar):  # halo exchange + SpMV + this rank's p.Ap in one kernel (+ fold + all-reduce)
                L.spgpuDhellspmvHaloDot(h, st.ap.data_ptr(), A.values.data_ptr(), A.indices.data_ptr(), A.hack_size,
                                        A.hack_offsets.data_ptr(), A.rs.data_ptr(), A.avg, rows, x_ptr, A.base, halo,
                                        links, peer.next_seq(), dres, ar)
        else:
            apply_A_dot = None

        peer_ar = None
        if world > 1 and args.allreduce == "peer":
            peer_ar = mg.PeerAllreduce(L, h, rank, world)
            allred = peer_ar
        else:
            allred = (lambda t: dist.all_reduce(t)) if world > 1 else None
        # the scalars of the device flavour are all-reduced by the last CTA of the kernel that produces them
        # (no separate launch) when the peer all-reduce is in use
        cg = krylov.Cg(L, h, st, apply_A, apply_A_dot, allred, ar_fused=peer_ar if apply_A_dot is not None else None,
                       pingpong=True)
        bvec = torch.rand(rows, generator=gen, device=device, dtype=torch.float64)
        iters = 10
        cg_out = {"iterations": iters}
        for flavour, fn in (("blocking", cg.step_blocking), ("device", cg.step_device)):
            cg.start(bvec)
            for _ in range(2):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = L.spgpuGetLaunchCount(h)
            a.record(stream)
            for _ in range(iters):
                fn()
            b.record(stream)
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_it = float(t.item()) / iters
            vec_bytes = krylov.CG_BYTES_PER_ROW_VECTOR_OPS[flavour] * w["total_rows"]
            if flavour == "device" and apply_A_dot is None:
                vec_bytes += 8 * w["total_rows"]          # separate p.Ap pass re-reads Ap
            cg_out[flavour] = {"ms_per_iteration": ms_it,
                               "algorithmic_gb_per_iteration": (bytes_total + vec_bytes) / 1e9,
                               "hbm_gbs": (bytes_total + vec_bytes) / (ms_it * 1e-3) / 1e9,
                               "frac_of_peak": (bytes_total + vec_bytes) / (ms_it * 1e-3) / 1e9 / (peak * world),
                               "kernels_per_iteration": (L.spgpuGetLaunchCount(h) - l0) / iters,
                               "residual_norm2_after": cg.residual_norm2() if flavour == "device" else st.rr_host}
        if os.environ.get("SPGPU_BENCH_TRACE"):
            # per-phase device time of the device flavour on this rank (events between the phases; a phase that
            # waits for a peer -- halo, all-reduce -- shows the wait)
            phases = {}
            for _ in range(6):
                evs = [("start", torch.cuda.Event(enable_timing=True))]
                evs[0][1].record(stream)

                def mark(label):
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(stream)
                    evs.append((label, e))
                cg.step_device(mark)
                torch.cuda.synchronize()
                for (_, a_), (lab, b_) in zip(evs[:-1], evs[1:]):
                    phases.setdefault(lab, []).append(a_.elapsed_time(b_))
            print(f"rank {rank} CG phases (ms, median of 6): " +
                  "; ".join(f"{k}: {float(np.median(v)):.4f}" for k, v in phases.items()), file=sys.stderr, flush=True)
            if peer is not None and args.halo == "fused":
                # the fused kernel's own record of its last 4 exchanges inside the iteration
                L.spgpuSetTuning(h, b"haloTrace", 1)
                for _ in range(4):
                    cg.step_device()
                buf = (ctypes.c_ulonglong * (8 * 4))()
                if L.spgpuHaloTraceRead(h, buf, peer.fseq - 3, 4) == 0:
                    tr = np.frombuffer(buf, dtype=np.uint64).reshape(4, 8).astype(np.int64)
                    for k in range(4):
                        r_ = tr[k]
                        print(f"rank {rank} CG exchange {peer.fseq - 3 + k}: start {(r_[0] - tr[0, 0]) / 1e3:9.1f} us  push {(r_[1] - r_[0]) / 1e3:6.1f} us  "
                              f"boundary first..last {(r_[6] - r_[0]) / 1e3:7.1f}..{(r_[7] - r_[0]) / 1e3:7.1f} us  "
                              f"waited lo {r_[2] / 1e3:7.1f} us in {r_[4]} blocks, hi {r_[3] / 1e3:7.1f} us in {r_[5]} blocks",
                              file=sys.stderr, flush=True)
                L.spgpuSetTuning(h, b"haloTrace", 0)
        # ---- the device-scalar iteration captured ONCE in a CUDA graph and replayed ---------------
        # (at N > 1 the halo / all-reduce sequence numbers then have to live in device memory:
        #  spgpuSetSeqCounters, include/spgpu_ext.h)
        graph_ok = apply_A_dot is not None and (world == 1 or (peer is not None and args.halo == "fused" and peer_ar is not None))
        if graph_ok:
            try:
                counters = torch.zeros(2, dtype=torch.int32, device=device)
                apply_eager = apply_A_dot
                if world > 1:
                    assert L.spgpuSetSeqCounters(h, counters.data_ptr(), counters.data_ptr() + 4) == 0
                    peer.to_device_seq(counters[0:1])
                    peer_ar.to_device_seq(counters[1:2])

                    def apply_A_dot_dev(_z, _x, dres, ar):
                        L.spgpuDhellspmvHaloDot(h, st.ap.data_ptr(), A.values.data_ptr(), A.indices.data_ptr(), A.hack_size,
                                                A.hack_offsets.data_ptr(), A.rs.data_ptr(), A.avg, rows, x_ptr, A.base, halo,
                                                links, 0, dres, ar)
                        L.spgpuHaloSeqAdvance(h)
                    cg.apply_A_dot = apply_A_dot_dev
                cg.pingpong = False                      # ONE captured iteration is replayed: its scalar slots are frozen
                cg.start(bvec)
                for _ in range(2):
                    cg.step_device()
                barrier()
                g = torch.cuda.CUDAGraph()
                cap = torch.cuda.Stream(device=device)
                L.spgpuSetStream(h, cap.cuda_stream)
                try:
                    with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                        cg.step_device()
                finally:
                    L.spgpuSetStream(h, stream.cuda_stream)
                torch.cuda.synchronize()
                for _ in range(2):
                    g.replay()
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cur = torch.cuda.current_stream(device)
                a.record(cur)
                for _ in range(iters):
                    g.replay()
                b.record(cur)
                barrier()
                t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_it = float(t.item()) / iters
                vec_bytes = krylov.CG_BYTES_PER_ROW_VECTOR_OPS["device"] * w["total_rows"]
                cg_out["graph"] = {"ms_per_iteration": ms_it,
                                   "algorithmic_gb_per_iteration": (bytes_total + vec_bytes) / 1e9,
                                   "hbm_gbs": (bytes_total + vec_bytes) / (ms_it * 1e-3) / 1e9,
                                   "frac_of_peak": (bytes_total + vec_bytes) / (ms_it * 1e-3) / 1e9 / (peak * world),
                                   "residual_norm2_after": cg.residual_norm2(),
                                   "what": "the device flavour captured once in a CUDA graph and replayed"}
                del g
                if world > 1:
                    peer.from_device_seq(counters[0:1])
                    peer_ar.from_device_seq(counters[1:2])
                    L.spgpuSetSeqCounters(h, 0, 0)
                cg.apply_A_dot = apply_eager
            except Exception as exc:
                print(f"CG graph flavour failed: {exc!r}", file=sys.stderr, flush=True)
        cg_out["allreduce"] = args.allreduce if world > 1 else "none"
        if peer_ar is not None:
            peer_ar.close()
        del st, cg, bvec

    # ---------------- CPU baseline (rank 0, N=1) --------------------------------
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            if args.matrix:
                cb = cpu_baseline_mtx(w)
            else:
                res = cpu_baseline(args.workload, size=args.size)
                cb = res[0] if res else None
        except Exception as exc:          # a missing checker must not void the GPU measurement
            print(f"cpu_baseline leg failed: {exc!r}", file=sys.stderr, flush=True)
            cb = None

    if rendezvous is not None:
        rendezvous.close()
    if peer is not None:
        peer.close()
    dev_status = L.spgpuGetDeviceStatus(h, 0)
    if rank == 0:
        cfg = workload_config(args.workload, args.size, world, args.halo, args.overlap, rows=w["total_rows"], nnz=int(nnz_total),
                              flush_l2=flush_l2, alpha=w["alpha"], beta=w["beta"]) if not args.matrix else \
            {"workload": w["label"], "rows": w["total_rows"], "nnz": int(nnz_total), "parallelism": "one GPU",
             "l2": "flushed" if flush_l2 else "inputs larger than L2", "alpha": str(w["alpha"]), "beta": str(w["beta"])}
        out = {
            "metric": "spmv_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"S": "f32", "D": "f64", "C": "c64", "Z": "c128"}[w["sym"]], "data": "synthetic",
            "config": cfg,
            "hbm_gbs": bytes_total / (ms_step * 1e-3) / 1e9,
            "hbm_frac_of_peak": bytes_total / (ms_step * 1e-3) / 1e9 / (peak * world),
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        out["device_status"] = dev_status          # 0, or SPGPU_DEVSTATUS_TIMEOUT if a wait for a peer GPU gave up
        if verified is not None:
            out["verified_vs_global_columns"] = verified
        if cg_out is not None:
            out["cg"] = cg_out
        if ref_kernels is not None:
            out["reference_kernels"] = ref_kernels
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    # Teardown in dependency order, then a NORMAL return: the driver's exit hooks (which record the shared
    # objects this process loaded) must run.  The stream is torch's (see above), so nothing torch still holds
    # refers to a stream the handle owned.
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream(device))
    del step, op, z, y, x_ext, w, scratch
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    L.spgpuSetStream(h, None)
    L.spgpuDestroy(h)
    sys.stderr.flush()
    os.dup2(json_fd, 1)                  # give stdout back (exit hooks may print)
    os.close(json_fd)
    return 0


if __name__ == "__main__":
    sys.exit(main())
