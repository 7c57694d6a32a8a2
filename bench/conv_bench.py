#!/usr/bin/env python
"""Device-side format construction vs the host conversions.

    python bench/conv_bench.py > gpurun_out/r1_conv.json

  * CSR -> HELL on the device (spgpuCsrToHellLayoutDevice + spgpuDcsrToHellDevice) for the 512^3
    7-point Laplacian (cfg5: 134 M rows, 938 M nnz), timed with CUDA events around both calls;
  * COO -> HDIA on the device (spgpuHdiaHackOffsetsFromCooDevice + spgpuDcooToHdiaDevice) for the
    128^3 27-point stencil (cfg2: 55.7 M nnz), likewise;
  * the host route of the reference ABI (cooToEll + ellToHell, cooToHdia: our bit-exact C port,
    one thread, as in the reference) on a 128^3 / 64^3 sample, scaled per non-zero.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from spgpu_b200 import capi, device_build as DB, formats as F, generators as G
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.spgpuCreate(ctypes.byref(h), 0) == 0
    stream = torch.cuda.ExternalStream(L.spgpuGetStream(h))
    torch.cuda.set_stream(stream)
    out = {}

    # ---- CSR -> HELL, cfg5 ------------------------------------------------------------------
    n = 512
    A = DB.hell_laplace3d_7pt(n)
    N = A.nrows
    rs64 = A.rs.to(torch.int64)
    rowptr64 = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    torch.cumsum(rs64, 0, out=rowptr64[1:])
    nnz = int(rowptr64[-1].item())
    cols = torch.empty(nnz, dtype=torch.int32, device="cuda")
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    rows = torch.arange(N, device="cuda", dtype=torch.int64)
    at = A.hack_offsets.to(torch.int64)[rows // 32] + rows % 32
    for k in range(7):
        live = rs64 > k
        src, dst = (at + 32 * k)[live], (rowptr64[:-1] + k)[live]
        cols[dst] = A.indices[src]
        vals[dst] = A.values[src]
        del src, dst, live
    total_ref = A.values.numel()
    del rows, at, rs64, A
    rowptr = rowptr64.to(torch.int32)
    del rowptr64
    rs = torch.empty(N, dtype=torch.int32, device="cuda")
    hoff = torch.empty(N // 32, dtype=torch.int32, device="cuda")
    hv = torch.empty(total_ref, dtype=torch.float64, device="cuda")
    hi = torch.empty(total_ref, dtype=torch.int32, device="cuda")
    total = ctypes.c_longlong(0)
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        assert L.spgpuCsrToHellLayoutDevice(h, N, rowptr.data_ptr(), 32, rs.data_ptr(), hoff.data_ptr(), ctypes.byref(total)) == 0
        L.spgpuDcsrToHellDevice(h, N, rowptr.data_ptr(), cols.data_ptr(), vals.data_ptr(), 0, 32, hoff.data_ptr(), 0, hv.data_ptr(), hi.data_ptr())
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.min(ts[1:]))
    moved = nnz * 12 * 2 + N * 4 * 3          # CSR read + HELL written + rowPtr / rS / (hackOffsets)
    out["csr_to_hell_device_cfg5"] = {"rows": N, "nnz": nnz, "ms": ms, "mnnz_per_s": nnz / ms / 1e3, "gbs_moved": moved / ms / 1e6}
    ridx = torch.empty(N, dtype=torch.int32, device="cuda")
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        assert L.spgpuCsrToOhellLayoutDevice(h, N, rowptr.data_ptr(), 32, ridx.data_ptr(), rs.data_ptr(), hoff.data_ptr(), ctypes.byref(total)) == 0
        L.spgpuDcsrToOhellDevice(h, N, rowptr.data_ptr(), cols.data_ptr(), vals.data_ptr(), 0, 32, hoff.data_ptr(), ridx.data_ptr(), 0,
                                 hv.data_ptr(), hi.data_ptr())
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    out["csr_to_ohell_device_cfg5"] = {"rows": N, "nnz": nnz, "ms": float(np.min(ts[1:])), "mnnz_per_s": nnz / float(np.min(ts[1:])) / 1e3}
    del cols, vals, rowptr, rs, hoff, hv, hi, ridx
    torch.cuda.empty_cache()

    # ---- COO -> HDIA, cfg2 ------------------------------------------------------------------
    n = 128
    N = n ** 3
    r = torch.arange(N, device="cuda", dtype=torch.int64)
    x, y, z = r % n, (r // n) % n, r // (n * n)
    R_, C_, V_ = [], [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (z + dz >= 0) & (z + dz < n) & (y + dy >= 0) & (y + dy < n) & (x + dx >= 0) & (x + dx < n)
                rr = r[ok]
                R_.append(rr.to(torch.int32)); C_.append((rr + (dz * n + dy) * n + dx).to(torch.int32))
                V_.append(torch.full((rr.numel(),), 26.0 if (dz, dy, dx) == (0, 0, 0) else -1.0, dtype=torch.float64, device="cuda"))
    rows, cols, vals = torch.cat(R_), torch.cat(C_), torch.cat(V_)
    nnz = rows.numel()
    hoff = torch.empty(N // 32 + 1, dtype=torch.int32, device="cuda")
    height = ctypes.c_int(0)
    ts = []
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        assert L.spgpuHdiaHackOffsetsFromCooDevice(h, ctypes.byref(height), hoff.data_ptr(), 32, N, N, nnz, rows.data_ptr(), cols.data_ptr(), 0) == 0
        if it == 0:
            hv = torch.zeros(height.value * 32, dtype=torch.float64, device="cuda")
            off = torch.empty(height.value, dtype=torch.int32, device="cuda")
        hv.zero_()
        assert L.spgpuDcooToHdiaDevice(h, hv.data_ptr(), off.data_ptr(), hoff.data_ptr(), 32, N, N, nnz, rows.data_ptr(), cols.data_ptr(),
                                       vals.data_ptr(), 0) == 0
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.min(ts[1:]))
    out["coo_to_hdia_device_cfg2"] = {"rows": N, "nnz": nnz, "hack_diagonals": height.value, "ms": ms, "mnnz_per_s": nnz / ms / 1e3}

    # ---- the host route on a sample (one thread, like the reference) --------------------------
    coo = G.laplace3d_7pt(128)
    t0 = time.perf_counter()
    F.ell_to_hell(F.coo_to_ell(coo), 32)
    dt = time.perf_counter() - t0
    out["coo_to_ell_to_hell_host_128^3"] = {"nnz": coo.nnz, "ms": dt * 1e3, "mnnz_per_s": coo.nnz / dt / 1e6,
                                            "cfg5_extrapolated_s": dt * 937951232 / coo.nnz}
    coo = G.stencil3d_27pt(64)
    t0 = time.perf_counter()
    F.coo_to_hdia(coo, 32)
    dt = time.perf_counter() - t0
    out["coo_to_hdia_host_64^3"] = {"nnz": coo.nnz, "ms": dt * 1e3, "mnnz_per_s": coo.nnz / dt / 1e6,
                                    "cfg2_extrapolated_s": dt * 55742968 / coo.nnz}
    print(json.dumps(out, indent=1))
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    os._exit(0)


if __name__ == "__main__":
    main()
