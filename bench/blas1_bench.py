#!/usr/bin/env python
"""HBM GB/s of the BLAS-1 companions through the C ABI on one B200.

    python bench/blas1_bench.py [--n 134217728] > gpurun_out/blas1.json

Vectors are 1 GiB each (double, n = 2^27: the length of a cfg5 vector), far larger than L2.
Each op is timed with CUDA events on the handle's stream over `reps` back-to-back launches
(the blocking reductions include their host synchronisation, as a caller pays it).
Prints one JSON object: op -> {ms, algorithmic GB moved, GB/s, fraction of the measured HBM peak}.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 27)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sweep", action="store_true", help="also sweep the reductions' redInflight x redBlocksPerSm knobs")
    args = ap.parse_args()
    import torch
    from spgpu_b200 import capi
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.spgpuCreate(ctypes.byref(h), 0) == 0
    stream = torch.cuda.ExternalStream(L.spgpuGetStream(h))
    torch.cuda.set_stream(stream)
    n = args.n
    T = capi.TYPES["D"]
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.rand(n, dtype=torch.float64, device="cuda")
    z = torch.empty(n, dtype=torch.float64, device="cuda")
    w = torch.rand(n, dtype=torch.float64, device="cuda")
    m = n // 16
    idx = torch.randperm(n, device="cuda")[:m].to(torch.int32)
    xv = torch.rand(m, dtype=torch.float64, device="cuda")
    dres = torch.zeros(4, dtype=torch.float64, device="cuda")
    X, Y, Z = x.data_ptr(), y.data_ptr(), z.data_ptr()
    ops = {
        "spgpuDaxpby  z=by+ax": (lambda: L.spgpuDaxpby(h, Z, n, T.scalar(0.5), Y, T.scalar(1.5), X), 24 * n),
        "spgpuDaxpby  in place (z=y)": (lambda: L.spgpuDaxpby(h, Y, n, T.scalar(0.999), Y, T.scalar(1e-3), X), 24 * n),
        "spgpuDscal": (lambda: L.spgpuDscal(h, Z, n, T.scalar(1.5), X), 16 * n),
        "spgpuDdot (blocking)": (lambda: L.spgpuDdot(h, n, X, Y), 16 * n),
        "spgpuDnrm2 (blocking)": (lambda: L.spgpuDnrm2(h, n, X), 8 * n),
        "spgpuDamax (blocking)": (lambda: L.spgpuDamax(h, n, X), 8 * n),
        "spgpuDasum (blocking)": (lambda: L.spgpuDasum(h, n, X), 8 * n),
        "spgpuDdotDev (device result)": (lambda: L.spgpuDdotDev(h, n, X, Y, dres.data_ptr()), 16 * n),
        "spgpuDcgUpdateDev": (lambda: L.spgpuDcgUpdateDev(h, Z, Y, X, w.data_ptr(), n, dres.data_ptr() + 8, dres.data_ptr() + 16,
                                                          dres.data_ptr(), None), 48 * n),
        "spgpuDgath (n/16 random indices)": (lambda: L.spgpuDgath(h, xv.data_ptr(), m, idx.data_ptr(), 0, X), 20 * m),
        "spgpuDscat (n/16 random indices, beta=2)": (lambda: L.spgpuDscat(h, Z, m, xv.data_ptr(), idx.data_ptr(), 0,
                                                                           T.scalar(2.0)), 28 * m),
    }
    dres[1] = 1.0
    dres[2] = 1e9
    out = {"n": n, "peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs"}
    for name, (fn, nbytes) in ops.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.reps
        out[name] = {"ms": ms, "algorithmic_gb": nbytes / 1e9, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    if args.sweep:
        sweep = {}
        red = {k: v for k, v in ops.items() if any(t in k for t in ("DdotDev", "Dnrm2", "Damax", "Ddot (blocking)"))}
        for infl in (2, 4, 8):
            for bps in (2, 4, 6, 8):
                L.spgpuSetTuning(h, b"redInflight", infl)
                L.spgpuSetTuning(h, b"redBlocksPerSm", bps)
                row = {}
                for name, (fn, nbytes) in red.items():
                    for _ in range(2):
                        fn()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    for _ in range(args.reps):
                        fn()
                    b.record(stream)
                    torch.cuda.synchronize()
                    row[name] = round(nbytes / (a.elapsed_time(b) / args.reps) / 1e6 / peak, 4)
                sweep[f"inflight={infl},blocksPerSm={bps}"] = row
        L.spgpuSetTuning(h, b"redInflight", 0)
        L.spgpuSetTuning(h, b"redBlocksPerSm", 8)
        out["reduction_sweep_frac_of_peak"] = sweep
    print(json.dumps(out, indent=1))
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    os._exit(0)


if __name__ == "__main__":
    main()
