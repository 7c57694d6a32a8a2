#!/usr/bin/env python
"""What a pure streaming kernel reaches on SHORT problems with a cold L2 (context for cfg1 / cfg2).

    python bench/stream_floor.py > gpurun_out/r1_stream_floor.json

MEASURED_PEAKS.json's HBM figure comes from a long copy.  cfg1 moves 80 MB (12 us at that rate)
and cfg2 489 MB (75 us): launch ramp, the last partial wave and the cold L2 are a visible part of
such a run whatever the kernel does.  This times `spgpuDscal` (z = a*x: n read + n written, no
gather, no metadata) through the same C ABI, with the same protocol bench.py uses for cfg1/cfg2
(L2 evicted by reading a 512 MB scratch before every launch, one CUDA-event pair per launch), at
the byte counts of the SpMV configurations.  The result is the practical ceiling a same-sized SpMV
is compared with in profiles/README.md.
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
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
    T = capi.TYPES["D"]
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.int64, device="cuda")
    out = {"peak_gbs": peak, "op": "spgpuDscal (n doubles read + n written), cold L2, one event pair per launch"}
    for label, total in (("cfg1 (80 MB)", 79.95e6), ("cfg2 (489 MB)", 489e6), ("cfg4 (1.70 GB)", 1.70e9)):
        n = int(total // 16)
        x = torch.rand(n, dtype=torch.float64, device="cuda")
        z = torch.empty(n, dtype=torch.float64, device="cuda")
        ts = []
        for it in range(13):
            scratch.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            L.spgpuDscal(h, z.data_ptr(), n, T.scalar(1.5), x.data_ptr())
            b.record(stream)
            b.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts))
        out[label] = {"bytes": 16 * n, "ms": ms, "min_ms": float(np.min(ts)), "gbs": 16 * n / ms / 1e6,
                      "frac_of_measured_peak": 16 * n / ms / 1e6 / peak}
        del x, z
    print(json.dumps(out, indent=1))
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    os._exit(0)


if __name__ == "__main__":
    main()
