#!/usr/bin/env python
"""Where do the 4096-slot rows of cfg3 cost their time?  The same Pareto matrix (float HELL, 2^22 rows, mean 16) with its 128
forced spike rows (a) spread evenly (cfg3), (b) one per hack in the FIRST 128 hacks, (c) one per hack in the LAST 128 hacks,
(d) absent.  A warp walks a spike row alone; if (c) is slower than (b) by about one such walk, the kernel ends with
a tail of spike rows that started late."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from spgpu_b200 import capi, device_build as DB
    from spgpu_b200.generators import powerlaw_lengths
    L = capi.SpgpuLib(os.environ["SPGPU_LIB"]) if os.environ.get("SPGPU_LIB") else capi.lib()
    h = ctypes.c_void_p()
    assert L.spgpuCreate(ctypes.byref(h), 0) == 0
    stream = torch.cuda.Stream()
    L.spgpuSetStream(h, stream.cuda_stream)
    torch.cuda.set_stream(stream)
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.int64, device="cuda")
    R = 1 << 22
    base = powerlaw_lengths(R, 16, 4096, 0, 7)
    placements = {"spread (cfg3)": np.arange(16384, R, 32768), "first 128 hacks": np.arange(128) * 32 + 5,
                  "last 128 hacks": R - 1 - (np.arange(128) * 32 + 5), "none": np.zeros(0, dtype=np.int64)}
    T = capi.TYPES["S"]
    for name, where in placements.items():
        lens_np = base.copy()
        lens_np[where] = 4096
        lens = torch.from_numpy(lens_np).cuda()
        gen = torch.Generator(device="cuda"); gen.manual_seed(8)
        lo = torch.zeros(R, dtype=torch.int64, device="cuda"); hi = torch.full((R,), R - 1, dtype=torch.int64, device="cuda")
        _r, cols = DB._strided_columns(lens, lo, hi, gen)
        vals = torch.rand(cols.numel(), device="cuda", generator=gen, dtype=torch.float32) * 2 - 1
        A = DB.hell_from_rows(lens, cols, vals, R)
        x = torch.rand(R, device="cuda", dtype=torch.float32) * 2 - 1
        z = torch.zeros(R, dtype=torch.float32, device="cuda")
        ts = []
        for it in range(10):
            scratch.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            L.spgpuShellspmv(h, z.data_ptr(), 0, T.scalar(1.0), A.values.data_ptr(), A.indices.data_ptr(), 32, A.hack_offsets.data_ptr(),
                             A.rs.data_ptr(), 0, 16, R, x.data_ptr(), T.scalar(0.0), 0)
            b.record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(a.elapsed_time(b))
        print(json.dumps({"spikes": name, "nnz": int(A.nnz), "stored": int(A.values.numel()), "ms": round(float(np.mean(ts)), 5),
                          "min_ms": round(float(np.min(ts)), 5)}), flush=True)
        del A, lens, cols, vals, x, z
    torch.cuda.set_stream(torch.cuda.default_stream())
    L.spgpuSetStream(h, None)
    L.spgpuDestroy(h)


if __name__ == "__main__":
    main()
