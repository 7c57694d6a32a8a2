#!/usr/bin/env python
"""cfg2 (128^3 27-point stencil, double HDIA) on one GPU, back to back: the plain SpMV, the fused SpMV + dot, and the
pieces of a CG iteration -- where do the 0.30 ms per iteration of examples/mg_hdia.c go?"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from spgpu_b200 import capi, device_build as DB
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.spgpuCreate(ctypes.byref(h), 0) == 0
    stream = torch.cuda.Stream()
    L.spgpuSetStream(h, stream.cuda_stream)
    torch.cuda.set_stream(stream)
    A = DB.hdia_stencil27(n)
    rows = A.nrows
    T = capi.TYPES["D"]
    x = torch.rand(rows, dtype=torch.float64, device="cuda")
    z = torch.zeros(rows, dtype=torch.float64, device="cuda")
    r = torch.rand(rows, dtype=torch.float64, device="cuda")
    xx = torch.zeros(rows, dtype=torch.float64, device="cuda")
    s = torch.ones(4, dtype=torch.float64, device="cuda")
    one, zero = T.scalar(1.0), T.scalar(0.0)
    dm, off, ho = A.values.data_ptr(), A.offsets.data_ptr(), A.hack_offsets.data_ptr()
    sp = s.data_ptr()
    L.spgpuReserveScratch(h, 1 << 22)
    ops = {
        "spgpuDhdiaspmv": lambda: L.spgpuDhdiaspmv(h, z.data_ptr(), 0, one, dm, off, 32, ho, rows, rows, x.data_ptr(), zero),
        "spgpuDhdiaspmvHaloDot (no neighbours)": lambda: L.spgpuDhdiaspmvHaloDot(h, z.data_ptr(), dm, off, 32, ho, rows, rows, x.data_ptr(), 0, None, 1, sp + 8, None),
        "spgpuDcgUpdateDev": lambda: L.spgpuDcgUpdateDev(h, xx.data_ptr(), r.data_ptr(), x.data_ptr(), z.data_ptr(), rows, sp, sp + 8, sp + 16, None),
        "spgpuDaxpbyDev": lambda: L.spgpuDaxpbyDev(h, x.data_ptr(), rows, sp + 16, sp, 1.0, x.data_ptr(), 0, 0, 1.0, r.data_ptr()),
    }
    out = {}
    for name, fn in ops.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(50):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        out[name] = round(a.elapsed_time(b) / 50, 5)
        s.fill_(1.0)
    print(json.dumps(out, indent=1))
    torch.cuda.set_stream(torch.cuda.default_stream())
    L.spgpuSetStream(h, None)
    L.spgpuDestroy(h)


if __name__ == "__main__":
    main()
