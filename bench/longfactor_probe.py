#!/usr/bin/env python
"""hellLongFactor (a row deeper than factor x avgNnzPerRow slots, at least 32, is finished by the whole warp) on a family
of irregular HELL matrices: Pareto row lengths with several means / caps / spike densities, uniform lengths, float and
double.  Event-timed with an L2 flush between launches; JSON lines on stdout."""
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
    L = capi.SpgpuLib(os.environ["SPGPU_LIB"]) if os.environ.get("SPGPU_LIB") else capi.lib()      # A/B builds
    h = ctypes.c_void_p()
    assert L.spgpuCreate(ctypes.byref(h), 0) == 0
    stream = torch.cuda.Stream()
    L.spgpuSetStream(h, stream.cuda_stream)
    torch.cuda.set_stream(stream)
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.int64, device="cuda")
    R = 1 << 21
    factors = [int(f) for f in (sys.argv[1].split(",") if len(sys.argv) > 1 else "2,3,4,6".split(","))]
    cases = []
    for dtype in (torch.float32, torch.float64):
        for mean, maxlen, spike in ((16, 4096, 32768), (8, 4096, 32768), (32, 4096, 32768), (16, 512, 4096), (64, 1024, 0), (16, 64, 0)):
            cases.append(("pareto", dtype, mean, maxlen, spike))
        cases.append(("uniform", dtype, 16, 32, 0))
    for kind, dtype, mean, maxlen, spike in cases:
        if kind == "pareto":
            lens, cols, vals = DB.powerlaw_entries(R, mean=mean, maxlen=maxlen, spike_every=spike, dtype=dtype)
        else:
            gen = torch.Generator(device="cuda"); gen.manual_seed(5)
            lens = torch.randint(0, maxlen + 1, (R,), device="cuda", generator=gen)
            lo = torch.zeros(R, dtype=torch.int64, device="cuda"); hi = torch.full((R,), R - 1, dtype=torch.int64, device="cuda")
            _r, cols = DB._strided_columns(lens, lo, hi, gen)
            vals = (torch.rand(cols.numel(), device="cuda", generator=gen, dtype=torch.float32) * 2 - 1).to(dtype)
        A = DB.hell_from_rows(lens, cols, vals, R)
        avg = max(1, int(round(A.nnz / R)))
        s = "S" if dtype == torch.float32 else "D"
        T = capi.TYPES[s]
        x = (torch.rand(R, device="cuda", dtype=torch.float32) * 2 - 1).to(dtype)
        z = torch.zeros(R, dtype=dtype, device="cuda")
        fn = getattr(L, f"spgpu{s}hellspmv")
        out = {"case": f"{kind} mean {mean} max {maxlen} spike_every {spike}", "type": s, "avg": avg, "nnz": int(A.nnz),
               "padding": float(A.values.numel()) / max(1, int(A.nnz))}
        ref = None
        for f in factors:
            assert L.spgpuSetTuning(h, b"hellLongFactor", f) == 0
            ts = []
            for it in range(8):
                scratch.sum()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                fn(h, z.data_ptr(), 0, T.scalar(1.0), A.values.data_ptr(), A.indices.data_ptr(), 32, A.hack_offsets.data_ptr(),
                   A.rs.data_ptr(), 0, avg, R, x.data_ptr(), T.scalar(0.0), 0)
                b.record(stream)
                torch.cuda.synchronize()
                if it >= 2:
                    ts.append(a.elapsed_time(b))
            out[f"factor{f}_ms"] = round(float(np.mean(ts)), 5)
            zc = z.double().cpu()
            if ref is None:
                ref = zc
            else:
                out[f"factor{f}_maxdiff"] = float((zc - ref).abs().max())
        print(json.dumps(out), flush=True)
        del A, lens, cols, vals, x, z
    torch.cuda.set_stream(torch.cuda.default_stream())
    L.spgpuSetStream(h, None)
    L.spgpuDestroy(h)


if __name__ == "__main__":
    main()
