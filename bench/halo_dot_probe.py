#!/usr/bin/env python
"""One GPU, emulated neighbours: the variants of the HELL row-block kernel on the SAME slab (a 256-plane half of the
512^3 Laplacian with local columns, what rank 0 of 2 multiplies): plain, fused-halo, +dot without / with the halo code.
Event-timed; run under ncu for per-kernel metrics."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from spgpu_b200 import capi, device_build as DB
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    parts = int(os.environ.get("PROBE_PARTS", "2"))           # the slab is rank 1 of `parts` ranks (both neighbours when parts > 2)
    L = capi.SpgpuLib(os.environ["SPGPU_LIB"]) if os.environ.get("SPGPU_LIB") else capi.lib()      # A/B builds
    h = ctypes.c_void_p()
    assert L.spgpuCreate(ctypes.byref(h), 0) == 0
    stream = torch.cuda.Stream()
    L.spgpuSetStream(h, stream.cuda_stream)
    torch.cuda.set_stream(stream)
    plane = n * n
    per = n // parts
    A = DB.hell_laplace3d_7pt(n, per if parts > 2 else 0, 2 * per if parts > 2 else per, local_columns=True)
    rows = A.nrows
    x = torch.rand(A.ncols, dtype=torch.float64, device="cuda")
    z = torch.zeros(rows, dtype=torch.float64, device="cuda")
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    pflags = torch.zeros(16, dtype=torch.int32, device="cuda")
    pz = torch.zeros(plane, dtype=torch.float64, device="cuda")
    lk = capi.HaloLinks()
    lk.peerHiLowerZone, lk.peerFlagsHi = pz.data_ptr(), pflags.data_ptr()
    if parts > 2:
        pzl = torch.zeros(plane, dtype=torch.float64, device="cuda")
        pfl = torch.zeros(16, dtype=torch.int32, device="cuda")
        lk.peerLoUpperZone, lk.peerFlagsLo = pzl.data_ptr(), pfl.data_ptr()
    lk.myFlags = flags.data_ptr()
    flags[4:8] = 1 << 30                      # every exchange "has arrived" and every push "has been consumed"
    T = capi.TYPES["D"]
    one, zero = T.scalar(1.0), T.scalar(0.0)
    cm, rp, ho, rs = A.values.data_ptr(), A.indices.data_ptr(), A.hack_offsets.data_ptr(), A.rs.data_ptr()
    seq = [0]

    def nxt():
        seq[0] += 1
        return seq[0]
    variants = {
        "spgpuDhellspmv (plain)": lambda: L.spgpuDhellspmv(h, z.data_ptr(), 0, one, cm, rp, 32, ho, rs, 0, 7, rows, x.data_ptr(), zero, 0),
        "spgpuDhellspmvHalo (emulated upper neighbour)": lambda: L.spgpuDhellspmvHalo(h, z.data_ptr(), 0, one, cm, rp, 32, ho, rs, 7, rows, x.data_ptr(), zero, 0, plane, ctypes.byref(lk), nxt()),
        "spgpuDhellspmvHaloDot (no neighbours)": lambda: L.spgpuDhellspmvHaloDot(h, z.data_ptr(), cm, rp, 32, ho, rs, 7, rows, x.data_ptr(), 0, plane, None, 1, res.data_ptr(), None),
        "spgpuDhellspmvHaloDot (emulated upper neighbour)": lambda: L.spgpuDhellspmvHaloDot(h, z.data_ptr(), cm, rp, 32, ho, rs, 7, rows, x.data_ptr(), 0, plane, ctypes.byref(lk), nxt(), res.data_ptr(), None),
    }
    out = {}
    if len(sys.argv) > 3:
        for kv in sys.argv[3].split(","):
            k, v = kv.split("=")
            assert L.spgpuSetTuning(h, k.encode(), int(v)) == 0
        out["tuning"] = sys.argv[3]
    for name, fn in variants.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / reps
    print(json.dumps(out, indent=1))
    assert L.spgpuGetDeviceStatus(h, 0) == 0
    torch.cuda.set_stream(torch.cuda.default_stream())
    L.spgpuSetStream(h, None)
    L.spgpuDestroy(h)


if __name__ == "__main__":
    main()
