"""Same box, same buffers: OUR kernels against the REFERENCE'S OWN kernels (its unmodified
sources compiled for sm_100a, oracle/_ref/libspgpu_ref.so).  BASELINE.md section 2 asks for
this "reference kernels on B200" column; it lives in tests/ because only tests may execute
anything under oracle/.  Writes gpurun_out/ref_vs_ours.json and requires that we are at
least as fast as the reference on its own headline formats."""
import json
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, stream, reps=10):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def test_ours_vs_reference_kernels(ours, ref, gpu_handle, ref_handle):
    import torch
    from spgpu_b200 import device_build as DB
    T = util.TYPES["D"]
    out = {}
    legacy = torch.cuda.default_stream()

    # ---- double HELL, 7-point Laplacian 256^3 (1/8 of BASELINE configs[4]) ----
    A = DB.hell_laplace3d_7pt(256)
    x = torch.rand(A.ncols, dtype=torch.float64, device="cuda")
    z1 = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
    z2 = torch.zeros_like(z1)

    def hell(L, h, z):
        return lambda: L.spgpuDhellspmv(h, z.data_ptr(), 0, T.scalar(1.0), A.values.data_ptr(), A.indices.data_ptr(), 32,
                                        A.hack_offsets.data_ptr(), A.rs.data_ptr(), 0, 7, A.nrows, x.data_ptr(),
                                        T.scalar(0.0), 0)
    t_ours = _time(hell(ours, gpu_handle, z1), legacy)
    t_ref = _time(hell(ref, ref_handle, z2), legacy)
    torch.cuda.synchronize()
    assert torch.allclose(z1, z2, rtol=0, atol=1e-11)
    bytes_hell = A.nnz * 12 + 4 * A.nrows + 4 * A.hack_offsets.numel() + 16 * A.nrows
    out["hell_d_lap7_256"] = {"ours_ms": t_ours, "reference_ms": t_ref, "speedup": t_ref / t_ours,
                              "ours_gbs": bytes_hell / t_ours / 1e6, "reference_gbs": bytes_hell / t_ref / 1e6}
    del A, x, z1, z2

    # ---- double HDIA, 27-point stencil 128^3 (BASELINE configs[1]) ----
    H = DB.hdia_stencil27(128)
    x = torch.rand(H.ncols, dtype=torch.float64, device="cuda")
    z1 = torch.zeros(H.nrows, dtype=torch.float64, device="cuda")
    z2 = torch.zeros_like(z1)

    def hdia(L, h, z):
        return lambda: L.spgpuDhdiaspmv(h, z.data_ptr(), 0, T.scalar(1.0), H.values.data_ptr(), H.offsets.data_ptr(), 32,
                                        H.hack_offsets.data_ptr(), H.nrows, H.ncols, x.data_ptr(), T.scalar(0.0))
    t_ours = _time(hdia(ours, gpu_handle, z1), legacy)
    t_ref = _time(hdia(ref, ref_handle, z2), legacy)
    torch.cuda.synchronize()
    assert torch.allclose(z1, z2, rtol=0, atol=1e-10)
    bytes_hdia = H.cells_in_range * 8 + 4 * H.offsets.numel() + 4 * H.hack_offsets.numel() + 16 * H.nrows
    out["hdia_d_st27_128"] = {"ours_ms": t_ours, "reference_ms": t_ref, "speedup": t_ref / t_ours,
                              "ours_gbs": bytes_hdia / t_ours / 1e6, "reference_gbs": bytes_hdia / t_ref / 1e6,
                              "note": "reference time is wall time per call incl. its per-call texture bind (a blocking "
                                      "cudaMemcpyToSymbol under oracle/texshim.h); L2 not flushed for either"}

    # ---- dot on 64 M doubles ----
    n = 1 << 26
    a = torch.rand(n, dtype=torch.float64, device="cuda")
    t_ours = _time(lambda: ours.spgpuDdot(gpu_handle, n, a.data_ptr(), a.data_ptr()), legacy, 5)
    t_ref = _time(lambda: ref.spgpuDdot(ref_handle, n, a.data_ptr(), a.data_ptr()), legacy, 5)
    out["ddot_64M"] = {"ours_ms": t_ours, "reference_ms": t_ref, "speedup": t_ref / t_ours,
                       "ours_gbs": 8 * n / t_ours / 1e6, "reference_gbs": 8 * n / t_ref / 1e6}

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_vs_ours.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
    for k, v in out.items():
        assert v["speedup"] >= 1.0, (k, v)
