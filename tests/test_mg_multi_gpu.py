"""The REAL multi-rank paths, on a box with at least two GPUs (skipped on a one-GPU box: there the same code runs
with emulated neighbours, tests/test_mg_gpu.py, and with several ranks on one device, tests/test_mg_capi_gpu.py).

  * bench.py under torchrun (one process per GPU, CUDA IPC): the fused SpMV + halo kernel of every rank against
    the same rows multiplied with global columns, bit for bit; the CG flavours against each other;
  * the C-level API (one process, peer access): FUSED exchange on distinct devices for all four value types;
  * examples/mg_cg.c with one rank per device.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from spgpu_b200 import capi, formats as F, generators as G
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _torchrun(n, args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "bench.py"), "--gpus", str(n)] + args
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,size", [("cfg5", 128), ("cfg2", 64), ("cfg4", 200000)])
def test_partitioned_bench_is_verified_on_two_gpus(workload, size):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    d = _torchrun(2, ["--workload", workload, "--size", str(size), "--steps", "6", "--warmup", "3", "--no-cpu", "--no-e2e"])
    assert d["n_gpus"] == 2 and d["verified_vs_global_columns"] is True and d["device_status"] == 0
    assert d["gpu_launches"] == 6                                  # ONE kernel per partitioned SpMV
    if workload == "cfg5":
        cg = d["cg"]
        assert cg["device"]["kernels_per_iteration"] <= 5          # no separate all-reduce launches
        rb, rd = cg["blocking"]["residual_norm2_after"], cg["device"]["residual_norm2_after"]
        assert abs(rb - rd) <= 1e-9 * abs(rb), (rb, rd)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_c_api_fused_exchange_on_two_devices(ours, gpu_handle, dtype):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from tests.test_mg_capi_gpu import Mg, scalars
    s = util.sym_of(dtype)
    t = util.TYPES[s]
    coo = G.laplace3d_7pt(24)
    vals = coo.vals.astype(dtype)
    if t.is_complex:
        vals = (vals + 0.25j * np.random.default_rng(1).standard_normal(vals.shape[0])).astype(dtype)
    coo = F.Coo(coo.rows, coo.cols, vals, coo.nrows, coo.ncols, coo.base)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    x = G.random_vector(coo.nrows, dtype, 1, -1, 1)
    y = G.random_vector(coo.nrows, dtype, 2, -1, 1)
    alpha, beta = scalars(dtype)
    single = util.dev_spmv(ours, gpu_handle, "hell", hell, util.upload(hell), x, y, alpha, beta, avg=7)
    mg = Mg(ours, list(range(min(_gpus(), 4))))
    try:
        assert ours.spgpuMgExchange(mg.h) == capi.MG_FUSED
        A = mg.matrix(hell)
        vx, vy, vz = mg.vector(A, x), mg.vector(A, y), mg.vector(A)
        for _ in range(5):
            assert getattr(ours, f"spgpuMg{s}hellspmv")(mg.h, vz, vy, t.scalar(alpha), A, vx, t.scalar(beta)) == 0
        assert ours.spgpuMgSynchronize(mg.h) == 0
        assert np.array_equal(mg.get(vz, coo.nrows, dtype).view(np.uint8), single.view(np.uint8))
        # the same through the EVENTS exchange
        assert ours.spgpuMgSetExchange(mg.h, capi.MG_EVENTS) == 0
        assert getattr(ours, f"spgpuMg{s}hellspmv")(mg.h, vz, vy, t.scalar(alpha), A, vx, t.scalar(beta)) == 0
        assert ours.spgpuMgSynchronize(mg.h) == 0
        assert np.array_equal(mg.get(vz, coo.nrows, dtype).view(np.uint8), single.view(np.uint8))
        for v in (vx, vy, vz):
            ours.spgpuMgVectorDestroy(v)
        ours.spgpuMgMatrixDestroy(A)
    finally:
        mg.close()


def test_c_example_one_rank_per_device(tmp_path):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    from tests.test_c_driver_gpu import build_mg
    exe = build_mg(tmp_path)
    p = subprocess.run([exe, "64", str(min(_gpus(), 8)), "10", "60"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "fused into the SpMV kernel" in p.stdout and "OK" in p.stdout and "(0 rows over 1e-12)" in p.stdout
