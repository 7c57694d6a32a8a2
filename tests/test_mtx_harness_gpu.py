"""MatrixMarket files through the benchmark harness (bench.build_mtx_workload): C reader -> device-side
layout builders -> SpMV through the C ABI, every format, against the oracle on the matrix that was written."""
import numpy as np
import pytest

import bench
from spgpu_b200 import formats as F, generators as G, mmio
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["ell", "hell", "ohell", "dia", "hdia"])
@pytest.mark.parametrize("sym", ["D", "S"])
def test_mtx_file_through_the_harness(ours, gpu_handle, tmp_path, fmt, sym):
    import torch
    dt = np.float64 if sym == "D" else np.float32
    cases = [(G.laplace3d_7pt(14), "symmetric"), (G.stencil3d_27pt(9), "symmetric")]
    if fmt not in ("dia",):
        cases.append((G.random_coo(777, 650, (0, 12), 5, np.float64, 0), "general"))
    for k, (coo, symmetry) in enumerate(cases):
        path = str(tmp_path / f"m{k}.mtx")
        mmio.write_coo(path, coo, "real", symmetry)
        w = bench.build_mtx_workload(path, fmt, ours, gpu_handle, torch.device("cuda", 0), sym)
        assert w["nnz"] == coo.nnz and w["rows"] == coo.nrows
        x = G.random_vector(coo.ncols, dt, 1, -1, 1)
        dx = util.to_dev(x)
        dz = torch.full((coo.nrows,), float("nan"), dtype=torch.float64 if sym == "D" else torch.float32, device="cuda")
        step = bench.make_step(ours, gpu_handle, w, dx.data_ptr(), dz.data_ptr(), 0)
        step()
        torch.cuda.synchronize()
        c = F.Coo(coo.rows, coo.cols, coo.vals.astype(dt), coo.nrows, coo.ncols, coo.base)
        want = util.oracle_spmv("ell", F.coo_to_ell(c), x, None, 1.0, 0.0)
        util.assert_rows_close(dz.cpu().numpy(), want, util.row_scale(c, x, None, 1.0, 0.0), sym, f"mtx/{fmt}/{sym}")
        cb = bench.cpu_baseline_mtx(w, budget_s=0.2, repeats=1)
        assert cb["value"] > 0 and cb["kind"] == "port"
