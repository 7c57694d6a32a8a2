"""examples/spmv_mtx.c: a C-only consumer (no Python between the file and the GPU) -- MatrixMarket reader,
the reference's conversion calls, SpMV in three formats, checked in C against a host loop."""
import os
import subprocess

import numpy as np
import pytest

from spgpu_b200 import generators as G, mmio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / "spmv_mtx")
    cmd = ["gcc", "-O2", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", "spmv_mtx.c"), f"-I{ROOT}/include",
           "-I/usr/local/cuda/include", f"-L{ROOT}/spgpu_b200/lib", "-lspgpu", f"-Wl,-rpath,{ROOT}/spgpu_b200/lib",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_c_driver_compiles_against_the_headers(tmp_path):
    build(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["lap7_sym", "st27_sym", "random_general"])
def test_c_driver_runs_and_agrees_with_its_host_loop(tmp_path, case):
    exe = build(tmp_path)
    coo, sym = {"lap7_sym": (G.laplace3d_7pt(20), "symmetric"), "st27_sym": (G.stencil3d_27pt(12), "symmetric"),
                "random_general": (G.random_coo(1500, 1300, (0, 14), 9, np.float64, 0), "general")}[case]
    path = str(tmp_path / "m.mtx")
    mmio.write_coo(path, coo, "real", sym)
    p = subprocess.run([exe, path, "5"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"{coo.nnz} non-zeros" in p.stdout
    assert p.stdout.count(" ok") == 3 and "FAILED" not in p.stdout


def build_mg(tmp_path, name="mg_cg"):
    exe = str(tmp_path / name)
    cmd = ["gcc", "-O2", "-fopenmp", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", name + ".c"), f"-I{ROOT}/include",
           "-I/usr/local/cuda/include", f"-L{ROOT}/spgpu_b200/lib", "-lspgpu", f"-Wl,-rpath,{ROOT}/spgpu_b200/lib",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_multi_gpu_c_driver_compiles_against_the_headers(tmp_path):
    build_mg(tmp_path)
    build_mg(tmp_path, "mg_hdia")


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [1, 2, 4])
def test_multi_gpu_c_driver_runs(tmp_path, ranks):
    """examples/mg_cg.c: the partitioned 7-point Laplacian + CG through include/spgpu_mg.h with no Python; more ranks
    than devices puts several ranks on one device (EVENTS exchange), one rank per device uses the fused exchange"""
    exe = build_mg(tmp_path)
    p = subprocess.run([exe, "32", str(ranks), "5", "40"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "OK" in p.stdout and "FAILED" not in p.stdout and "(0 rows over 1e-12)" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [1, 2, 3])
def test_multi_gpu_hdia_c_driver_runs(tmp_path, ranks):
    """examples/mg_hdia.c: the 27-point stencil converted with the library's own cooToHdia, partitioned by
    spgpuMgDhdiaCreate, multiplied through spgpuMgDhdiaspmv / spgpuMgDspmv, solved with CG -- no Python"""
    exe = build_mg(tmp_path, "mg_hdia")
    p = subprocess.run([exe, "24", str(ranks), "5", "25"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "OK" in p.stdout and "FAILED" not in p.stdout and "(0 rows over 1e-12)" in p.stdout and "fits" in p.stdout
