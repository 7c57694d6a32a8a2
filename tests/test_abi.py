"""The C-ABI boundary: libspgpu.so loads without a GPU and exports every symbol
the headers under include/ declare; the reference's own test programs compile
and link against it unmodified."""
import ctypes
import os
import re
import subprocess

import pytest

from spgpu_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


def declared_symbols(header):
    """function names declared by a header, from the preprocessed text"""
    out = subprocess.run(["gcc", "-E", "-P", f"-I{CUDA_INC}", f"-I{ROOT}/include", os.path.join(ROOT, "include", header)],
                         capture_output=True, text=True, check=True).stdout
    names = set(re.findall(r"\b((?:spgpu[A-Z][A-Za-z0-9]*|compute[A-Z]\w*|cooTo\w+|coo2dia|ellTo\w+|diaToHdia|getHdiaHacksCount))\s*\(", out))
    return {n for n in names if not n.endswith("_t")}


def test_library_exports_every_declared_symbol(ours):
    for header, listed in (("spgpu.h", capi.abi_symbols()), ("spgpu_ext.h", capi.abi_symbols() + capi.ext_symbols()),
                           ("spgpu_mg.h", capi.abi_symbols() + capi.ext_symbols() + capi.mg_symbols())):
        declared = declared_symbols(header)
        assert len(declared) > 100
        missing = sorted(n for n in declared if not ours.has(n))
        assert not missing, f"{header} declares symbols the library does not export: {missing}"
        assert set(listed) >= declared - {"spgpuHandleStruct"}, sorted(declared - set(listed))


def test_library_exports_the_matrixmarket_reader(ours):
    declared = declared_symbols("spgpu_mm.h") - declared_symbols("spgpu.h")
    assert declared == {"spgpuMmLoadProperties", "spgpuMmLoadMatrixToCoo", "spgpuMmUnfoldedSymmetricSize",
                        "spgpuMmUnfoldSymmetric", "spgpuMmLoadDenseVector"}
    assert not [n for n in declared if not ours.has(n)]


def test_every_c_symbol_of_the_reference_library_is_exported():
    """link-level drop-in: whatever an application could have linked from the reference's libspgpu
    (its own build, oracle/_ref/libspgpu_ref.so) resolves against ours; `merge` / `mergesort` are
    ell.c's sort helpers leaking out of the reference (not declared in any header)"""
    from tests import util
    if not os.path.exists(util.REF_PATH):
        pytest.skip("oracle/_ref/libspgpu_ref.so not built (needs /root/reference at build time)")

    def exported(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
        return {ln.split()[2] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] == "T"}
    theirs = {n for n in exported(util.REF_PATH) if not n.startswith("_")}
    missing = theirs - exported(capi.LIB_PATH) - {"merge", "mergesort"}
    assert len(theirs) > 100 and not missing, sorted(missing)


def test_handle_struct_layout_is_the_reference_abi():
    """reference core.h:60-82: two pointers then nine ints, in this order"""
    S = capi.SpgpuHandleStruct
    names = [f[0] for f in S._fields_]
    assert names == ["currentStream", "defaultStream", "device", "warpSize", "maxThreadsPerBlock", "maxGridSizeX",
                     "maxGridSizeY", "maxGridSizeZ", "multiProcessorCount", "capabilityMajor", "capabilityMinor"]
    assert S.currentStream.offset == 0 and S.defaultStream.offset == 8 and S.device.offset == 16
    assert S.capabilityMinor.offset == 48 and ctypes.sizeof(S) == 56


def test_create_without_a_gpu_reports_an_error_code(ours):
    """no CPU fallback: on a box without a device spgpuCreate must not report success
    (reference core.c:33-38 returns SPGPU_UNSPECIFIED when the property query fails)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert ours.spgpuCreate(ctypes.byref(h), 0) == capi.SPGPU_UNSPECIFIED
    if h.value:
        ours.spgpuDestroy(h)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        capi.SpgpuLib(str(tmp_path / "libspgpu.so"))


@pytest.mark.skipif(not os.path.exists("/root/reference/src/tests/ctest.c"), reason="reference sources not present")
def test_reference_drivers_link_against_us(tmp_path):
    """drop-in: reference src/tests/{ctest,testSparseVector}.c compile against include/
    and link against libspgpu.so with no source change"""
    for src, extra in (("ctest.c", []), ("testSparseVector.c", ["-DTEST_DOUBLE"])):
        exe = tmp_path / (src + ".out")
        cmd = ["gcc", "-w", *extra, f"-I{ROOT}/include", f"-I{ROOT}/include/core", "-I/root/reference/src",
               f"-I{CUDA_INC}", f"/root/reference/src/tests/{src}", "-o", str(exe),
               f"-L{ROOT}/spgpu_b200/lib", "-lspgpu", "-L/usr/local/cuda/lib64", "-lcudart", "-lm"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
