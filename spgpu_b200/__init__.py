"""spgpu_b200 -- B200-native (sm_100a) implementation of spGPU's SpMV hot path.

The product is the C-ABI shared library `spgpu_b200/lib/libspgpu.so` (CUDA
kernels + C host layer under `spgpu_b200/csrc/`, headers under `include/`).
This package is the thin Python side: a ctypes binding of that ABI (`capi`),
numpy wrappers of the host conversions (`formats`), synthetic matrix generators
for the benchmark configurations (`generators`, `device_build`) and the
row-partitioned multi-GPU layer (`mg`).
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
