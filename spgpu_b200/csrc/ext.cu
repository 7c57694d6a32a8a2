/*
 * Additive host-side helpers (include/spgpu_ext.h): CUDA IPC + raw allocations for the
 * one-process-per-GPU ranks of the multi-GPU layer.  The device code lives in ext_krylov.cu
 * (device-scalar Krylov kernels) and ext_halo.cu (halo exchange, fused SpMV + halo, all-reduce).
 */
#include <cstring>
#include <cuda_runtime.h>
#include "spgpu_internal.h"
#include "spgpu_ext.h"

extern "C" int spgpuIpcGetHandle(void* devPtr, void* handle64)
{
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, devPtr);
	if (e != cudaSuccess)
		return (int)e;
	memcpy(handle64, &h, sizeof(h));
	return 0;
}

extern "C" int spgpuIpcOpenHandle(const void* handle64, void** devPtr)
{
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	return (int)cudaIpcOpenMemHandle(devPtr, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int spgpuIpcCloseHandle(void* devPtr)
{
	return (int)cudaIpcCloseMemHandle(devPtr);
}

extern "C" int spgpuDeviceAlloc(void** devPtr, size_t bytes)
{
	return (int)cudaMalloc(devPtr, bytes);
}

extern "C" int spgpuDeviceFree(void* devPtr)
{
	return (int)cudaFree(devPtr);
}
