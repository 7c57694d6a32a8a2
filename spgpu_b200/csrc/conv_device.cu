/*
 * Device-side format construction: CSR -> HELL without leaving the
 * GPU, producing bit for bit what the reference's host route (cooToEll + ellToHell,
 * reference ell.c:39-80, hell.c:46-104) produces from the same entries: rS, hackOffsets
 * (element offsets, one per hack, no terminator) and the slot placement
 * hackOffsets[h] + k*hackSize + row%hackSize.  The host route is serial O(nnz) and needs
 * an ELL intermediate of maxRowLength x rows; this one is three small kernels and a
 * prefix sum (CUB DeviceScan, toolkit header library -- not on the SpMV path).
 */
#include <cub/device/device_scan.cuh>

#include "launch.cuh"
#include "numeric.cuh"

/* rS[i] = rowPtr[i+1] - rowPtr[i]; one warp per 32 rows also leaves the warp maximum */
__global__ void __launch_bounds__(256)
csr_row_sizes_kernel(const int* __restrict__ rowPtr, int rows, int* __restrict__ rS)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < rows)
		rS[i] = rowPtr[i + 1] - rowPtr[i];
}

/* slab size of every hack = hackSize * longest row of the hack (one warp per hack) */
__global__ void __launch_bounds__(256)
hack_sizes_kernel(const int* __restrict__ rS, int rows, int hackSize, int hacks, int* __restrict__ sizes)
{
	const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (warp >= hacks)
		return;
	const long long first = (long long)warp * hackSize;
	int longest = 0;
	for (int r = lane; r < hackSize; r += 32) {
		const long long row = first + r;
		if (row < rows)
			longest = max(longest, rS[row]);
	}
	longest = __reduce_max_sync(SPGPU_FULL_MASK, longest);
	if (lane == 0)
		sizes[warp] = longest * hackSize;
}

/* one warp per row: the row's entries are read coalesced and written to their HELL slots */
template <typename T>
__global__ void __launch_bounds__(256)
csr_to_hell_scatter_kernel(const int* __restrict__ rowPtr, const int* __restrict__ cols,
	const T* __restrict__ vals, int rows, int csrBase, int hellBase, int hackSize,
	const int* __restrict__ hackOffsets, T* __restrict__ hellValues, int* __restrict__ hellIndices)
{
	const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (row >= rows)
		return;
	const int begin = rowPtr[row] - csrBase, end = rowPtr[row + 1] - csrBase;
	const long long at = (long long)hackOffsets[row / hackSize] + row % hackSize;
	for (int e = begin + lane; e < end; e += 32) {
		const long long to = at + (long long)(e - begin) * hackSize;
		hellValues[to] = vals[e];
		hellIndices[to] = cols[e] - csrBase + hellBase;
	}
}

/*
 * Step 1 (blocking): row sizes, hack offsets and the allocation size.  dRs (rows ints) and
 * dHackOffsets (ceil(rows/hackSize) ints) are device outputs; *totalElements (host) is
 * the number of elements the HELL value / index arrays need.  Returns 0 on success.
 */
extern "C" int spgpuCsrToHellLayoutDevice(spgpuHandle_t handle, int rows, const int* dRowPtr,
	int hackSize, int* dRs, int* dHackOffsets, long long* totalElements)
{
	*totalElements = 0;
	if (rows <= 0)
		return 0;
	if (hackSize <= 0 || hackSize % 32)
		return SPGPU_UNSUPPORTED;
	cudaStream_t s = handle->currentStream;
	const int hacks = (rows + hackSize - 1) / hackSize;
	csr_row_sizes_kernel<<<spgpu_ceil_div(rows, 256), 256, 0, s>>>(dRowPtr, rows, dRs);
	spgpu_count_launch(handle);

	size_t tempBytes = 0;
	cub::DeviceScan::ExclusiveSum(NULL, tempBytes, (const int*)NULL, (int*)NULL, hacks, s);
	const size_t sizesBytes = ((size_t)hacks * sizeof(int) + 255) & ~(size_t)255;
	char* scratch = (char*)spgpuScratch(handle, sizesBytes + tempBytes + 256);
	if (!scratch)
		return SPGPU_OUTOFMEMORY;
	int* sizes = reinterpret_cast<int*>(scratch);
	hack_sizes_kernel<<<spgpu_ceil_div((long long)hacks * 32, 256), 256, 0, s>>>(dRs, rows, hackSize, hacks, sizes);
	spgpu_count_launch(handle);
	cub::DeviceScan::ExclusiveSum(scratch + sizesBytes, tempBytes, sizes, dHackOffsets, hacks, s);

	int lastOffset = 0, lastSize = 0;
	cudaMemcpyAsync(&lastOffset, dHackOffsets + hacks - 1, sizeof(int), cudaMemcpyDeviceToHost, s);
	cudaMemcpyAsync(&lastSize, sizes + hacks - 1, sizeof(int), cudaMemcpyDeviceToHost, s);
	if (cudaStreamSynchronize(s) != cudaSuccess)
		return SPGPU_UNSPECIFIED;
	*totalElements = (long long)lastOffset + lastSize;
	/* the reference ABI stores element offsets as int: refuse matrices that do not fit */
	if (lastOffset < 0 || *totalElements > 2147483647LL)
		return SPGPU_UNSUPPORTED;
	return SPGPU_SUCCESS;
}

template <typename T>
static void csr_to_hell_fill(spgpuHandle_t handle, int rows, const int* dRowPtr, const int* dCols,
	const T* dVals, int csrBase, int hackSize, const int* dHackOffsets, int hellBase,
	T* dHellValues, int* dHellIndices)
{
	if (rows <= 0)
		return;
	csr_to_hell_scatter_kernel<T><<<spgpu_ceil_div((long long)rows * 32, 256), 256, 0, handle->currentStream>>>(
		dRowPtr, dCols, dVals, rows, csrBase, hellBase, hackSize, dHackOffsets, dHellValues, dHellIndices);
	spgpu_count_launch(handle);
}

/* Step 2 (asynchronous): place the entries.  Slots beyond a row's length are left untouched,
 * exactly like ellToHell. */
#define SPGPU_DEFINE_CSR2HELL(S, T)                                                      \
	extern "C" void spgpu##S##csrToHellDevice(spgpuHandle_t handle, int rows,             \
		const int* dRowPtr, const int* dCols, const T* dVals, int csrBase, int hackSize,  \
		const int* dHackOffsets, int hellBase, T* dHellValues, int* dHellIndices)         \
	{                                                                                     \
		csr_to_hell_fill<T>(handle, rows, dRowPtr, dCols, dVals, csrBase, hackSize,       \
			dHackOffsets, hellBase, dHellValues, dHellIndices);                           \
	}

SPGPU_DEFINE_CSR2HELL(S, float)
SPGPU_DEFINE_CSR2HELL(D, double)
SPGPU_DEFINE_CSR2HELL(C, cuFloatComplex)
SPGPU_DEFINE_CSR2HELL(Z, cuDoubleComplex)
