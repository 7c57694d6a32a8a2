/*
 * Device-side format construction: CSR -> HELL without leaving the
 * GPU, producing bit for bit what the reference's host route (cooToEll + ellToHell,
 * reference ell.c:39-80, hell.c:46-104) produces from the same entries: rS, hackOffsets
 * (element offsets, one per hack, no terminator) and the slot placement
 * hackOffsets[h] + k*hackSize + row%hackSize.  The host route is serial O(nnz) and needs
 * an ELL intermediate of maxRowLength x rows; this one is three small kernels and a
 * prefix sum (CUB DeviceScan, toolkit header library -- not on the SpMV path).
 *
 * Also here: the OHELL row order (ellToOell's mergesort, reference ell.c:84-157) as one radix
 * sort, and COO -> HDIA (computeHdiaHackOffsetsFromCoo + cooToHdia, reference
 * hdia.cpp:161-349) as sort + unique + binary searches.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

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

/* hack sizes -> hackOffsets (exclusive prefix sum) and the allocation size; blocks on the stream.
 * `scratch` holds sizesBytes for the per-hack sizes followed by tempBytes of CUB workspace. */
static int hack_layout_from_rs(spgpuHandle_t handle, int rows, int hackSize, const int* dRs,
	int* dHackOffsets, char* scratch, size_t sizesBytes, size_t tempBytes, long long* totalElements)
{
	cudaStream_t s = handle->currentStream;
	const int hacks = (rows + hackSize - 1) / hackSize;
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
	if (lastOffset < 0 || *totalElements > 2147483647LL)
		return SPGPU_UNSUPPORTED;
	return SPGPU_SUCCESS;
}

/*
 * CSR -> HELL placement.  One LANE per HELL row (HELL row `row` takes CSR row rIdx[row], or row
 * itself when rIdx is NULL): for a given slot k the 32 lanes of a warp write 32 consecutive
 * elements of the slot row -- whole 128/256-byte runs -- and each lane walks its own CSR row
 * front to back, so its reads stay in the lines it already pulled into L1.  (One warp per row
 * with the lanes across the entries, the obvious mapping, leaves 25 of 32 lanes idle on a
 * 7-entry row and writes 8 bytes per 32-byte sector: 31 ms for the 512^3 Laplacian against
 * what this kernel takes, profiles/README.md.)  Rows longer than LANE_DEPTH entries are finished
 * by the whole warp striding over the rest, so a 4096-entry spike row does not serialise.
 */
#define CSR2HELL_LANE_DEPTH 64

template <typename T>
__global__ void __launch_bounds__(256)
csr_to_hell_place_kernel(const int* __restrict__ rowPtr, const int* __restrict__ cols,
	const T* __restrict__ vals, int rows, int csrBase, int hellBase, int hackSize,
	const int* __restrict__ hackOffsets, const int* __restrict__ rIdx,
	T* __restrict__ hellValues, int* __restrict__ hellIndices)
{
	const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	if (row - lane >= rows)
		return;
	const bool live = row < rows;
	int begin = 0, n = 0;
	long long at = 0;
	if (live) {
		const int src = rIdx ? rIdx[row] : (int)row;
		begin = rowPtr[src] - csrBase;
		n = rowPtr[src + 1] - csrBase - begin;
		at = (long long)hackOffsets[row / hackSize] + row % hackSize;
	}
	const int mine = min(n, CSR2HELL_LANE_DEPTH);
	const int deepest = __reduce_max_sync(SPGPU_FULL_MASK, mine);
	for (int k0 = 0; k0 < deepest; k0 += 4) {
		T v[4];
		int c[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			if (k0 + u < mine) {
				v[u] = vals[begin + k0 + u];
				c[u] = cols[begin + k0 + u];
			}
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			if (k0 + u < mine) {
				const long long to = at + (long long)(k0 + u) * hackSize;
				hellValues[to] = v[u];
				hellIndices[to] = c[u] - csrBase + hellBase;
			}
		}
	}
	/* long rows: all 32 lanes on one row */
	unsigned todo = __ballot_sync(SPGPU_FULL_MASK, n > CSR2HELL_LANE_DEPTH);
	while (todo) {
		const int r = __ffs(todo) - 1;
		todo &= todo - 1;
		const int rb = __shfl_sync(SPGPU_FULL_MASK, begin, r);
		const int rn = __shfl_sync(SPGPU_FULL_MASK, n, r);
		const long long rat = (long long)__shfl_sync(SPGPU_FULL_MASK, (unsigned long long)at, r);
		for (int k = CSR2HELL_LANE_DEPTH + lane; k < rn; k += 32) {
			const long long to = rat + (long long)k * hackSize;
			hellValues[to] = vals[rb + k];
			hellIndices[to] = cols[rb + k] - csrBase + hellBase;
		}
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
	/* the reference ABI stores element offsets as int: matrices that do not fit are refused */
	return hack_layout_from_rs(handle, rows, hackSize, dRs, dHackOffsets, scratch, sizesBytes, tempBytes, totalElements);
}

template <typename T>
static void csr_to_hell_fill(spgpuHandle_t handle, int rows, const int* dRowPtr, const int* dCols,
	const T* dVals, int csrBase, int hackSize, const int* dHackOffsets, int hellBase,
	T* dHellValues, int* dHellIndices)
{
	if (rows <= 0)
		return;
	csr_to_hell_place_kernel<T><<<spgpu_ceil_div(rows, 256), 256, 0, handle->currentStream>>>(
		dRowPtr, dCols, dVals, rows, csrBase, hellBase, hackSize, dHackOffsets, NULL, dHellValues, dHellIndices);
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


/* ======================================================================================
 * OHELL: rows ordered by length (reference ellToOell, ell.c:161-202).
 *
 * The reference's bottom-up mergesort takes the RIGHT run on ties (ell.c:93: strict `>`), so
 * its result is a total order: length descending, and among equal lengths ORIGINAL ROW INDEX
 * DESCENDING.  One descending radix sort of the 64-bit key (length << 32 | row) gives exactly
 * that; rIdx is the key's low word.
 * ====================================================================================== */

__global__ void __launch_bounds__(256)
ohell_keys_kernel(const int* __restrict__ rowPtr, int rows, unsigned long long* __restrict__ keys)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < rows)
		keys[i] = ((unsigned long long)(unsigned)(rowPtr[i + 1] - rowPtr[i]) << 32) | (unsigned)i;
}

__global__ void __launch_bounds__(256)
ohell_unpack_kernel(const unsigned long long* __restrict__ keys, int rows, int* __restrict__ rIdx, int* __restrict__ rS)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < rows) {
		const unsigned long long k = keys[i];
		rIdx[i] = (int)(unsigned)(k & 0xffffffffull);
		rS[i] = (int)(unsigned)(k >> 32);
	}
}

/*
 * Step 1 of CSR -> OHELL (blocking): dRidx[i] = CSR row stored as HELL row i (ellToOell's
 * order), dRs[i] = its length, dHackOffsets / *totalElements as for plain HELL.
 */
extern "C" int spgpuCsrToOhellLayoutDevice(spgpuHandle_t handle, int rows, const int* dRowPtr,
	int hackSize, int* dRidx, int* dRs, int* dHackOffsets, long long* totalElements)
{
	*totalElements = 0;
	if (rows <= 0)
		return 0;
	if (hackSize <= 0 || hackSize % 32)
		return SPGPU_UNSUPPORTED;
	cudaStream_t s = handle->currentStream;
	const int hacks = (rows + hackSize - 1) / hackSize;
	size_t sortBytes = 0, scanBytes = 0;
	cub::DeviceRadixSort::SortKeysDescending(NULL, sortBytes, (const unsigned long long*)NULL,
		(unsigned long long*)NULL, rows, 0, 64, s);
	cub::DeviceScan::ExclusiveSum(NULL, scanBytes, (const int*)NULL, (int*)NULL, hacks, s);
	const size_t keyBytes = ((size_t)rows * sizeof(unsigned long long) + 255) & ~(size_t)255;
	const size_t sizesBytes = ((size_t)hacks * sizeof(int) + 255) & ~(size_t)255;
	const size_t tempBytes = sortBytes > scanBytes ? sortBytes : scanBytes;
	char* scratch = (char*)spgpuScratch(handle, 2 * keyBytes + sizesBytes + tempBytes + 256);
	if (!scratch)
		return SPGPU_OUTOFMEMORY;
	unsigned long long* keysIn = reinterpret_cast<unsigned long long*>(scratch);
	unsigned long long* keysOut = reinterpret_cast<unsigned long long*>(scratch + keyBytes);
	char* rest = scratch + 2 * keyBytes;
	ohell_keys_kernel<<<spgpu_ceil_div(rows, 256), 256, 0, s>>>(dRowPtr, rows, keysIn);
	spgpu_count_launch(handle);
	size_t tb = tempBytes;
	cub::DeviceRadixSort::SortKeysDescending(rest + sizesBytes, tb, keysIn, keysOut, rows, 0, 64, s);
	ohell_unpack_kernel<<<spgpu_ceil_div(rows, 256), 256, 0, s>>>(keysOut, rows, dRidx, dRs);
	spgpu_count_launch(handle);
	return hack_layout_from_rs(handle, rows, hackSize, dRs, dHackOffsets, rest, sizesBytes, tempBytes, totalElements);
}

#define SPGPU_DEFINE_CSR2OHELL(S, T)                                                     \
	extern "C" void spgpu##S##csrToOhellDevice(spgpuHandle_t handle, int rows,            \
		const int* dRowPtr, const int* dCols, const T* dVals, int csrBase, int hackSize,  \
		const int* dHackOffsets, const int* dRidx, int hellBase, T* dHellValues,          \
		int* dHellIndices)                                                                \
	{                                                                                     \
		if (rows <= 0)                                                                    \
			return;                                                                       \
		csr_to_hell_place_kernel<T><<<spgpu_ceil_div(rows, 256), 256, 0,                  \
			handle->currentStream>>>(dRowPtr, dCols, dVals, rows, csrBase, hellBase,      \
			hackSize, dHackOffsets, dRidx, dHellValues, dHellIndices);                    \
		spgpu_count_launch(handle);                                                       \
	}

SPGPU_DEFINE_CSR2OHELL(S, float)
SPGPU_DEFINE_CSR2OHELL(D, double)
SPGPU_DEFINE_CSR2OHELL(C, cuFloatComplex)
SPGPU_DEFINE_CSR2OHELL(Z, cuDoubleComplex)

/* ======================================================================================
 * COO -> HDIA (reference computeHdiaHackOffsetsFromCoo + cooToHdia, hdia.cpp:161-349).
 *
 * The reference collects, per hack, the set of diagonals its entries lie on in a std::map
 * (ascending), stores col-row of each as the hack's offsets and copies every value to
 * cell (hackOffsets[h] + position of its diagonal, row % hackSize).  Equivalent, in
 * parallel: the sorted, de-duplicated list of keys (hack << diagBits | diagonal + rows - 1) IS
 * the offsets array (low bits) in storage order, hackOffsets[h] is the lower bound of
 * (h << diagBits) in that list, and an entry's diagonal slot is found by binary search inside its
 * hack's segment.  Entries are independent, so the COO order does not matter; duplicates of
 * one (row, col) race (the reference keeps the last one in COO order).
 * ====================================================================================== */

/* key layout: the diagonal col-row lies in (-rows, cols), so diagonal + rows - 1 needs
 * bit_width(rows + cols) bits; the radix sort then only visits the bits that can differ */
struct HdiaKeyCode {
	int diagBits;
	int bias;          /* rows - 1 */
	int keyBits;       /* diagBits + bit_width(hacks) */
};

static int bit_width_ll(long long v)
{
	int b = 0;
	while (v > 0) { ++b; v >>= 1; }
	return b > 0 ? b : 1;
}

static HdiaKeyCode hdia_key_code(int rows, int cols, int hackSize)
{
	HdiaKeyCode kc;
	kc.diagBits = bit_width_ll((long long)rows + cols);
	kc.bias = rows - 1;
	kc.keyBits = kc.diagBits + bit_width_ll((rows + hackSize - 1) / hackSize);
	return kc;
}

__global__ void __launch_bounds__(256)
hdia_keys_kernel(const int* __restrict__ cooRows, const int* __restrict__ cooCols, int nnz, int base,
	int hackSize, HdiaKeyCode kc, unsigned long long* __restrict__ keys)
{
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e < nnz) {
		const int r = cooRows[e] - base, c = cooCols[e] - base;
		keys[e] = ((unsigned long long)(unsigned)(r / hackSize) << kc.diagBits) | (unsigned long long)(unsigned)(c - r + kc.bias);
	}
}

/* hackOffsets[h] = number of unique keys below (h << diagBits), h = 0..hacks */
__global__ void __launch_bounds__(256)
hdia_hack_offsets_kernel(const unsigned long long* __restrict__ uniq, const int* __restrict__ count,
	int hacks, HdiaKeyCode kc, int* __restrict__ hackOffsets)
{
	const int h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h > hacks)
		return;
	const unsigned long long want = (unsigned long long)(unsigned)h << kc.diagBits;
	int lo = 0, hi = *count;
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (uniq[mid] < want) lo = mid + 1; else hi = mid;
	}
	hackOffsets[h] = lo;
}

__global__ void __launch_bounds__(256)
hdia_offsets_kernel(const unsigned long long* __restrict__ uniq, const int* __restrict__ count, HdiaKeyCode kc,
	int* __restrict__ offsets)
{
	const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (d < *count)
		offsets[d] = (int)(uniq[d] & ((1ull << kc.diagBits) - 1ull)) - kc.bias;
}

template <typename T>
__global__ void __launch_bounds__(256)
coo_to_hdia_scatter_kernel(const int* __restrict__ cooRows, const int* __restrict__ cooCols,
	const T* __restrict__ cooVals, int nnz, int base, int hackSize, const int* __restrict__ hackOffsets,
	const int* __restrict__ offsets, T* __restrict__ hdiaValues)
{
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nnz)
		return;
	const int r = cooRows[e] - base, c = cooCols[e] - base;
	const int h = r / hackSize, diag = c - r;
	int lo = hackOffsets[h], hi = hackOffsets[h + 1];
	while (lo < hi) {                                  /* the hack's offsets are ascending */
		const int mid = (lo + hi) >> 1;
		if (offsets[mid] < diag) lo = mid + 1; else hi = mid;
	}
	hdiaValues[(long long)lo * hackSize + r % hackSize] = cooVals[e];
}

/* sorted unique (hack, diagonal) keys of the COO entries into handle scratch; returns the
 * device pointers (valid until the next scratch user) or NULL on allocation failure */
static bool hdia_unique_keys(spgpuHandle_t handle, int hackSize, HdiaKeyCode kc, int nnz, const int* dCooRows,
	const int* dCooCols, int base, unsigned long long** uniq, int** dCount)
{
	cudaStream_t s = handle->currentStream;
	size_t sortBytes = 0, selBytes = 0;
	cub::DeviceRadixSort::SortKeys(NULL, sortBytes, (const unsigned long long*)NULL, (unsigned long long*)NULL, nnz, 0, kc.keyBits, s);
	cub::DeviceSelect::Unique(NULL, selBytes, (const unsigned long long*)NULL, (unsigned long long*)NULL, (int*)NULL, nnz, s);
	const size_t keyBytes = ((size_t)nnz * sizeof(unsigned long long) + 255) & ~(size_t)255;
	size_t tempBytes = sortBytes > selBytes ? sortBytes : selBytes;
	char* scratch = (char*)spgpuScratch(handle, 2 * keyBytes + 256 + tempBytes + 256);
	if (!scratch)
		return false;
	unsigned long long* a = reinterpret_cast<unsigned long long*>(scratch);
	unsigned long long* b = reinterpret_cast<unsigned long long*>(scratch + keyBytes);
	int* count = reinterpret_cast<int*>(scratch + 2 * keyBytes);
	void* temp = scratch + 2 * keyBytes + 256;
	hdia_keys_kernel<<<spgpu_ceil_div(nnz, 256), 256, 0, s>>>(dCooRows, dCooCols, nnz, base, hackSize, kc, a);
	spgpu_count_launch(handle);
	size_t tb = tempBytes;
	cub::DeviceRadixSort::SortKeys(temp, tb, a, b, nnz, 0, kc.keyBits, s);
	tb = tempBytes;
	cub::DeviceSelect::Unique(temp, tb, b, a, count, nnz, s);
	*uniq = a;
	*dCount = count;
	return true;
}

/*
 * Device twin of computeHdiaHackOffsetsFromCoo (blocking): fills dHackOffsets (hacks+1
 * entries, unit: diagonals) and *allocationHeight = total number of hack-diagonals.
 */
extern "C" int spgpuHdiaHackOffsetsFromCooDevice(spgpuHandle_t handle, int* allocationHeight,
	int* dHackOffsets, int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* dCooRowIndices, const int* dCooColsIndices, int cooBaseIndex)
{
	*allocationHeight = 0;
	if (hackSize <= 0 || hackSize % 32)
		return SPGPU_UNSUPPORTED;
	cudaStream_t s = handle->currentStream;
	const int hacks = (rowsCount + hackSize - 1) / hackSize;
	if (nonZerosCount <= 0) {
		if (cudaMemsetAsync(dHackOffsets, 0, (size_t)(hacks + 1) * sizeof(int), s) != cudaSuccess)
			return SPGPU_UNSPECIFIED;
		return cudaStreamSynchronize(s) == cudaSuccess ? SPGPU_SUCCESS : SPGPU_UNSPECIFIED;
	}
	unsigned long long* uniq;
	int* dCount;
	const HdiaKeyCode kc = hdia_key_code(rowsCount, columnsCount, hackSize);
	if (!hdia_unique_keys(handle, hackSize, kc, nonZerosCount, dCooRowIndices, dCooColsIndices, cooBaseIndex, &uniq, &dCount))
		return SPGPU_OUTOFMEMORY;
	hdia_hack_offsets_kernel<<<spgpu_ceil_div(hacks + 1, 256), 256, 0, s>>>(uniq, dCount, hacks, kc, dHackOffsets);
	spgpu_count_launch(handle);
	int count = 0;
	cudaMemcpyAsync(&count, dCount, sizeof(int), cudaMemcpyDeviceToHost, s);
	if (cudaStreamSynchronize(s) != cudaSuccess)
		return SPGPU_UNSPECIFIED;
	*allocationHeight = count;
	/* dM has count*hackSize cells: refuse what the int-based ABI cannot address */
	if ((long long)count * hackSize > 2147483647LL)
		return SPGPU_UNSUPPORTED;
	return SPGPU_SUCCESS;
}

template <typename T>
static int coo_to_hdia_fill(spgpuHandle_t handle, T* dHdiaValues, int* dHdiaOffsets, const int* dHackOffsets,
	int hackSize, int rows, int cols, int nnz, const int* dCooRows, const int* dCooCols, const T* dCooVals, int base)
{
	if (nnz <= 0)
		return SPGPU_SUCCESS;
	if (hackSize <= 0 || hackSize % 32)
		return SPGPU_UNSUPPORTED;
	cudaStream_t s = handle->currentStream;
	unsigned long long* uniq;
	int* dCount;
	const HdiaKeyCode kc = hdia_key_code(rows, cols, hackSize);
	if (!hdia_unique_keys(handle, hackSize, kc, nnz, dCooRows, dCooCols, base, &uniq, &dCount))
		return SPGPU_OUTOFMEMORY;
	/* at most nnz unique keys; the kernel stops at *dCount */
	hdia_offsets_kernel<<<spgpu_ceil_div(nnz, 256), 256, 0, s>>>(uniq, dCount, kc, dHdiaOffsets);
	spgpu_count_launch(handle);
	coo_to_hdia_scatter_kernel<T><<<spgpu_ceil_div(nnz, 256), 256, 0, s>>>(dCooRows, dCooCols, dCooVals, nnz, base,
		hackSize, dHackOffsets, dHdiaOffsets, dHdiaValues);
	spgpu_count_launch(handle);
	return SPGPU_SUCCESS;
}

/*
 * Device twin of cooToHdia (asynchronous): fills dHdiaOffsets (allocationHeight entries) and
 * places the values; like the reference it leaves all other cells untouched, so the caller
 * zero-fills dHdiaValues first (reference diaPerf.cpp:195-196).
 */
#define SPGPU_DEFINE_COO2HDIA(S, T)                                                      \
	extern "C" int spgpu##S##cooToHdiaDevice(spgpuHandle_t handle, T* dHdiaValues,        \
		int* dHdiaOffsets, const int* dHackOffsets, int hackSize, int rowsCount,          \
		int columnsCount, int nonZerosCount, const int* dCooRowIndices,                   \
		const int* dCooColsIndices, const T* dCooValues, int cooBaseIndex)                \
	{                                                                                     \
		return coo_to_hdia_fill<T>(handle, dHdiaValues, dHdiaOffsets, dHackOffsets,       \
			hackSize, rowsCount, columnsCount, nonZerosCount, dCooRowIndices,             \
			dCooColsIndices, dCooValues, cooBaseIndex);                                   \
	}

SPGPU_DEFINE_COO2HDIA(S, float)
SPGPU_DEFINE_COO2HDIA(D, double)
SPGPU_DEFINE_COO2HDIA(C, cuFloatComplex)
SPGPU_DEFINE_COO2HDIA(Z, cuDoubleComplex)
