/*
 * ELL SpMV for sm_100a:  z = alpha*A*x + beta*y, A in pitched column-major
 * ELLPACK.
 *
 * Replaces reference kernels/ell_spmv_base.cuh:99-146 (host entry),
 * ell_spmv_base_template.cuh:178-425 (rS given) and ell_spmv_base_nors.cuh
 * (rS == NULL: every row uses maxNnzPerRow slots, padding = value 0 with a
 * valid index).  cMPitch / rPPitch are in elements and may differ.
 *
 * Same warp-per-32-rows walk as HELL (spmv_slots.cuh) with base = row and
 * stride = pitch; the product k*pitch is formed in 64 bits (134 M rows x 7
 * slots already passes 2^31).
 */
#include "launch.cuh"
#include "spmv_slots.cuh"
#include "spmv_ell_bulk.cuh"

template <typename T, int UNROLL, int MINB>
__global__ void __launch_bounds__(128, MINB)
ell_spmv_kernel(T* __restrict__ z, const T* y, T alpha,
	const T* __restrict__ cM, const int* __restrict__ rP, int cMPitch,
	int rPPitch, const int* __restrict__ rS, const int* __restrict__ rIdx,
	int maxNnzPerRow, int rows, const T* __restrict__ x, T beta, int baseIndex,
	int longCut, int allocated)
{
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	if (i - lane >= (unsigned)rows)
		return;
	const bool live = i < (unsigned)rows;

	const int len = live ? (rS ? ld_stream(rS + i) : maxNnzPerRow) : 0;
	const bool useBeta = Num<T>::nonzero(beta);
	const unsigned out = (live && rIdx) ? (unsigned)__ldg(rIdx + i) : i;
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = y[out];

	T acc = warp_rows_dot<T, UNROLL, 0>(cM + i, rP + i, cMPitch, rPPitch, len, longCut, allocated, x, baseIndex);

	if (live)
		z[out] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
}

/* bulk-async variant (hellVariant = 3): short regular rows, aligned arrays */
template <typename T, int UNROLL>
static bool ell_spmv_try_bulk(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* cM, const int* rP, int cMPitch, int rPPitch, const int* rS,
	const int* rIdx, int maxNnzPerRow, int rows, const T* x, T beta,
	int baseIndex, int longCut)
{
	const int tileRows = HB_CONSUMER_WARPS * 32;
	const size_t stageBytes = (size_t)maxNnzPerRow * tileRows * (sizeof(T) + sizeof(int)) + tileRows * sizeof(int);
	int stages = (int)((110 * 1024 - 256) / stageBytes);
	if (stages > 4) stages = 4;
	if (stages < 2 || maxNnzPerRow <= 0 || rows < 8 * tileRows)
		return false;
	if ((((size_t)cM | (size_t)rP | (size_t)rS) & 15) != 0 ||
	    ((size_t)cMPitch * sizeof(T)) % 16 != 0 || ((size_t)rPPitch * sizeof(int)) % 16 != 0)
		return false;
	if (cudaFuncSetAttribute(ell_spmv_bulk_kernel<T, UNROLL>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024)) != cudaSuccess)
		return false;
	const size_t smem = stages * stageBytes + 2 * stages * sizeof(uint64_t);
	const int tiles = (rows + tileRows - 1) / tileRows;
	int grid = 2 * handle->multiProcessorCount;
	if (grid > tiles) grid = tiles;
	ell_spmv_bulk_kernel<T, UNROLL><<<grid, HB_THREADS, smem, handle->currentStream>>>(
		z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, maxNnzPerRow, rows, x, beta, baseIndex, longCut, stages);
	spgpu_count_launch(handle);
	return true;
}

template <typename T, int UNROLL>
static void ell_spmv_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* cM, const int* rP, int cMPitch, int rPPitch, const int* rS,
	const int* rIdx, int avgNnzPerRow, int maxNnzPerRow, int rows, const T* x,
	T beta, int baseIndex)
{
	if (rows <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int block = 128;
	const unsigned grid = spgpu_ceil_div(rows, block);
	if (t->hellVariant == 3 &&
	    ell_spmv_try_bulk<T, UNROLL>(handle, z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, maxNnzPerRow,
			rows, x, beta, baseIndex, spgpu_long_cut(t, avgNnzPerRow)))
		return;
	/* every slot below maxNnzPerRow exists in an ELL allocation, so short regular
	 * matrices can read them without waiting for rS (spmv_slots.cuh); only worth it
	 * when little of that is padding (avg close to max) */
	int allocated = 0;
	if (t->hellVariant != 1 && maxNnzPerRow > 0 && maxNnzPerRow <= UNROLL &&
	    (rS == NULL || 4LL * avgNnzPerRow >= 3LL * maxNnzPerRow))
		allocated = maxNnzPerRow;
	bool dense = !Num<T>::is_complex;
	if (t->hellBlock >= 256) dense = true;
	else if (t->hellBlock > 0 && t->hellBlock <= 64) dense = false;
	if (dense)
		ell_spmv_kernel<T, UNROLL, 12><<<grid, block, 0, handle->currentStream>>>(
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, maxNnzPerRow, rows, x,
			beta, baseIndex, spgpu_long_cut(t, avgNnzPerRow), allocated);
	else
		ell_spmv_kernel<T, UNROLL, 8><<<grid, block, 0, handle->currentStream>>>(
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, maxNnzPerRow, rows, x,
			beta, baseIndex, spgpu_long_cut(t, avgNnzPerRow), allocated);
	spgpu_count_launch(handle);
}

#define SPGPU_DEFINE_ELLSPMV(S, T, U)                                         \
	extern "C" void spgpu##S##ellspmv(spgpuHandle_t handle, T* z, const T* y,  \
		T alpha, const T* cM, const int* rP, int cMPitch, int rPPitch,         \
		const int* rS, const int* rIdx, int avgNnzPerRow, int maxNnzPerRow,    \
		int rows, const T* x, T beta, int baseIndex)                           \
	{                                                                          \
		ell_spmv_launch<T, U>(handle, z, y, alpha, cM, rP, cMPitch, rPPitch,   \
			rS, rIdx, avgNnzPerRow, maxNnzPerRow, rows, x, beta, baseIndex);   \
	}

SPGPU_DEFINE_ELLSPMV(S, float, 8)
SPGPU_DEFINE_ELLSPMV(D, double, 8)
SPGPU_DEFINE_ELLSPMV(C, cuFloatComplex, 8)
SPGPU_DEFINE_ELLSPMV(Z, cuDoubleComplex, 4)
