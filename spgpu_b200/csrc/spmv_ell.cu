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
	grid_dependency_wait();
	grid_launch_dependents();
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

/*
 * Short regular rows (every row fits one round: maxNnzPerRow <= UNROLL, little padding) with the
 * slot count NS as a compile-time constant and ROWS rows per lane (ellRows = 1 or 2).  With 4-5
 * slots a row is ~80 bytes and a warp's whole memory phase is two dependent round trips
 * (indices -> x); what limits a matrix like the 2-D 5-point Laplacian is the bytes the SM has in
 * flight during that phase, i.e. resident warps x rows per lane.  An exact NS keeps only NS
 * (index, value) pairs in registers instead of UNROLL, which buys the occupancy: ROWS = 1 runs at
 * 64 warps per SM, ROWS = 2 (a CTA covers 256 consecutive rows; lane t owns rows t and t+128, so
 * every warp-level load stays one coalesced run) at 32-40 with twice the loads per lane.
 */
template <typename T, int NS, int ROWS, int MINB, int BLOCK = 128>
__global__ void __launch_bounds__(BLOCK, MINB * 128 / BLOCK)
ell_spmv_short_kernel(T* __restrict__ z, const T* y, T alpha,
	const T* __restrict__ cM, const int* __restrict__ rP, int cMPitch,
	int rPPitch, const int* __restrict__ rS, const int* __restrict__ rIdx,
	int rows, const T* __restrict__ x, T beta, int baseIndex)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const unsigned first = blockIdx.x * ((unsigned)BLOCK * ROWS) + threadIdx.x;
	const bool useBeta = Num<T>::nonzero(beta);
	int col[ROWS][NS];
	T a[ROWS][NS];
	int len[ROWS];
#pragma unroll
	for (int r = 0; r < ROWS; ++r) {
		const unsigned i = first + (unsigned)BLOCK * r;
		const bool live = i < (unsigned)rows;
		len[r] = live ? (rS ? ld_stream(rS + i) : NS) : 0;
#pragma unroll
		for (int u = 0; u < NS; ++u) {
			col[r][u] = baseIndex;
			a[r][u] = Num<T>::zero();
			if (live) {
				col[r][u] = ld_stream(rP + i + (long long)u * rPPitch);
				a[r][u] = ld_stream(cM + i + (long long)u * cMPitch);
			}
		}
	}
#pragma unroll
	for (int r = 0; r < ROWS; ++r) {
		const unsigned i = first + (unsigned)BLOCK * r;
		const bool live = i < (unsigned)rows;
		const unsigned out = (live && rIdx) ? (unsigned)__ldg(rIdx + i) : i;
		T yv = Num<T>::zero();
		if (useBeta && live)
			yv = y[out];
		T xv[NS];
#pragma unroll
		for (int u = 0; u < NS; ++u) {
			xv[u] = Num<T>::zero();
			if (u < len[r])
				xv[u] = ld_keep(x + (col[r][u] - baseIndex));
		}
		T acc = Num<T>::zero();
#pragma unroll
		for (int u = 0; u < NS; ++u)
			if (u < len[r])
				acc = Num<T>::fma(a[r][u], xv[u], acc);
		if (live)
			z[out] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
	}
}

template <typename T, int NS>
static void ell_spmv_short_launch(spgpuHandle_t handle, int rowsPerLane, T* z, const T* y, T alpha,
	const T* cM, const int* rP, int cMPitch, int rPPitch, const int* rS, const int* rIdx,
	int rows, const T* x, T beta, int baseIndex)
{
	/* registers ~ 20 + ROWS*NS*(1 + words per value) (ptxas -v): the CTA count that fits without spills */
	constexpr int W = sizeof(T) / 4;
	constexpr int MINB1 = NS * (1 + W) <= 16 ? 16 : NS * (1 + W) <= 24 ? 12 : 8;
	constexpr int MINB2 = 2 * NS * (1 + W) <= 32 ? 10 : 2 * NS * (1 + W) <= 48 ? 8 : 5;
	if (rowsPerLane == 3)              /* ellRows = 3 / 4: one row per lane in CTAs of 256 / 512 threads (fewer CTA launches) */
		spgpu_launch_dep(handle, ell_spmv_short_kernel<T, NS, 1, MINB1, 256>, spgpu_ceil_div(rows, 256), 256, 
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, rows, x, beta, baseIndex);
	else if (rowsPerLane == 4)
		spgpu_launch_dep(handle, ell_spmv_short_kernel<T, NS, 1, MINB1, 512>, spgpu_ceil_div(rows, 512), 512, 
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, rows, x, beta, baseIndex);
	else if (rowsPerLane == 5)         /* 64-thread CTAs */
		spgpu_launch_dep(handle, ell_spmv_short_kernel<T, NS, 1, MINB1, 64>, spgpu_ceil_div(rows, 64), 64, 
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, rows, x, beta, baseIndex);
	else if (rowsPerLane >= 2)
		spgpu_launch_dep(handle, ell_spmv_short_kernel<T, NS, 2, MINB2>, spgpu_ceil_div(rows, 256), 128, 
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, rows, x, beta, baseIndex);
	else
		spgpu_launch_dep(handle, ell_spmv_short_kernel<T, NS, 1, MINB1>, spgpu_ceil_div(rows, 128), 128, 
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, rows, x, beta, baseIndex);
	spgpu_count_launch(handle);
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
	    (rS == NULL || 4LL * avgNnzPerRow >= 3LL * maxNnzPerRow) &&
	    /* the lanes past the last row read too: the pitches must cover the last warp */
	    (long long)cMPitch >= (((long long)rows + 31) & ~31LL) && (long long)rPPitch >= (((long long)rows + 31) & ~31LL))
		allocated = maxNnzPerRow;
	if (allocated > 0 && t->ellRows >= 0) {
#define SPGPU_ELL_SHORT(NS) case NS: ell_spmv_short_launch<T, NS>(handle, t->ellRows, z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, rows, x, beta, baseIndex); return;
		switch (allocated) {
			SPGPU_ELL_SHORT(1) SPGPU_ELL_SHORT(2) SPGPU_ELL_SHORT(3) SPGPU_ELL_SHORT(4)
			SPGPU_ELL_SHORT(5) SPGPU_ELL_SHORT(6) SPGPU_ELL_SHORT(7) SPGPU_ELL_SHORT(8)
		default: break;
		}
#undef SPGPU_ELL_SHORT
	}
	bool dense = !Num<T>::is_complex;
	if (t->hellBlock >= 256) dense = true;
	else if (t->hellBlock > 0 && t->hellBlock <= 64) dense = false;
	if (dense)
		spgpu_launch_dep(handle, ell_spmv_kernel<T, UNROLL, 12>, grid, block, 
			z, y, alpha, cM, rP, cMPitch, rPPitch, rS, rIdx, maxNnzPerRow, rows, x,
			beta, baseIndex, spgpu_long_cut(t, avgNnzPerRow), allocated);
	else
		spgpu_launch_dep(handle, ell_spmv_kernel<T, UNROLL, 8>, grid, block, 
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
