/*
 * HELL SpMV for sm_100a:  z = alpha*A*x + beta*y, A in "hacked ELL".
 *
 * Replaces reference kernels/hell_spmv_base.cuh:103-157 (host entry) and
 * hell_spmv_base_template.cuh:19-357 (kernels).  Same contract: hackSize is a
 * multiple of 32, hackOffsets[h] is the ELEMENT offset of hack h (no
 * terminator), rS is mandatory and slots >= rS[i] are never read, rIdx (or
 * NULL) redirects the output row, beta == 0 means y is not read, z may alias y.
 *
 * Design: one warp per 32 consecutive rows (so a warp never straddles a hack),
 * the row walk is warp_rows_dot() of spmv_slots.cuh.  One launch per call, no
 * other host API traffic (the reference also issues cudaFuncSetCacheConfig on
 * every call).  The dead "large vector" loop of the reference is not kept
 * (SURVEY 2.2); indices are 64-bit where products can pass 2^31.
 */
#include "launch.cuh"
#include "spmv_slots.cuh"

template <typename T, int UNROLL>
__global__ void __launch_bounds__(1024)
hell_spmv_kernel(T* __restrict__ z, const T* y, T alpha,
	const T* __restrict__ cM, const int* __restrict__ rP, int hackSize,
	const int* __restrict__ hackOffsets, const int* __restrict__ rS,
	const int* __restrict__ rIdx, int rows, const T* __restrict__ x, T beta,
	int baseIndex, int longCut)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	const long long warpRow = i - lane;
	if (warpRow >= rows)
		return;                       /* whole warp past the end */
	const bool live = i < rows;

	const int hack = (int)(warpRow / hackSize);
	const long long at = (long long)__ldg(hackOffsets + hack) + (warpRow % hackSize) + lane;
	const int len = live ? ld_stream(rS + i) : 0;
	const bool useBeta = Num<T>::nonzero(beta);
	const long long out = (live && rIdx) ? (long long)__ldg(rIdx + i) : i;
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = y[out];

	T acc = warp_rows_dot<T, UNROLL>(cM + at, rP + at, hackSize, hackSize, len, longCut, x, baseIndex);

	if (live)
		z[out] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
}

template <typename T, int UNROLL>
static void hell_spmv_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* cM, const int* rP, int hackSize, const int* hackOffsets,
	const int* rS, const int* rIdx, int avgNnzPerRow, int rows, const T* x,
	T beta, int baseIndex)
{
	if (rows <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int block = spgpu_block(t->hellBlock);
	const unsigned grid = spgpu_ceil_div(rows, block);
	hell_spmv_kernel<T, UNROLL><<<grid, block, 0, handle->currentStream>>>(
		z, y, alpha, cM, rP, hackSize, hackOffsets, rS, rIdx, rows, x, beta,
		baseIndex, spgpu_long_cut(t, avgNnzPerRow));
	spgpu_count_launch(handle);
}

#define SPGPU_DEFINE_HELLSPMV(S, T, U)                                        \
	extern "C" void spgpu##S##hellspmv(spgpuHandle_t handle, T* z, const T* y, \
		T alpha, const T* cM, const int* rP, int hackSize,                     \
		const int* hackOffsets, const int* rS, const int* rIdx,                \
		int avgNnzPerRow, int rows, const T* x, T beta, int baseIndex)         \
	{                                                                          \
		hell_spmv_launch<T, U>(handle, z, y, alpha, cM, rP, hackSize,          \
			hackOffsets, rS, rIdx, avgNnzPerRow, rows, x, beta, baseIndex);    \
	}

SPGPU_DEFINE_HELLSPMV(S, float, 8)
SPGPU_DEFINE_HELLSPMV(D, double, 8)
SPGPU_DEFINE_HELLSPMV(C, cuFloatComplex, 8)
SPGPU_DEFINE_HELLSPMV(Z, cuDoubleComplex, 4)
