/*
 * HDIA SpMV for sm_100a:  z = alpha*A*x + beta*y, A in "hacked DIA".
 *
 * Replaces reference kernels/hdia_spmv_base.cuh:99-145 (host entry) and
 * hdia_spmv_base_template.cuh:19-252 (kernels).  Hack h (hackSize rows, a
 * multiple of 32) owns diagonals [hackOffsets[h], hackOffsets[h+1]) of
 * offsets[]; cell (d, r) is dM[d*hackSize + r]; a cell contributes only if
 * 0 <= i + offsets[d] < cols.  hackOffsets has hacks+1 entries.
 *
 * The whole matrix is ONE contiguous stream (hack after hack, diagonal after
 * diagonal), so consecutive warps read consecutive memory.  One warp owns 32
 * consecutive rows of one hack; its diagonal offsets are loaded 32 at a time by
 * the lanes and broadcast by shuffle (the reference stages them in shared
 * memory with only `volatile` between writer and readers).
 * Per diagonal the warp reads 32 adjacent cells and 32 adjacent x entries.
 */
#include "launch.cuh"
#include "numeric.cuh"

template <typename T, int UNROLL>
__global__ void __launch_bounds__(1024)
hdia_spmv_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, int hackSize,
	const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	const long long warpRow = i - lane;
	if (warpRow >= rows)
		return;
	const bool live = i < rows;
	const bool useBeta = Num<T>::nonzero(beta);
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = y[i];

	const int hack = (int)(warpRow / hackSize);
	const int first = __ldg(hackOffsets + hack);
	const int diags = __ldg(hackOffsets + hack + 1) - first;
	const T* cell = dM + (long long)first * hackSize + (warpRow % hackSize) + lane;
	const int* offs = offsets + first;
	T acc = Num<T>::zero();

	for (int j0 = 0; j0 < diags; j0 += 32) {
		const int mineOff = (j0 + lane < diags) ? ld_stream(offs + j0 + lane) : 0;
		const int n = min(32, diags - j0);
		for (int u0 = 0; u0 < n; u0 += UNROLL) {
			T a[UNROLL];
			T xv[UNROLL];
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const int jj = u0 + u;
				const int off = __shfl_sync(SPGPU_FULL_MASK, mineOff, jj & 31);
				const long long c = i + off;
				const bool on = live && jj < n && c >= 0 && c < cols;
				a[u] = on ? ld_stream(cell + (long long)(j0 + jj) * hackSize) : Num<T>::zero();
				xv[u] = on ? ld_keep(x + c) : Num<T>::zero();
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
				acc = Num<T>::fma(a[u], xv[u], acc);
		}
	}

	if (live)
		z[i] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
}

template <typename T, int UNROLL>
static void hdia_spmv_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* dM, const int* offsets, int hackSize, const int* hackOffsets,
	int rows, int cols, const T* x, T beta)
{
	if (rows <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int block = spgpu_block(t->hdiaBlock);
	hdia_spmv_kernel<T, UNROLL><<<spgpu_ceil_div(rows, block), block, 0, handle->currentStream>>>(
		z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
	spgpu_count_launch(handle);
}

#define SPGPU_DEFINE_HDIASPMV(S, T, U)                                        \
	extern "C" void spgpu##S##hdiaspmv(spgpuHandle_t handle, T* z, const T* y, \
		T alpha, const T* dM, const int* offsets, int hackSize,                \
		const int* hackOffsets, int rows, int cols, const T* x, T beta)        \
	{                                                                          \
		hdia_spmv_launch<T, U>(handle, z, y, alpha, dM, offsets, hackSize,     \
			hackOffsets, rows, cols, x, beta);                                 \
	}

SPGPU_DEFINE_HDIASPMV(S, float, 8)
SPGPU_DEFINE_HDIASPMV(D, double, 8)
SPGPU_DEFINE_HDIASPMV(C, cuFloatComplex, 8)
SPGPU_DEFINE_HDIASPMV(Z, cuDoubleComplex, 4)
