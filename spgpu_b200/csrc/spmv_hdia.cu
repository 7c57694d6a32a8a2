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
#include <climits>
#include "launch.cuh"
#include "numeric.cuh"

/*
 * HACK > 0: hackSize known at compile time -> cell addresses are base +
 * immediate.  All index arithmetic is 32-bit: the in-range test
 * 0 <= row+off < cols is ONE unsigned compare; lanes past the last row get
 * cols = 0 and offsets past a hack's last diagonal get INT_MIN, so both fail
 * that same compare without extra predicates.  The matrix cells of a round are
 * loaded without waiting for the offsets (they all exist in the slab); only the
 * x gather and the FMA depend on the in-range test.
 */
template <typename T, int UNROLL, int HACK, int MINB>
__global__ void __launch_bounds__(128, MINB)
hdia_spmv_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, int hackSizeRt,
	const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta)
{
	const int hackSize = HACK > 0 ? HACK : hackSizeRt;
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	const unsigned warpRow = i - lane;
	if (warpRow >= (unsigned)rows)
		return;
	const bool live = i < (unsigned)rows;
	const unsigned colsEff = live ? (unsigned)cols : 0u;
	const bool useBeta = Num<T>::nonzero(beta);
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = y[i];

	const unsigned hack = warpRow / (unsigned)hackSize;
	const int first = __ldg(hackOffsets + hack);
	const int diags = __ldg(hackOffsets + hack + 1) - first;
	const T* cell = dM + (long long)first * hackSize + (warpRow % (unsigned)hackSize) + lane;
	const int* offs = offsets + first;
	T acc = Num<T>::zero();

	for (int j0 = 0; j0 < diags; j0 += 32) {
		const int mineOff = (j0 + (int)lane < diags) ? ld_stream(offs + j0 + lane) : INT_MIN;
		const int n = min(32, diags - j0);
		for (int u0 = 0; u0 < n; u0 += UNROLL) {
			const T* cp = cell + (long long)(j0 + u0) * hackSize;
			T a[UNROLL];
			T xv[UNROLL];
			bool on[UNROLL];
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				a[u] = Num<T>::zero();
				if (u0 + u < n)                       /* warp-uniform: cell exists */
					a[u] = ld_stream(cp + (long long)u * hackSize);
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const int off = __shfl_sync(SPGPU_FULL_MASK, mineOff, u0 + u);
				const int c = (int)i + off;
				on[u] = (unsigned)c < colsEff;
				xv[u] = Num<T>::zero();
				if (on[u])
					xv[u] = ld_keep(x + c);
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
				acc = on[u] ? Num<T>::fma(a[u], xv[u], acc) : acc;
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
	const unsigned grid = spgpu_ceil_div(rows, 128);
	cudaStream_t s = handle->currentStream;
	/* occupancy knob (registers vs resident warps): hdiaBlock >=256 -> 48 warps, 192 -> 40, else 32 */
	if (hackSize == 32) {
		if (t->hdiaBlock >= 256)      hdia_spmv_kernel<T, UNROLL, 32, 12><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		else if (t->hdiaBlock == 224) hdia_spmv_kernel<T, 4, 32, 12><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		else if (t->hdiaBlock == 192) hdia_spmv_kernel<T, UNROLL, 32, 10><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		else                          hdia_spmv_kernel<T, UNROLL, 32, 8><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
	} else if (hackSize == 64) {
		hdia_spmv_kernel<T, UNROLL, 64, 8><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
	} else {
		hdia_spmv_kernel<T, UNROLL, 0, 8><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
	}
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
