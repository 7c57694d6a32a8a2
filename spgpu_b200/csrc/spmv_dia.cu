/*
 * DIA SpMV for sm_100a:  z = alpha*A*x + beta*y, A stored by diagonals.
 *
 * Replaces reference kernels/dia_spmv_base.cuh:99-142 (host entry) and
 * dia_spmv_base_template.cuh:20-216 (kernels).  Diagonal j holds A(i, i+off_j)
 * at dM[i + j*dMPitch]; a cell contributes only if 0 <= i+off_j < cols (cells
 * outside are not read).  No baseIndex, no rIdx.
 *
 * One row per lane.  For one diagonal the 32 lanes read 32 consecutive matrix
 * cells and 32 consecutive x entries, so both streams are coalesced by
 * construction; x goes through the read-only path and is served by L1/L2 when
 * neighbouring diagonals overlap.  Offsets are fetched 32 at a time by the
 * lanes of a warp and broadcast by shuffle (no shared memory, no block
 * barrier, unlike the reference's per-128-diagonal __syncthreads).
 */
#include <climits>
#include "launch.cuh"
#include "numeric.cuh"

template <typename T, int UNROLL>
__global__ void __launch_bounds__(128, 8)
dia_spmv_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, int dMPitch, int rows, int cols, int diags,
	const T* __restrict__ x, T beta)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	if (i - lane >= (unsigned)rows)
		return;
	const bool live = i < (unsigned)rows;
	/* 0 <= row+off < cols as one unsigned compare; dead lanes get cols = 0 and
	 * offsets past the last diagonal INT_MIN, so both fail it */
	const unsigned colsEff = live ? (unsigned)cols : 0u;
	const bool useBeta = Num<T>::nonzero(beta);
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = y[i];

	const long long pitch = dMPitch;
	const T* cp = dM + i;
	T acc = Num<T>::zero();

	for (int j0 = 0; j0 < diags; j0 += 32) {
		const int mineOff = (j0 + (int)lane < diags) ? __ldg(offsets + j0 + lane) : INT_MIN;
		const int n = min(32, diags - j0);
		for (int u0 = 0; u0 < n; u0 += UNROLL) {
			T a[UNROLL];
			T xv[UNROLL];
			bool on[UNROLL];
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const int off = __shfl_sync(SPGPU_FULL_MASK, mineOff, u0 + u);
				const int c = (int)i + off;
				on[u] = (unsigned)c < colsEff;
				a[u] = Num<T>::zero();
				xv[u] = Num<T>::zero();
				if (on[u]) {                       /* cells outside the matrix are not read */
					a[u] = ld_stream(cp + u * pitch);
					xv[u] = ld_keep(x + c);
				}
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
				acc = Num<T>::fma(a[u], xv[u], acc);
			cp += UNROLL * pitch;
		}
		/* cp advanced by ceil(n/UNROLL)*UNROLL diagonals; n == 32 except in the last chunk */
	}

	if (live)
		z[i] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
}

template <typename T, int UNROLL>
static void dia_spmv_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* dM, const int* offsets, int dMPitch, int rows, int cols, int diags,
	const T* x, T beta)
{
	if (rows <= 0)
		return;
	spgpu_launch_dep(handle, dia_spmv_kernel<T, UNROLL>, spgpu_ceil_div(rows, 128), 128, 
		z, y, alpha, dM, offsets, dMPitch, rows, cols, diags, x, beta);
	spgpu_count_launch(handle);
}

#define SPGPU_DEFINE_DIASPMV(S, T, U)                                         \
	extern "C" void spgpu##S##diaspmv(spgpuHandle_t handle, T* z, const T* y,  \
		T alpha, const T* dM, const int* offsets, int dMPitch, int rows,       \
		int cols, int diags, const T* x, T beta)                               \
	{                                                                          \
		dia_spmv_launch<T, U>(handle, z, y, alpha, dM, offsets, dMPitch, rows, \
			cols, diags, x, beta);                                             \
	}

SPGPU_DEFINE_DIASPMV(S, float, 8)
SPGPU_DEFINE_DIASPMV(D, double, 8)
SPGPU_DEFINE_DIASPMV(C, cuFloatComplex, 8)
SPGPU_DEFINE_DIASPMV(Z, cuDoubleComplex, 4)
