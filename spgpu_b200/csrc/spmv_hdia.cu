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
#include "spmv_hdia_body.cuh"
#include "spmv_hdia_bulk.cuh"
#include "spmv_hdia_slab.cuh"

template <typename T, int UNROLL, int HACK, int MINB, bool PREDICATED = false, int BLOCK = 128>
__global__ void __launch_bounds__(BLOCK, MINB)
hdia_spmv_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, int hackSizeRt,
	const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta, int prefetchHacks)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const HdiaArgs<T> a = { z, y, alpha, dM, offsets, hackSizeRt, hackOffsets, rows, cols, x, beta, prefetchHacks };
	T unused;
	hdia_warp_rows_value<T, UNROLL, HACK, PREDICATED>(a, (blockIdx.x * BLOCK + threadIdx.x) & ~31u, unused);
}

/*
 * Variant with shared-memory staging of the x windows (hdiaVariant = 2).  Stencil matrices
 * have runs of consecutive offsets (.., o-1, o, o+1, ..): the 32 rows of a warp then need
 * the 32+L-1 consecutive x entries x[row0+o .. row0+o+31+L-1] for the L diagonals of a
 * run.  The warp loads that window ONCE into its slice of shared memory (two coalesced
 * loads) and every diagonal of the run reads its operand at lane+d -- conflict-free --
 * instead of issuing L overlapping, mostly misaligned global loads.  Runs are found with
 * one shuffle + ballot per 32 diagonals.  Matrix cells are still loaded 8 diagonals at a
 * time before they are consumed.
 */
template <typename T, int HACK>
__global__ void __launch_bounds__(128, 8)
hdia_spmv_staged_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, int hackSizeRt,
	const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta)
{
	constexpr int U = 8;
	__shared__ T win[4][64];
	const int hackSize = HACK > 0 ? HACK : hackSizeRt;
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	const unsigned warpRow = i - lane;
	if (warpRow >= (unsigned)rows)
		return;
	T* w = win[threadIdx.x >> 5];
	const bool live = i < (unsigned)rows;
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
		const int n = min(32, diags - j0);
		const int mineOff = ((int)lane < n) ? ld_stream(offs + j0 + lane) : INT_MIN;
		const int prevOff = __shfl_up_sync(SPGPU_FULL_MASK, mineOff, 1);
		/* bit j set: diagonal j starts a run (its offset is not the previous one + 1) */
		const unsigned starts = __ballot_sync(SPGPU_FULL_MASK, (int)lane < n && (lane == 0 || mineOff != prevOff + 1));
		for (int u0 = 0; u0 < n; u0 += U) {
			const T* cp = cell + (long long)(j0 + u0) * hackSize;
			T a[U];
#pragma unroll
			for (int u = 0; u < U; ++u) {
				a[u] = Num<T>::zero();
				if (u0 + u < n)
					a[u] = ld_stream(cp + (long long)u * hackSize);
			}
			/* walk the runs that intersect diagonals [u0, u0+U) */
			int d = u0;
			const int dEnd = min(u0 + U, n);
			while (d < dEnd) {
				/* run containing d ends at the next start bit after d (or at dEnd) */
				const unsigned later = starts & ~((2u << d) - 1u);
				const int runEnd = min(later ? (__ffs(later) - 1) : n, dEnd);
				const int L = runEnd - d;
				const int o = __shfl_sync(SPGPU_FULL_MASK, mineOff, d);
				const int c0 = (int)warpRow + o + (int)lane;
				__syncwarp();
				w[lane] = ((unsigned)c0 < (unsigned)cols) ? ld_keep(x + c0) : Num<T>::zero();
				if ((int)lane < L - 1) {
					const int c1 = c0 + 32;
					w[32 + lane] = ((unsigned)c1 < (unsigned)cols) ? ld_keep(x + c1) : Num<T>::zero();
				}
				__syncwarp();
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int dd = u0 + u;
					if (dd >= d && dd < runEnd) {            /* warp-uniform */
						const int c = (int)i + o + (dd - d);
						const bool on = live && (unsigned)c < (unsigned)cols;
						const T xv = w[lane + (dd - d)];
						acc = on ? Num<T>::fma(a[u], xv, acc) : acc;
					}
				}
				d = runEnd;
			}
		}
	}

	if (live)
		z[i] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
}

/*
 * Persistent variant with software prefetch of the per-hack metadata (hdiaVariant = 5).
 * The direct kernel's warps each walk the dependent chain hackOffsets -> offsets -> x once
 * and stall on it (ncu: 37 % of the stall samples sit on the first use of the x gather).
 * Here a warp walks 32-row units u, u + W, u + 2W, ... (W = warps in the grid) and, while
 * it multiplies unit u, already holds the hackOffsets pair and the first 32 offsets of unit
 * u + W and has the hackOffsets pair of unit u + 2W in flight.
 */
template <typename T, int UNROLL, int HACK>
__global__ void __launch_bounds__(128, 8)
hdia_spmv_persistent_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, int hackSizeRt,
	const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta)
{
	const int hackSize = HACK > 0 ? HACK : hackSizeRt;
	const unsigned lane = threadIdx.x & 31;
	const unsigned units = ((unsigned)rows + 31u) >> 5;
	const unsigned W = (gridDim.x * blockDim.x) >> 5;
	const unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const bool useBeta = Num<T>::nonzero(beta);

	/* pipeline registers: metadata of the current unit, the next one, and the pair after that */
	int first0 = 0, cnt0 = 0, off0 = INT_MIN;
	int first1 = 0, cnt1 = 0;
	unsigned u = g;
	if (u < units) {
		const unsigned hk = (u * 32u) / (unsigned)hackSize;
		first0 = __ldg(hackOffsets + hk);
		cnt0 = __ldg(hackOffsets + hk + 1) - first0;
		off0 = ((int)lane < cnt0) ? ld_stream(offsets + first0 + lane) : INT_MIN;
	}
	if (u + W < units) {
		const unsigned hk = ((u + W) * 32u) / (unsigned)hackSize;
		first1 = __ldg(hackOffsets + hk);
		cnt1 = __ldg(hackOffsets + hk + 1) - first1;
	}

	for (; u < units; u += W) {
		/* prefetch: offsets of the next unit (its pair is already here), pair of the one after */
		int off1 = INT_MIN, first2 = 0, cnt2 = 0;
		if (u + W < units)
			off1 = ((int)lane < cnt1) ? ld_stream(offsets + first1 + lane) : INT_MIN;
		if (u + 2 * W < units) {
			const unsigned hk = ((u + 2 * W) * 32u) / (unsigned)hackSize;
			first2 = __ldg(hackOffsets + hk);
			cnt2 = __ldg(hackOffsets + hk + 1) - first2;
		}

		const unsigned warpRow = u * 32u;
		const unsigned i = warpRow + lane;
		const bool live = i < (unsigned)rows;
		const unsigned colsEff = live ? (unsigned)cols : 0u;
		T yv = Num<T>::zero();
		if (useBeta && live)
			yv = y[i];
		const T* cell = dM + (long long)first0 * hackSize + (warpRow % (unsigned)hackSize) + lane;
		T acc = Num<T>::zero();
		for (int j0 = 0; j0 < cnt0; j0 += 32) {
			const int mineOff = j0 == 0 ? off0
				: ((j0 + (int)lane < cnt0) ? ld_stream(offsets + first0 + j0 + lane) : INT_MIN);
			const int n = min(32, cnt0 - j0);
			for (int u0 = 0; u0 < n; u0 += UNROLL) {
				const T* cp = cell + (long long)(j0 + u0) * hackSize;
				T a[UNROLL];
				T xv[UNROLL];
				bool on[UNROLL];
#pragma unroll
				for (int k = 0; k < UNROLL; ++k) {
					a[k] = Num<T>::zero();
					if (u0 + k < n)
						a[k] = ld_stream(cp + (long long)k * hackSize);
				}
#pragma unroll
				for (int k = 0; k < UNROLL; ++k) {
					const int c = (int)i + __shfl_sync(SPGPU_FULL_MASK, mineOff, u0 + k);
					on[k] = (unsigned)c < colsEff;
					xv[k] = Num<T>::zero();
					if (on[k])
						xv[k] = ld_keep(x + c);
				}
#pragma unroll
				for (int k = 0; k < UNROLL; ++k)
					acc = on[k] ? Num<T>::fma(a[k], xv[k], acc) : acc;
			}
		}
		if (live)
			z[i] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);

		first0 = first1; cnt0 = cnt1; off0 = off1;
		first1 = first2; cnt1 = cnt2;
	}
}

template <typename T, int UNROLL, int HACK>
static bool hdia_spmv_try_bulk(spgpuHandle_t handle, T* z, const T* y, T alpha, const T* dM,
	const int* offsets, const int* hackOffsets, int rows, int cols, const T* x, T beta)
{
	const int stages = 3;
	const size_t budget = 110 * 1024 - 256;
	const size_t perDiag = (size_t)HACK * sizeof(T) + sizeof(int);
	const int capD = (int)((budget / stages - 8 * sizeof(int)) / perDiag) & ~3;
	if (capD < 8 || rows < 8 * HDB_WARPS * 32)
		return false;
	if ((((size_t)dM | (size_t)offsets) & 15) != 0)
		return false;
	if (cudaFuncSetAttribute(hdia_spmv_bulk_kernel<T, HACK, UNROLL>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024)) != cudaSuccess)
		return false;
	const size_t stageBytes = (size_t)capD * HACK * sizeof(T) + (size_t)(capD + 8) * sizeof(int);
	const size_t smem = stages * stageBytes + 2 * stages * sizeof(uint64_t);
	const int tiles = (rows + HDB_WARPS * 32 - 1) / (HDB_WARPS * 32);
	int grid = 2 * handle->multiProcessorCount;
	if (grid > tiles) grid = tiles;
	hdia_spmv_bulk_kernel<T, HACK, UNROLL><<<grid, HDB_THREADS, smem, handle->currentStream>>>(
		z, y, alpha, dM, offsets, hackOffsets, rows, cols, x, beta, capD, stages);
	spgpu_count_launch(handle);
	return true;
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
	/* the side variants index the 32 offsets a warp holds without a wrap guard: their rounds must divide 32 */
	constexpr int U8 = UNROLL == 9 ? 8 : UNROLL;
	if (t->hdiaVariant == 4) {
		bool done = false;
		if (hackSize == 32)      done = hdia_spmv_try_bulk<T, U8, 32>(handle, z, y, alpha, dM, offsets, hackOffsets, rows, cols, x, beta);
		else if (hackSize == 64) done = hdia_spmv_try_bulk<T, U8, 64>(handle, z, y, alpha, dM, offsets, hackOffsets, rows, cols, x, beta);
		if (done)
			return;
	}
	if ((t->hdiaVariant == 6 || t->hdiaVariant == 7) && hackSize == 32) {
		/* per-warp slab through the bulk-copy engine; hdiaBlock (1..32) caps the diagonals per slice */
		const int cap = (t->hdiaBlock > 0 && t->hdiaBlock <= 32) ? t->hdiaBlock : 0;
		constexpr int UX = sizeof(T) > 8 ? 8 : 16;               /* x gathers a lane keeps in flight */
		const bool done = t->hdiaVariant == 6
			? hdia_spmv_try_slab<T, UX, 6>(handle, z, y, alpha, dM, offsets, hackOffsets, rows, cols, x, beta, cap)
			: hdia_spmv_try_slab<T, 2 * UX, 4>(handle, z, y, alpha, dM, offsets, hackOffsets, rows, cols, x, beta, cap);
		if (done)
			return;
	}
	if (t->hdiaVariant == 5) {
		long long want = (long long)handle->multiProcessorCount * 8;
		if (want > (long long)grid) want = grid;
		if (hackSize == 32) hdia_spmv_persistent_kernel<T, U8, 32><<<(unsigned)want, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		else                hdia_spmv_persistent_kernel<T, U8, 0><<<(unsigned)want, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		spgpu_count_launch(handle);
		return;
	}
	if (t->hdiaVariant == 2) {
		if (hackSize == 32) hdia_spmv_staged_kernel<T, 32><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		else                hdia_spmv_staged_kernel<T, 0><<<grid, 128, 0, s>>>(z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta);
		spgpu_count_launch(handle);
		return;
	}
	/* hdiaPrefetch tuning key: waves of resident CTAs (8 per SM, 128 / hackSize hacks each) a warp looks ahead; 0 = 2, < 0 off */
	const int pfWaves = t->hdiaPrefetch == 0 ? 2 : t->hdiaPrefetch;
	const int pf = pfWaves < 0 ? 0 : pfWaves * 8 * handle->multiProcessorCount * (hackSize <= 128 ? 128 / hackSize : 1);
	/* occupancy / round-size knob (registers vs resident warps): hdiaBlock >=256 -> 48 warps, 192 -> 40, 176 -> 36,
	 * 160 -> 24 with twice the unroll, 64 -> 64-thread CTAs, 8 -> rounds of 8 diagonals, else 32 warps, rounds of 9 */
	if (hackSize == 32 && t->hdiaVariant == 3) {
		spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 32, 8, true>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
	} else if (hackSize == 32) {
		if (t->hdiaBlock >= 256)      spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 32, 12>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else if (t->hdiaBlock == 224) spgpu_launch_dep(handle, hdia_spmv_kernel<T, 4, 32, 12>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else if (t->hdiaBlock == 8)   spgpu_launch_dep(handle, hdia_spmv_kernel<T, U8, 32, 8>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else if (t->hdiaBlock == 64)  spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 32, 16, false, 64>, spgpu_ceil_div(rows, 64), 64, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else if (t->hdiaBlock == 176) spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 32, 9>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else if (t->hdiaBlock == 160) spgpu_launch_dep(handle, hdia_spmv_kernel<T, 2 * U8, 32, 6>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else if (t->hdiaBlock == 192) spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 32, 10>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
		else                          spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 32, 8>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
	} else if (hackSize == 64) {
		spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 64, 8>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
	} else {
		spgpu_launch_dep(handle, hdia_spmv_kernel<T, UNROLL, 0, 8>, grid, 128, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, x, beta, pf);
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

/* Diagonals per round.  9, not 8: a warp walks ceil(diags/UNROLL) dependent rounds (cells + x
 * gathers of a round are in flight together), and stencils come with 5, 7, 9, 19 or 27 diagonals
 * -- 27 is three rounds of 9 but four of 8 (77.8 us vs 86.0 us on the 128^3 27-point stencil,
 * i.e. 0.96 vs 0.87 of the measured HBM peak); 9 is never more rounds than 8 and still fits 64
 * registers for double. */
SPGPU_DEFINE_HDIASPMV(S, float, 9)
SPGPU_DEFINE_HDIASPMV(D, double, 9)
SPGPU_DEFINE_HDIASPMV(C, cuFloatComplex, 9)
SPGPU_DEFINE_HDIASPMV(Z, cuDoubleComplex, 4)
