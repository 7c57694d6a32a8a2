/*
 * HDIA SpMV, per-warp slab variant for sm_100a (hdiaVariant = 6 / 7), hackSize 32.
 *
 * With hackSize 32 a warp owns one whole hack, and the cells of a hack are ONE contiguous
 * run of diags*32 elements at dM + hackOffsets[h]*32.  The direct kernel fetches that run 8-9
 * diagonals at a time into registers, so a warp walks ceil(diags/9) dependent rounds after
 * the hackOffsets -> offsets head of the chain (5-6 DRAM round trips for a 27-point stencil)
 * and 32 resident warps only just cover the HBM latency.  Here lane 0 hands the whole run
 * to the bulk-copy engine (cp.async.bulk, SASS UBLKCP) as soon as hackOffsets[h] is known:
 * one instruction puts the warp's complete slab in flight (6.9 KB for 27 double diagonals),
 * destination = the warp's private slice of shared memory, completion = the warp's private
 * mbarrier.  No registers hold matrix cells, so a lane can keep UX (16 or 32) x gathers in
 * flight instead of 8, and the chain shrinks to hackOffsets -> offsets -> x.
 * Nothing is shared between warps: no producer warp, no __syncthreads, no stage ring
 * (contrast spmv_hdia_bulk.cuh, whose 4 consumer warps per CTA could not cover the gathers).
 *
 * Hacks with more than capD diagonals are walked capD diagonals at a time through the same
 * slice.  A slab always exists in full (HDIA allocates whole hacks), also for the last,
 * ragged hack.
 *
 * Measured on cfg2 (profiles/README.md): 98 us (16 gathers) / 114 us (32) against 78 us for
 * the direct kernel with rounds of 9 -- 8 KB of shared memory per warp caps the SM at 24 warps
 * and the carve-out leaves little L1 for the overlapping x windows, which costs more than
 * the shorter chain wins.  Not the default; kept selectable (and tested) as the record of it.
 */
#ifndef SPGPU_SPMV_HDIA_SLAB_CUH_
#define SPGPU_SPMV_HDIA_SLAB_CUH_

#include <climits>
#include "spmv_hell_bulk.cuh"

#define HDS_WARPS 4

template <typename T, int UX, int MINB>
__global__ void __launch_bounds__(HDS_WARPS * 32, MINB)
hdia_spmv_slab_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta, int capD)
{
	static_assert(32 % UX == 0, "a round must not straddle the 32 offsets a warp holds");
	extern __shared__ __align__(128) unsigned char hs_smem[];

	const unsigned warp = threadIdx.x >> 5;
	const unsigned lane = threadIdx.x & 31;
	const unsigned hack = blockIdx.x * HDS_WARPS + warp;
	const unsigned warpRow = hack * 32u;
	if (warpRow >= (unsigned)rows)
		return;
	T* slab = reinterpret_cast<T*>(hs_smem) + (size_t)warp * capD * 32;
	uint64_t* bar = reinterpret_cast<uint64_t*>(hs_smem + (size_t)HDS_WARPS * capD * 32 * sizeof(T)) + warp;

	if (lane == 0) {
		hb_mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	const unsigned i = warpRow + lane;
	const bool live = i < (unsigned)rows;
	const unsigned colsEff = live ? (unsigned)cols : 0u;
	const bool useBeta = Num<T>::nonzero(beta);
	const int first = __ldg(hackOffsets + hack);
	const int cnt = __ldg(hackOffsets + hack + 1) - first;
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = y[i];
	T acc = Num<T>::zero();
	unsigned phase = 0;

	for (int j0 = 0; j0 < cnt; j0 += capD) {
		const int n = min(capD, cnt - j0);                       /* <= 32 */
		if (lane == 0) {
			const unsigned bytes = (unsigned)n * 32u * (unsigned)sizeof(T);
			hb_mbar_expect_tx(bar, bytes);
			hb_bulk_g2s(slab, dM + (size_t)(first + j0) * 32, bytes, bar, hb_policy_evict_first());
		}
		const int mineOff = ((int)lane < n) ? ld_stream(offsets + first + j0 + lane) : INT_MIN;
		bool arrived = false;
		for (int u0 = 0; u0 < n; u0 += UX) {
			T xv[UX];
			unsigned onMask = 0;
#pragma unroll
			for (int u = 0; u < UX; ++u) {
				/* lanes >= n hold INT_MIN, which fails the range test */
				const int c = (int)i + __shfl_sync(SPGPU_FULL_MASK, mineOff, u0 + u);
				const bool on = (unsigned)c < colsEff;
				xv[u] = Num<T>::zero();
				if (on)
					xv[u] = ld_keep(x + c);
				onMask |= (on ? 1u : 0u) << u;
			}
			if (!arrived) {
				hb_mbar_wait(bar, phase);
				arrived = true;
			}
			const T* sp = slab + u0 * 32 + lane;
#pragma unroll
			for (int u = 0; u < UX; ++u)
				if (onMask & (1u << u))
					acc = Num<T>::fma(sp[u * 32], xv[u], acc);
		}
		phase ^= 1u;
		__syncwarp();                                            /* slice is free for the next copy */
	}

	if (live)
		z[i] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
}

/* diagonals per warp slice: 8 KB per warp (32 KB per CTA), at most the 32 offsets a warp holds */
template <typename T>
static inline int hdia_slab_cap()
{
	const int cap = (int)(8192 / (32 * sizeof(T)));
	return cap > 32 ? 32 : cap;
}

template <typename T, int UX, int MINB>
static bool hdia_spmv_try_slab(spgpuHandle_t handle, T* z, const T* y, T alpha, const T* dM,
	const int* offsets, const int* hackOffsets, int rows, int cols, const T* x, T beta, int capOverride)
{
	if (((size_t)dM & 15) != 0)
		return false;
	int capD = hdia_slab_cap<T>();
	if (capOverride > 0 && capOverride < capD)
		capD = capOverride;
	const size_t smem = (size_t)HDS_WARPS * capD * 32 * sizeof(T) + HDS_WARPS * sizeof(uint64_t);
	/* function attributes are per device: set on every call (a process may drive several GPUs) */
	if (cudaFuncSetAttribute(hdia_spmv_slab_kernel<T, UX, MINB>,
			cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) != cudaSuccess)
		return false;
	hdia_spmv_slab_kernel<T, UX, MINB><<<spgpu_ceil_div(rows, HDS_WARPS * 32), HDS_WARPS * 32, smem, handle->currentStream>>>(
		z, y, alpha, dM, offsets, hackOffsets, rows, cols, x, beta, capD);
	spgpu_count_launch(handle);
	return true;
}

#endif
