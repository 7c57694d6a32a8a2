/*
 * ELL SpMV, bulk-async (TMA) pipelined variant -- the ELL twin of
 * spmv_hell_bulk.cuh (same mbarrier ring, same warp roles; read that header
 * first).  In ELL slot k of rows [r0, r0+256) is one contiguous run
 * cM[r0 + k*pitch .. +256), so a tile is loaded with maxNnzPerRow bulk copies
 * of values, as many of column indices and one of rS, into a stage laid out
 * vals[k][256] | idx[k][256] | rs[256].  Meant for short regular rows (the whole
 * padded slab is streamed); the host only selects it when maxNnzPerRow is small
 * and close to the average.
 */
#ifndef SPGPU_SPMV_ELL_BULK_CUH_
#define SPGPU_SPMV_ELL_BULK_CUH_

#include "spmv_hell_bulk.cuh"

template <typename T, int UNROLL>
__global__ void __launch_bounds__(HB_THREADS, 2)
ell_spmv_bulk_kernel(T* __restrict__ z, const T* y, T alpha,
	const T* __restrict__ cM, const int* __restrict__ rP, int cMPitch, int rPPitch,
	const int* __restrict__ rS, const int* __restrict__ rIdx, int maxNnz, int rows,
	const T* __restrict__ x, T beta, int baseIndex, int longCut, int stages)
{
	constexpr int TILE_ROWS = HB_CONSUMER_WARPS * 32;
	extern __shared__ __align__(128) unsigned char hb_smem[];

	const size_t valBytes = (size_t)maxNnz * TILE_ROWS * sizeof(T);
	const size_t idxBytes = (size_t)maxNnz * TILE_ROWS * sizeof(int);
	const size_t stageBytes = valBytes + idxBytes + TILE_ROWS * sizeof(int);
	uint64_t* full = reinterpret_cast<uint64_t*>(hb_smem + (size_t)stages * stageBytes);
	uint64_t* empty = full + stages;

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int fullTiles = rows / TILE_ROWS;                       /* the ragged tail goes direct */
	const int tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;

	if (threadIdx.x == 0) {
		for (int s = 0; s < stages; ++s) {
			hb_mbar_init(full + s, 1);
			hb_mbar_init(empty + s, HB_CONSUMER_WARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == HB_CONSUMER_WARPS) {
		if (lane == 0) {
			const uint64_t policy = hb_policy_evict_first();
			int stage = 0;
			unsigned phase = 0;
			for (int t = blockIdx.x; t < fullTiles; t += gridDim.x) {
				const size_t r0 = (size_t)t * TILE_ROWS;
				hb_mbar_wait(empty + stage, phase ^ 1u);
				unsigned char* base = hb_smem + (size_t)stage * stageBytes;
				const unsigned rsBytes = rS ? TILE_ROWS * sizeof(int) : 0;
				hb_mbar_expect_tx(full + stage, (unsigned)(valBytes + idxBytes) + rsBytes);
				for (int k = 0; k < maxNnz; ++k) {
					hb_bulk_g2s(base + (size_t)k * TILE_ROWS * sizeof(T), cM + r0 + (size_t)k * cMPitch,
						TILE_ROWS * sizeof(T), full + stage, policy);
					hb_bulk_g2s(base + valBytes + (size_t)k * TILE_ROWS * sizeof(int), rP + r0 + (size_t)k * rPPitch,
						TILE_ROWS * sizeof(int), full + stage, policy);
				}
				if (rS)
					hb_bulk_g2s(base + valBytes + idxBytes, rS + r0, rsBytes, full + stage, policy);
				if (++stage == stages) { stage = 0; phase ^= 1u; }
			}
		}
		return;
	}

	const bool useBeta = Num<T>::nonzero(beta);
	int stage = 0;
	unsigned phase = 0;
	for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
		const int row = t * TILE_ROWS + warp * 32 + lane;
		const bool live = row < rows;
		const unsigned out = (live && rIdx) ? (unsigned)__ldg(rIdx + row) : (unsigned)row;
		T yv = Num<T>::zero();
		if (useBeta && live)
			yv = y[out];
		T acc = Num<T>::zero();

		if (t >= fullTiles) {
			if (t * TILE_ROWS + warp * 32 < rows) {
				const int len = live ? (rS ? ld_stream(rS + row) : maxNnz) : 0;
				acc = warp_rows_dot<T, UNROLL, 0>(cM + row, rP + row, cMPitch, rPPitch, len, longCut, 0, x, baseIndex);
			}
		} else {
			hb_mbar_wait(full + stage, phase);
			const unsigned char* base = hb_smem + (size_t)stage * stageBytes;
			const T* sv = reinterpret_cast<const T*>(base) + warp * 32 + lane;
			const int* si = reinterpret_cast<const int*>(base + valBytes) + warp * 32 + lane;
			const int len = rS ? reinterpret_cast<const int*>(base + valBytes + idxBytes)[warp * 32 + lane] : maxNnz;
			for (int k0 = 0; k0 < maxNnz; k0 += UNROLL) {
				int col[UNROLL];
				T a[UNROLL];
				T xv[UNROLL];
#pragma unroll
				for (int u = 0; u < UNROLL; ++u) {
					const bool on = (k0 + u) < len;
					col[u] = baseIndex;
					a[u] = Num<T>::zero();
					if (on) {
						col[u] = si[(k0 + u) * TILE_ROWS];
						a[u] = sv[(k0 + u) * TILE_ROWS];
					}
				}
#pragma unroll
				for (int u = 0; u < UNROLL; ++u) {
					const bool on = (k0 + u) < len;
					xv[u] = Num<T>::zero();
					if (on)
						xv[u] = ld_keep(x + (col[u] - baseIndex));
				}
#pragma unroll
				for (int u = 0; u < UNROLL; ++u)
					acc = Num<T>::fma(a[u], xv[u], acc);
			}
			__syncwarp();
			if (lane == 0)
				hb_mbar_arrive(empty + stage);
			if (++stage == stages) { stage = 0; phase ^= 1u; }
		}

		if (live)
			z[out] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
	}
}

#endif
