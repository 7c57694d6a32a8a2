/*
 * HDIA SpMV, bulk-async (TMA) pipelined variant for sm_100a (hdiaVariant = 4).
 *
 * An HDIA matrix is one contiguous stream of cells (hack after hack, diagonal after
 * diagonal) plus one contiguous stream of diagonal offsets.  In the direct kernel every warp
 * first needs hackOffsets (L2), then its offsets (HBM), then x, while its cell loads sit in
 * registers -- 32 resident warps cannot quite hide that chain (89 % of the HBM peak on the
 * 128^3 27-point stencil).  Here a producer warp streams tiles of 4 hacks (128 rows) --
 * cells and offsets -- into a shared-memory ring with cp.async.bulk (UBLKCP) and an mbarrier
 * per stage; 4 consumer warps (one hack each) read cells and offsets from shared memory and
 * only gather x.  Same mbarrier protocol and helpers as spmv_hell_bulk.cuh.
 * Tiles with more diagonals than a stage holds, and the last tiles (whose 16-byte aligned
 * offset window could run past the array), are done by the same warps with direct loads.
 */
#ifndef SPGPU_SPMV_HDIA_BULK_CUH_
#define SPGPU_SPMV_HDIA_BULK_CUH_

#include <climits>
#include "spmv_hell_bulk.cuh"

#define HDB_WARPS 4
#define HDB_THREADS ((HDB_WARPS + 1) * 32)

template <typename T, int HACK, int UNROLL>
__global__ void __launch_bounds__(HDB_THREADS, 2)
hdia_spmv_bulk_kernel(T* __restrict__ z, const T* y, T alpha, const T* __restrict__ dM,
	const int* __restrict__ offsets, const int* __restrict__ hackOffsets, int rows, int cols,
	const T* __restrict__ x, T beta, int capD, int stages)
{
	constexpr int G = HDB_WARPS * 32 / HACK;              /* hacks per tile (128 rows) */
	constexpr int TILE_ROWS = HDB_WARPS * 32;
	extern __shared__ __align__(128) unsigned char hb_smem[];

	const size_t valBytes = (size_t)capD * HACK * sizeof(T);
	const size_t offBytes = (size_t)(capD + 8) * sizeof(int);
	const size_t stageBytes = valBytes + offBytes;
	uint64_t* full = reinterpret_cast<uint64_t*>(hb_smem + (size_t)stages * stageBytes);
	uint64_t* empty = full + stages;

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int hacks = (rows + HACK - 1) / HACK;
	const int tiles = (hacks + G - 1) / G;
	const int totalD = __ldg(hackOffsets + hacks);

	if (threadIdx.x == 0) {
		for (int s = 0; s < stages; ++s) {
			hb_mbar_init(full + s, 1);
			hb_mbar_init(empty + s, HDB_WARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	/* both sides classify a tile from the same three numbers */
#define HDB_CLASSIFY(t, d0, d1, d0a, nOff, direct)                                         \
	const int h0_ = (t) * G, h1_ = min(h0_ + G, hacks);                                    \
	const int d0 = __ldg(hackOffsets + h0_), d1 = __ldg(hackOffsets + h1_);               \
	const int d0a = d0 & ~3;                                                               \
	const int nOff = ((d1 - d0a) + 3) & ~3;                                                \
	const bool direct = (h0_ + G > hacks) || (d1 - d0) > capD || (d0a + nOff) > totalD || d1 == d0;

	if (warp == HDB_WARPS) {
		if (lane == 0) {
			const uint64_t policy = hb_policy_evict_first();
			int stage = 0;
			unsigned phase = 0;
			for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
				HDB_CLASSIFY(t, d0, d1, d0a, nOff, direct)
				if (direct)
					continue;
				hb_mbar_wait(empty + stage, phase ^ 1u);
				unsigned char* base = hb_smem + (size_t)stage * stageBytes;
				const unsigned vb = (unsigned)(d1 - d0) * HACK * sizeof(T), ob = (unsigned)nOff * sizeof(int);
				hb_mbar_expect_tx(full + stage, vb + ob);
				hb_bulk_g2s(base, dM + (size_t)d0 * HACK, vb, full + stage, policy);
				hb_bulk_g2s(base + valBytes, offsets + d0a, ob, full + stage, policy);
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
		const int hack = (t * TILE_ROWS + warp * 32) / HACK;
		const int inHack = (warp * 32) % HACK + lane;
		const bool warpLive = t * TILE_ROWS + warp * 32 < rows;
		const bool live = row < rows;
		const unsigned colsEff = live ? (unsigned)cols : 0u;
		HDB_CLASSIFY(t, d0, d1, d0a, nOff, direct)
		(void)nOff;
		T yv = Num<T>::zero();
		if (useBeta && live)
			yv = y[row];
		T acc = Num<T>::zero();
		int first = 0, cnt = 0;
		if (warpLive) {
			first = __ldg(hackOffsets + hack);
			cnt = __ldg(hackOffsets + hack + 1) - first;
		}

		if (direct) {
			if (warpLive) {
				const T* cell = dM + (long long)first * HACK + inHack;
				const int* offs = offsets + first;
				for (int j0 = 0; j0 < cnt; j0 += 32) {
					const int mineOff = (j0 + lane < cnt) ? ld_stream(offs + j0 + lane) : INT_MIN;
					const int n = min(32, cnt - j0);
					for (int u0 = 0; u0 < n; u0 += UNROLL) {
						T a[UNROLL];
						T xv[UNROLL];
						bool on[UNROLL];
#pragma unroll
						for (int u = 0; u < UNROLL; ++u) {
							a[u] = Num<T>::zero();
							if (u0 + u < n)
								a[u] = ld_stream(cell + (long long)(j0 + u0 + u) * HACK);
						}
#pragma unroll
						for (int u = 0; u < UNROLL; ++u) {
							const int c = row + __shfl_sync(SPGPU_FULL_MASK, mineOff, u0 + u);
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
			}
		} else {
			hb_mbar_wait(full + stage, phase);
			const unsigned char* base = hb_smem + (size_t)stage * stageBytes;
			const T* sv = reinterpret_cast<const T*>(base) + (size_t)(first - d0) * HACK + inHack;
			const int* so = reinterpret_cast<const int*>(base + valBytes) + (first - d0a);
			for (int j0 = 0; j0 < cnt; j0 += UNROLL) {
				T a[UNROLL];
				T xv[UNROLL];
				bool on[UNROLL];
#pragma unroll
				for (int u = 0; u < UNROLL; ++u) {
					const bool have = (j0 + u) < cnt;             /* warp-uniform */
					a[u] = Num<T>::zero();
					on[u] = false;
					xv[u] = Num<T>::zero();
					if (have) {
						a[u] = sv[(size_t)(j0 + u) * HACK];
						const int c = row + so[j0 + u];
						on[u] = (unsigned)c < colsEff;
						if (on[u])
							xv[u] = ld_keep(x + c);
					}
				}
#pragma unroll
				for (int u = 0; u < UNROLL; ++u)
					acc = on[u] ? Num<T>::fma(a[u], xv[u], acc) : acc;
			}
			__syncwarp();
			if (lane == 0)
				hb_mbar_arrive(empty + stage);
			if (++stage == stages) { stage = 0; phase ^= 1u; }
		}

		if (live)
			z[row] = spmv_epilogue<T>(acc, alpha, beta, useBeta, yv);
	}
#undef HDB_CLASSIFY
}

#endif
