/*
 * HELL SpMV, bulk-async (TMA) pipelined variant for sm_100a.
 *
 * A HELL matrix is one contiguous stream: hack after hack, and inside a hack
 * slot row after slot row.  So instead of every lane issuing its own 8-byte
 * loads (and holding their destinations in registers while they are in flight),
 * a persistent CTA streams the matrix through shared memory with the bulk copy
 * engine:
 *
 *   tile      = TILE_HACKS consecutive hacks (256 rows at hackSize 32) = ONE
 *               contiguous range [hackOffsets[t*G], hackOffsets[(t+1)*G]) of
 *               values, the same range of column indices, and 1 KB of rS
 *   producer  = one warp; per tile it arms the stage's "full" mbarrier with the
 *               byte count and issues three cp.async.bulk (UBLKCP) copies with
 *               an L2 evict-first policy (the matrix is read once; x should own
 *               the L2)
 *   consumers = 8 warps, one hack each per tile: wait on "full", read values /
 *               indices / row sizes from shared memory (lane = row, conflict
 *               free), gather x through the read-only path, FMA, store z, then
 *               arrive on the stage's "empty" mbarrier
 *   pipeline  = STAGES stages per CTA, two CTAs per SM: the bytes in flight are
 *               bounded by shared memory (up to ~200 KB per SM), not by
 *               registers or warp count.
 *
 * Tiles that do not fit a stage (hacks with very long rows) and the tail tiles
 * (hackOffsets has no terminator entry, so the end of the last hack is not
 * known from the array) are done by the same warps with the direct-load walk of
 * spmv_slots.cuh; both sides classify a tile from the same two hackOffsets
 * entries, so the stage sequence stays consistent.
 */
#ifndef SPGPU_SPMV_HELL_BULK_CUH_
#define SPGPU_SPMV_HELL_BULK_CUH_

#include <cstdint>
#include "spmv_slots.cuh"

#define HB_CONSUMER_WARPS 8
#define HB_THREADS ((HB_CONSUMER_WARPS + 1) * 32)
#define HB_SPIN_LIMIT (1u << 28)

__device__ __forceinline__ uint32_t hb_smem_addr(const void* p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void hb_mbar_init(uint64_t* bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(hb_smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void hb_mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(hb_smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void hb_mbar_arrive(uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(hb_smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool hb_mbar_try_wait(uint64_t* bar, unsigned parity)
{
	unsigned ok;
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t}"
		: "=r"(ok) : "r"(hb_smem_addr(bar)), "r"(parity) : "memory");
	return ok != 0;
}

/* bounded wait: a pipeline bug becomes a trap (launch error), never a hang */
__device__ __forceinline__ void hb_mbar_wait(uint64_t* bar, unsigned parity)
{
	unsigned spins = 0;
	while (!hb_mbar_try_wait(bar, parity)) {
		if (++spins > HB_SPIN_LIMIT)
			__trap();
	}
}

__device__ __forceinline__ uint64_t hb_policy_evict_first()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

/* global -> shared bulk copy (SASS: UBLKCP), completion counted on `bar` */
__device__ __forceinline__ void hb_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		:: "r"(hb_smem_addr(dst)), "l"(src), "r"(bytes), "r"(hb_smem_addr(bar)), "l"(policy) : "memory");
}

/*
 * Shared-memory layout (dynamic): per stage  vals[cap] | idx[cap] | rs[G*HACK],
 * then the barriers.  `cap` (elements per stage) is chosen by the host.
 */
template <typename T, int HACK, int UNROLL>
__global__ void __launch_bounds__(HB_THREADS, 2)
hell_spmv_bulk_kernel(T* __restrict__ z, const T* y, T alpha,
	const T* __restrict__ cM, const int* __restrict__ rP,
	const int* __restrict__ hackOffsets, const int* __restrict__ rS,
	const int* __restrict__ rIdx, int rows, const T* __restrict__ x, T beta,
	int baseIndex, int longCut, int cap, int stages)
{
	constexpr int G = HB_CONSUMER_WARPS * 32 / HACK;        /* hacks per tile (256 rows) */
	constexpr int TILE_ROWS = HB_CONSUMER_WARPS * 32;
	extern __shared__ __align__(128) unsigned char hb_smem[];

	const size_t stageBytes = (size_t)cap * (sizeof(T) + sizeof(int)) + TILE_ROWS * sizeof(int);
	uint64_t* full = reinterpret_cast<uint64_t*>(hb_smem + (size_t)stages * stageBytes);
	uint64_t* empty = full + stages;

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int hacks = (rows + HACK - 1) / HACK;
	const int tiles = (hacks + G - 1) / G;

	if (threadIdx.x == 0) {
		for (int s = 0; s < stages; ++s) {
			hb_mbar_init(full + s, 1);
			hb_mbar_init(empty + s, HB_CONSUMER_WARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == HB_CONSUMER_WARPS) {
		/* ===================== producer warp (one lane issues) ===================== */
		if (lane == 0) {
			const uint64_t policy = hb_policy_evict_first();
			int stage = 0;
			unsigned phase = 0;
			for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
				const int h0 = t * G, h1 = h0 + G;
				if (h1 >= hacks)
					continue;                               /* tail tile: consumers go direct */
				const int e0 = __ldg(hackOffsets + h0);
				const int n = __ldg(hackOffsets + h1) - e0;
				if (n > cap)
					continue;                               /* oversize tile: direct */
				hb_mbar_wait(empty + stage, phase ^ 1u);
				unsigned char* base = hb_smem + (size_t)stage * stageBytes;
				const unsigned vb = (unsigned)n * sizeof(T), ib = (unsigned)n * sizeof(int), rb = TILE_ROWS * sizeof(int);
				hb_mbar_expect_tx(full + stage, vb + ib + rb);
				if (n > 0) {
					hb_bulk_g2s(base, cM + e0, vb, full + stage, policy);
					hb_bulk_g2s(base + (size_t)cap * sizeof(T), rP + e0, ib, full + stage, policy);
				}
				hb_bulk_g2s(base + (size_t)cap * (sizeof(T) + sizeof(int)), rS + (size_t)h0 * HACK, rb, full + stage, policy);
				if (++stage == stages) { stage = 0; phase ^= 1u; }
			}
		}
		return;
	}

	/* ========================== consumer warps ========================== */
	const bool useBeta = Num<T>::nonzero(beta);
	int stage = 0;
	unsigned phase = 0;
	for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
		const int h0 = t * G, h1 = h0 + G;
		const int row = t * TILE_ROWS + warp * 32 + lane;           /* this lane's row */
		const int hack = (t * TILE_ROWS + warp * 32) / HACK;
		const int inHack = (warp * 32) % HACK + lane;
		bool direct = h1 >= hacks;
		int e0 = 0;
		if (!direct) {
			e0 = __ldg(hackOffsets + h0);
			direct = (__ldg(hackOffsets + h1) - e0) > cap;
		}
		const bool live = row < rows;
		const unsigned out = (live && rIdx) ? (unsigned)__ldg(rIdx + row) : (unsigned)row;
		T yv = Num<T>::zero();
		if (useBeta && live)
			yv = y[out];
		T acc = Num<T>::zero();

		if (direct) {
			if (t * TILE_ROWS + warp * 32 < rows) {                 /* warp-uniform */
				const long long at = (long long)__ldg(hackOffsets + hack) + inHack;
				const int len = live ? ld_stream(rS + row) : 0;
				acc = warp_rows_dot<T, UNROLL, HACK>(cM + at, rP + at, HACK, HACK, len, longCut, 0, x, baseIndex);
			}
		} else {
			/* slab of this warp's hack inside the tile (local element offsets) */
			const int lo = __ldg(hackOffsets + hack) - e0;
			const int height = (__ldg(hackOffsets + hack + 1) - e0 - lo) / HACK;
			hb_mbar_wait(full + stage, phase);
			const unsigned char* base = hb_smem + (size_t)stage * stageBytes;
			const T* sv = reinterpret_cast<const T*>(base) + lo + inHack;
			const int* si = reinterpret_cast<const int*>(base + (size_t)cap * sizeof(T)) + lo + inHack;
			const int len = reinterpret_cast<const int*>(base + (size_t)cap * (sizeof(T) + sizeof(int)))[warp * 32 + lane];
			for (int k0 = 0; k0 < height; k0 += UNROLL) {
				int col[UNROLL];
				T a[UNROLL];
				T xv[UNROLL];
#pragma unroll
				for (int u = 0; u < UNROLL; ++u) {
					const bool on = (k0 + u) < len;                 /* len <= height by construction */
					col[u] = baseIndex;
					a[u] = Num<T>::zero();
					if (on) {
						col[u] = si[(k0 + u) * HACK];
						a[u] = sv[(k0 + u) * HACK];
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
			/* this warp is done reading the stage */
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
