/*
 * The per-warp body of the direct HELL SpMV, shared by the plain kernel
 * (spmv_hell.cu) and the kernel fused with the NVLink halo exchange (ext.cu).
 * `warpRow` = first row of the 32 rows this warp owns; returns without doing
 * anything when the warp lies past the last row.
 */
#ifndef SPGPU_SPMV_HELL_BODY_CUH_
#define SPGPU_SPMV_HELL_BODY_CUH_

#include "spmv_slots.cuh"

template <typename T>
struct HellArgs {
	T* z;
	const T* y;
	T alpha;
	const T* cM;
	const int* rP;
	int hackSize;
	const int* hackOffsets;
	const int* rS;
	const int* rIdx;
	int rows;
	const T* x;
	T beta;
	int baseIndex;
	int longCut;
	int speculate;
	/* split mode (0 = off): rows are walked by their own warp only up to splitT slots; deeper
	 * slots are cut into chunks of splitT and queued for hell_tail_kernel (spmv_hell.cu) */
	int splitT;
	unsigned* workHeader;        /* [0] items queued, [1] items taken, [2] units that queued, [3] tail warps done;
	                              * lives in the handle, zero between calls (the tail kernel's last warp resets it) */
	uint4* workItems;            /* (32-row unit, chunk, its unit's fold entry, -) */
	int workCap;
	uint4* foldList;             /* (32-row unit, first item, items, items finished) per unit that queued */
	T* partials;                 /* one 32-lane partial sum per item */
	/* > 0: every warp asks L2 for the hackOffsets entry this many hacks ahead.  A warp's life is three dependent
	 * memory round trips -- hackOffsets -> slab -> x -- and the first one is a 4-byte read per 32 rows that nobody
	 * has touched before: HBM latency for 1 % of the bytes.  Prefetched a couple of waves ahead it is an L2 hit. */
	int prefetchHacks;
};

#define SPGPU_WORK_INVALID 0xffffffffu

/* returns the value stored for this lane's row in `zval` (zero for lanes without a row) */
template <typename T, int UNROLL, int HACK, class XG>
__device__ __forceinline__ void hell_warp_rows_value_x(const HellArgs<T>& a, unsigned warpRow, T& zval, const XG xg)
{
	zval = Num<T>::zero();
	const int hackSize = HACK > 0 ? HACK : a.hackSize;
	const unsigned lane = threadIdx.x & 31;
	if (warpRow >= (unsigned)a.rows)
		return;                       /* whole warp past the end */
	const unsigned i = warpRow + lane;
	const bool live = i < (unsigned)a.rows;

	const unsigned hack = warpRow / (unsigned)hackSize;
	const unsigned lastHack = ((unsigned)a.rows - 1u) / (unsigned)hackSize;
	if (a.prefetchHacks > 0 && lane == 0) {
		const unsigned ahead = hack + (unsigned)a.prefetchHacks;
		if (ahead <= lastHack)
			prefetch_l2(a.hackOffsets + ahead);
	}
	const int slab = __ldg(a.hackOffsets + hack);
	/* slab height of this hack = slots that exist for all of its rows; the last
	 * hack has no terminator entry, so it takes the predicated path */
	int allocated = 0;
	if (a.speculate && hack < lastHack)
		allocated = (__ldg(a.hackOffsets + hack + 1) - slab) / hackSize;
	int len = live ? ld_stream(a.rS + i) : 0;
	if (a.splitT > 0) {
		/* long hack: queue the slots beyond splitT as independent chunks so that other warps
		 * of the grid (the tail kernel) share a walk that would otherwise be this warp's
		 * serial critical path */
		const int longest = __reduce_max_sync(SPGPU_FULL_MASK, len);
		if (longest > a.splitT) {
			const int nchunks = (longest - 1) / a.splitT;            /* chunks c = 0.. cover [ (c+1)T, (c+2)T ) */
			unsigned base = 0, entry = 0;
			if (lane == 0)
				base = atomicAdd(a.workHeader, (unsigned)nchunks);
			base = __shfl_sync(SPGPU_FULL_MASK, base, 0);
			const bool fits = base + (unsigned)nchunks <= (unsigned)a.workCap;
			if (fits && lane == 0) {                                 /* fits => at most workCap units get here */
				entry = atomicAdd(a.workHeader + 2, 1u);
				a.foldList[entry] = make_uint4(warpRow >> 5, base, (unsigned)nchunks, 0u);
			}
			entry = __shfl_sync(SPGPU_FULL_MASK, entry, 0);
			for (unsigned c = lane; c < (unsigned)nchunks; c += 32)
				if (base + c < (unsigned)a.workCap)
					a.workItems[base + c] = fits ? make_uint4(warpRow >> 5, c, entry, 0u) : make_uint4(SPGPU_WORK_INVALID, 0u, 0u, 0u);
			if (fits)
				len = min(len, a.splitT);                            /* else: queue full, walk it all here */
		}
	}
	const bool useBeta = Num<T>::nonzero(a.beta);
	const unsigned out = (live && a.rIdx) ? (unsigned)__ldg(a.rIdx + i) : i;
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = a.y[out];

	const long long at = (long long)slab + (warpRow % (unsigned)hackSize) + lane;
	T acc = warp_rows_dot_x<T, UNROLL, HACK, XG>(a.cM + at, a.rP + at, hackSize, hackSize, len, a.longCut,
		allocated, xg, a.baseIndex);

	if (live) {
		zval = spmv_epilogue<T>(acc, a.alpha, a.beta, useBeta, yv);
		a.z[out] = zval;
	}
}

template <typename T, int UNROLL, int HACK>
__device__ __forceinline__ void hell_warp_rows_value(const HellArgs<T>& a, unsigned warpRow, T& zval)
{
	const XPlain<T> xg = { a.x };
	hell_warp_rows_value_x<T, UNROLL, HACK, XPlain<T> >(a, warpRow, zval, xg);
}

template <typename T, int UNROLL, int HACK>
__device__ __forceinline__ void hell_warp_rows(const HellArgs<T>& a, unsigned warpRow)
{
	T unused;
	hell_warp_rows_value<T, UNROLL, HACK>(a, warpRow, unused);
}

#endif
