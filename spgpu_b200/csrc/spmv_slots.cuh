/*
 * Row dot-product over "slotted" storage, shared by the ELL and HELL kernels.
 *
 * Both formats store slot k of a row at  valBase + k*valStride  (values) and
 * idxBase + k*idxStride (column indices):
 *     ELL   base = row,                        stride = pitch
 *     HELL  base = hackOffsets[h] + row%hack,  stride = hackSize
 * so lanes of a warp (= consecutive rows) read consecutive elements: every
 * warp-level load is one fully used 128/256-byte run.
 *
 * One warp owns 32 consecutive rows, one row per lane.  Work is done in two
 * warp-uniform phases:
 *   phase 1  slots [0, cut): each lane walks its own row, UNROLL slots per
 *            round; all index and value loads of a round are issued before the
 *            dependent x gathers, then the FMAs.  Lanes whose row is shorter
 *            are switched off by predication (select, not branch).
 *   phase 2  only when a few rows of the warp are much longer than the rest
 *            (spike rows): those rows are finished one at a time by ALL 32
 *            lanes striding over the remaining slots, followed by a shuffle
 *            reduction -- a 4096-slot row costs 128 warp rounds instead of 4096.
 *            `cut` adapts per warp (see below), so hacks whose rows are all
 *            long (length-sorted OHELL) stay in phase 1.
 *
 * STRIDE > 0 makes the slot stride a compile-time constant (HELL with the usual
 * hackSize 32/64): every load of a round is then base + immediate, with no
 * address arithmetic at all.
 *
 * `allocated` (HELL only) is the number of slots that exist in memory for every
 * row of the warp (the hack's slab height).  Slots below it can be loaded
 * WITHOUT waiting for rS -- their content may be padding garbage, so the lane
 * predicate is applied to the x gather and to the FMA instead of to the load.
 * This removes one dependent DRAM round trip (rS -> matrix) from every warp;
 * it costs no extra DRAM traffic because a slot row is fetched as whole lines
 * as soon as one row of the hack uses it.
 */
#ifndef SPGPU_SPMV_SLOTS_CUH_
#define SPGPU_SPMV_SLOTS_CUH_

#include "numeric.cuh"

/* rounds of 32 slots a warp keeps in flight while it finishes a spike row (phase 2).  8 and 16 were tried to shorten the
 * walk of a 4096-slot row: the extra registers spill (double: 12 -> 144 bytes) and EVERY matrix pays -- 14 irregular matrices
 * 3 - 15 % slower at 8, 30 - 50 % at 16, and the 512^3 Laplacian, which never enters phase 2, 2.144 -> 2.215 ms. */
#ifndef SPGPU_PHASE2_UNROLL
#define SPGPU_PHASE2_UNROLL 4
#endif

template <typename T, int UNROLL, int STRIDE, class XG>
__device__ __forceinline__ T warp_rows_dot_x(
	const T* __restrict__ vals, const int* __restrict__ idxs,   /* already at this lane's slot 0 */
	int valStrideRt, int idxStrideRt,
	int rowLen,              /* slots of this lane's row (0 for lanes past the end) */
	int longCut,             /* depth beyond which a row counts as a spike (host: 4 x avgNnzPerRow, >= 32) */
	int allocated,           /* slots guaranteed to exist for the whole warp, 0 = unknown */
	const XG xg,             /* how x[c] is fetched (numeric.cuh: XPlain, XZones) */
	int baseIndex)
{
	const int lane = threadIdx.x & 31;
	const long long valStride = STRIDE > 0 ? STRIDE : valStrideRt;
	const long long idxStride = STRIDE > 0 ? STRIDE : idxStrideRt;
	T acc = Num<T>::zero();
	int cut;

	if (allocated > 0 && allocated <= UNROLL) {
		/* ---- phase 1, short regular rows: unpredicated matrix loads ---- */
		int col[UNROLL];
		T a[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			if (u < allocated) {               /* warp-uniform */
				col[u] = ld_stream(idxs + u * idxStride);
				a[u] = ld_stream(vals + u * valStride);
			}
		}
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			if (u < allocated) {
				const bool on = u < rowLen;
				T xv = Num<T>::zero();
				if (on)
					xv = xg.ld(col[u] - baseIndex);
				acc = on ? Num<T>::fma(a[u], xv, acc) : acc;
			}
		}
		cut = allocated;
	} else {
		/* ---- phase 1, general: one row per lane, predicated loads ---- */
		const int longest = __reduce_max_sync(SPGPU_FULL_MASK, rowLen);
		cut = min(longest, longCut);
		if (longest > longCut) {
			/* Some row is much longer than the matrix average (longCut = 4 x avgNnzPerRow,
			 * at least 32).  If only a FEW rows are (spikes in an otherwise short hack), stop
			 * the row-per-lane walk at longCut and finish those rows cooperatively (phase 2).
			 * If MANY rows of the warp are that long (homogeneous long hacks: length-sorted
			 * OHELL, dense blocks), the row-per-lane walk is the efficient one -- it stays
			 * coalesced -- so it continues while at least SPIKE_ROWS rows are still active:
			 * cut = the smallest depth with fewer than SPIKE_ROWS longer rows (binary search
			 * on ballots). */
			constexpr int SPIKE_ROWS = 8;            /* 4 and 16 measure the same (bench/longfactor_probe.py, 14 matrices) */
			int lo = longCut, hi = longest;
			while (lo < hi) {
				const int mid = (lo + hi) >> 1;
				if (__popc(__ballot_sync(SPGPU_FULL_MASK, rowLen > mid)) < SPIKE_ROWS)
					hi = mid;
				else
					lo = mid + 1;
			}
			cut = lo;
		}
		const int mine = min(rowLen, cut);
		const T* vp = vals;
		const int* ip = idxs;
		for (int k0 = 0; k0 < cut; k0 += UNROLL) {
			int col[UNROLL];
			T a[UNROLL];
			T xv[UNROLL];
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const bool on = (k0 + u) < mine;
				col[u] = baseIndex;
				a[u] = Num<T>::zero();
				if (on) {                              /* partly used lines: 64-byte L2 fetch granularity */
					col[u] = ld_stream64(ip + u * idxStride);
					a[u] = ld_stream64(vp + u * valStride);
				}
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const bool on = (k0 + u) < mine;
				xv[u] = Num<T>::zero();
				if (on)
					xv[u] = xg.ld(col[u] - baseIndex);
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
				acc = Num<T>::fma(a[u], xv[u], acc);
			vp += UNROLL * valStride;
			ip += UNROLL * idxStride;
		}
	}

	/* ---- phase 2: spike rows, all lanes on one row ---- */
	/* One warp walks the whole row, and every round is two memory latencies (slot, then x): P2U rounds in flight. */
	constexpr int P2U = SPGPU_PHASE2_UNROLL;
	unsigned todo = __ballot_sync(SPGPU_FULL_MASK, rowLen > cut);
	while (todo) {
		const int r = __ffs(todo) - 1;
		todo &= todo - 1;
		const int len = __shfl_sync(SPGPU_FULL_MASK, rowLen, r);
		/* pointers of lane r's row: same base shifted by (r - lane) elements */
		const T* rv = vals + (r - lane);
		const int* ri = idxs + (r - lane);
		T part = Num<T>::zero();
		for (int k0 = cut + lane; k0 < len; k0 += 32 * P2U) {
			int col[P2U];
			T a[P2U];
#pragma unroll
			for (int u = 0; u < P2U; ++u) {
				const int k = k0 + 32 * u;
				const bool on = k < len;
				col[u] = baseIndex;
				a[u] = Num<T>::zero();
				if (on) {                              /* one row's slots are a whole stride apart: one element per line */
					col[u] = ld_stream64(ri + k * idxStride);
					a[u] = ld_stream64(rv + k * valStride);
				}
			}
#pragma unroll
			for (int u = 0; u < P2U; ++u) {
				const bool on = (k0 + 32 * u) < len;
				T xv = Num<T>::zero();
				if (on)
					xv = xg.ld(col[u] - baseIndex);
				part = Num<T>::fma(a[u], xv, part);
			}
		}
		part = warp_sum<T>(part);
		if (lane == r)
			acc = Num<T>::add(acc, part);
	}
	return acc;
}

template <typename T, int UNROLL, int STRIDE>
__device__ __forceinline__ T warp_rows_dot(
	const T* __restrict__ vals, const int* __restrict__ idxs, int valStrideRt, int idxStrideRt,
	int rowLen, int longCut, int allocated, const T* __restrict__ x, int baseIndex)
{
	const XPlain<T> xg = { x };
	return warp_rows_dot_x<T, UNROLL, STRIDE, XPlain<T> >(vals, idxs, valStrideRt, idxStrideRt, rowLen, longCut,
		allocated, xg, baseIndex);
}

#endif
