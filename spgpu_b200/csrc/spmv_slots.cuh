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
 *            are switched off by predicated loads (select, not branch), so the
 *            padding is never touched.
 *   phase 2  only when some row of the warp is longer than `cut` (spike rows):
 *            those rows are finished one at a time by ALL 32 lanes striding
 *            over the remaining slots, followed by a shuffle reduction -- a
 *            4096-slot row costs 128 warp rounds instead of 4096.
 * `cut` = min(longest row of the warp, longCut); longCut is chosen by the host
 * from avgNnzPerRow, so regular matrices (stencils) never enter phase 2.
 */
#ifndef SPGPU_SPMV_SLOTS_CUH_
#define SPGPU_SPMV_SLOTS_CUH_

#include "numeric.cuh"

template <typename T, int UNROLL>
__device__ __forceinline__ T warp_rows_dot(
	const T* __restrict__ vals, const int* __restrict__ idxs,   /* already at this lane's slot 0 */
	long long valStride, long long idxStride,
	int rowLen,              /* slots of this lane's row (0 for lanes past the end) */
	int longCut,
	const T* __restrict__ x, int baseIndex)
{
	const int lane = threadIdx.x & 31;
	T acc = Num<T>::zero();

	const int longest = __reduce_max_sync(SPGPU_FULL_MASK, rowLen);
	const int cut = min(longest, longCut);
	const int mine = min(rowLen, cut);

	/* ---- phase 1: one row per lane ---- */
	for (int k0 = 0; k0 < cut; k0 += UNROLL) {
		int col[UNROLL];
		T a[UNROLL];
		T xv[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			const int k = k0 + u;
			const bool on = k < mine;
			col[u] = on ? ld_stream(idxs + (long long)k * idxStride) : baseIndex;
			a[u] = on ? ld_stream(vals + (long long)k * valStride) : Num<T>::zero();
		}
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			const bool on = (k0 + u) < mine;
			xv[u] = on ? ld_keep(x + (col[u] - baseIndex)) : Num<T>::zero();
		}
#pragma unroll
		for (int u = 0; u < UNROLL; ++u)
			acc = Num<T>::fma(a[u], xv[u], acc);
	}

	/* ---- phase 2: spike rows, all lanes on one row ---- */
	unsigned todo = __ballot_sync(SPGPU_FULL_MASK, rowLen > cut);
	while (todo) {
		const int r = __ffs(todo) - 1;
		todo &= todo - 1;
		const int len = __shfl_sync(SPGPU_FULL_MASK, rowLen, r);
		/* pointers of lane r's row: same base shifted by (r - lane) elements */
		const T* rv = vals + (r - lane);
		const int* ri = idxs + (r - lane);
		T part = Num<T>::zero();
		for (int k0 = cut + lane; k0 < len; k0 += 32 * 4) {
			int col[4];
			T a[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int k = k0 + 32 * u;
				const bool on = k < len;
				col[u] = on ? ld_stream(ri + (long long)k * idxStride) : baseIndex;
				a[u] = on ? ld_stream(rv + (long long)k * valStride) : Num<T>::zero();
			}
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const bool on = (k0 + 32 * u) < len;
				T xv = on ? ld_keep(x + (col[u] - baseIndex)) : Num<T>::zero();
				part = Num<T>::fma(a[u], xv, part);
			}
		}
		part = warp_sum<T>(part);
		if (lane == r)
			acc = Num<T>::add(acc, part);
	}
	return acc;
}

#endif
