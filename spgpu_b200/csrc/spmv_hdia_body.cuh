/*
 * The per-warp body of the direct HDIA SpMV, shared by the plain kernel (spmv_hdia.cu) and the
 * kernel fused with the NVLink halo exchange (ext.cu).  `warpRow` = first row of the 32 rows this
 * warp owns; returns without doing anything when the warp lies past the last row.
 *
 * HACK > 0: hackSize known at compile time -> cell addresses are base + immediate.  All index
 * arithmetic is 32-bit: the in-range test 0 <= row+off < cols is ONE unsigned compare; lanes
 * past the last row get cols = 0 and offsets past a hack's last diagonal get INT_MIN, so both
 * fail that same compare without extra predicates.  The matrix cells of a round are loaded
 * without waiting for the offsets (they all exist in the slab); only the x gather and the FMA
 * depend on the in-range test.
 */
#ifndef SPGPU_SPMV_HDIA_BODY_CUH_
#define SPGPU_SPMV_HDIA_BODY_CUH_

#include <climits>
#include "numeric.cuh"

template <typename T>
struct HdiaArgs {
	T* z;
	const T* y;
	T alpha;
	const T* dM;
	const int* offsets;
	int hackSize;
	const int* hackOffsets;
	int rows;
	int cols;
	const T* x;
	T beta;
	/* > 0: a warp looks this many hacks ahead -- it asks L2 for that hack's hackOffsets entry and, at the end of its
	 * life, for that hack's slice of offsets[] (whose position it read at its start, off the critical path).  A warp
	 * lives hackOffsets -> offsets -> rounds of (cells, x): the first two are tiny reads nobody has touched before,
	 * HBM latency for 1 % of the bytes; prefetched a couple of waves ahead they are L2 hits. */
	int prefetchHacks;
};

/* one round of the direct kernel: UNROLL diagonals starting at diagonal u0 of the 32 whose offsets
 * the warp holds in mineOff; GUARD = the round may run past the hack's last diagonal (n) */
template <typename T, int UNROLL, bool GUARD, bool PREDICATED, class XG>
__device__ __forceinline__ T hdia_round(T acc, const T* __restrict__ cp, long long hackSize, int mineOff,
	int u0, int n, unsigned i, unsigned colsEff, const XG xg)
{
	T a[UNROLL];
	T xv[UNROLL];
	bool on[UNROLL];
	if (!PREDICATED) {
#pragma unroll
		for (int u = 0; u < UNROLL; ++u) {
			a[u] = Num<T>::zero();
			if (!GUARD || u0 + u < n)                 /* warp-uniform: cell exists */
				a[u] = ld_stream(cp + u * hackSize);
		}
	}
#pragma unroll
	for (int u = 0; u < UNROLL; ++u) {
		const int off = __shfl_sync(SPGPU_FULL_MASK, mineOff, u0 + u);
		const int c = (int)i + off;
		on[u] = (unsigned)c < colsEff && (!GUARD || u0 + u < n);   /* u0+u may pass lane 31 when UNROLL does not divide 32 */
		xv[u] = Num<T>::zero();
		if (PREDICATED)
			a[u] = Num<T>::zero();
		if (on[u]) {
			xv[u] = xg.ld(c);
			if (PREDICATED)                           /* cells outside the matrix are not read */
				a[u] = ld_stream(cp + u * hackSize);
		}
	}
#pragma unroll
	for (int u = 0; u < UNROLL; ++u)
		acc = PREDICATED ? Num<T>::fma(a[u], xv[u], acc)
		                 : (on[u] ? Num<T>::fma(a[u], xv[u], acc) : acc);
	return acc;
}

/* returns the value stored for this lane's row in `zval` (zero for lanes without a row) */
template <typename T, int UNROLL, int HACK, bool PREDICATED, class XG>
__device__ __forceinline__ void hdia_warp_rows_value_x(const HdiaArgs<T>& a, unsigned warpRow, T& zval, const XG xg)
{
	zval = Num<T>::zero();
	const int hackSize = HACK > 0 ? HACK : a.hackSize;
	const unsigned lane = threadIdx.x & 31;
	if (warpRow >= (unsigned)a.rows)
		return;
	const unsigned i = warpRow + lane;
	const bool live = i < (unsigned)a.rows;
	const unsigned colsEff = live ? (unsigned)a.cols : 0u;
	const bool useBeta = Num<T>::nonzero(a.beta);
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = a.y[i];

	const unsigned hack = warpRow / (unsigned)hackSize;
	int aheadFirst = -1;
	if (a.prefetchHacks > 0) {
		const unsigned lastHack = ((unsigned)a.rows - 1u) / (unsigned)hackSize;
		const unsigned ahead = hack + (unsigned)a.prefetchHacks;
		if (ahead <= lastHack) {
			aheadFirst = __ldg(a.hackOffsets + ahead);          /* used only at the very end */
			if (lane == 0 && ahead + (unsigned)a.prefetchHacks <= lastHack)
				prefetch_l2(a.hackOffsets + ahead + a.prefetchHacks);
		}
	}
	const int first = __ldg(a.hackOffsets + hack);
	const int diags = __ldg(a.hackOffsets + hack + 1) - first;
	const T* cell = a.dM + (long long)first * hackSize + (warpRow % (unsigned)hackSize) + lane;
	const int* offs = a.offsets + first;
	T acc = Num<T>::zero();

	for (int j0 = 0; j0 < diags; j0 += 32) {
		const int mineOff = (j0 + (int)lane < diags) ? ld_stream(offs + j0 + lane) : INT_MIN;
		const int n = min(32, diags - j0);
		int u0 = 0;
		/* full rounds need no per-diagonal guard; the last, partial round does */
		for (; u0 + UNROLL <= n; u0 += UNROLL)
			acc = hdia_round<T, UNROLL, false, PREDICATED, XG>(acc, cell + (long long)(j0 + u0) * hackSize, hackSize, mineOff, u0, n, i, colsEff, xg);
		if (u0 < n)
			acc = hdia_round<T, UNROLL, true, PREDICATED, XG>(acc, cell + (long long)(j0 + u0) * hackSize, hackSize, mineOff, u0, n, i, colsEff, xg);
	}

	if (live) {
		zval = spmv_epilogue<T>(acc, a.alpha, a.beta, useBeta, yv);
		a.z[i] = zval;
	}
	if (aheadFirst >= 0 && lane < 2)
		prefetch_l2(a.offsets + aheadFirst + 32 * lane);        /* up to 64 diagonals of the hack ahead */
}

template <typename T, int UNROLL, int HACK, bool PREDICATED>
__device__ __forceinline__ void hdia_warp_rows_value(const HdiaArgs<T>& a, unsigned warpRow, T& zval)
{
	const XPlain<T> xg = { a.x };
	hdia_warp_rows_value_x<T, UNROLL, HACK, PREDICATED, XPlain<T> >(a, warpRow, zval, xg);
}

#endif
