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
};

/* returns the value stored for this lane's row in `zval` (zero for lanes without a row) */
template <typename T, int UNROLL, int HACK>
__device__ __forceinline__ void hell_warp_rows_value(const HellArgs<T>& a, unsigned warpRow, T& zval)
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
	const int slab = __ldg(a.hackOffsets + hack);
	/* slab height of this hack = slots that exist for all of its rows; the last
	 * hack has no terminator entry, so it takes the predicated path */
	int allocated = 0;
	if (a.speculate && hack < lastHack)
		allocated = (__ldg(a.hackOffsets + hack + 1) - slab) / hackSize;
	const int len = live ? ld_stream(a.rS + i) : 0;
	const bool useBeta = Num<T>::nonzero(a.beta);
	const unsigned out = (live && a.rIdx) ? (unsigned)__ldg(a.rIdx + i) : i;
	T yv = Num<T>::zero();
	if (useBeta && live)
		yv = a.y[out];

	const long long at = (long long)slab + (warpRow % (unsigned)hackSize) + lane;
	T acc = warp_rows_dot<T, UNROLL, HACK>(a.cM + at, a.rP + at, hackSize, hackSize, len, a.longCut,
		allocated, a.x, a.baseIndex);

	if (live) {
		zval = spmv_epilogue<T>(acc, a.alpha, a.beta, useBeta, yv);
		a.z[out] = zval;
	}
}

template <typename T, int UNROLL, int HACK>
__device__ __forceinline__ void hell_warp_rows(const HellArgs<T>& a, unsigned warpRow)
{
	T unused;
	hell_warp_rows_value<T, UNROLL, HACK>(a, warpRow, unused);
}

#endif
