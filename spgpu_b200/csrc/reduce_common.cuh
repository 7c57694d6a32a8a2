/*
 * Building blocks shared by the reductions (blas1_reduce.cu) and the fused Krylov
 * kernels (ext.cu): a 2-double accumulator, the warp-shuffle + shared-memory block
 * reduction, and the "last CTA folds the per-CTA partials in index order" finish that
 * makes every reduction deterministic for a given launch shape.
 */
#ifndef SPGPU_REDUCE_COMMON_CUH_
#define SPGPU_REDUCE_COMMON_CUH_

#include "numeric.cuh"

/* accumulator: up to two doubles (re, im) or (value, unused) */
struct alignas(16) Acc2 { double a, b; };

template <bool IS_MAX>
__device__ __forceinline__ Acc2 combine(Acc2 p, Acc2 q)
{
	if (IS_MAX)
		return { fmax(p.a, q.a), 0.0 };
	return { p.a + q.a, p.b + q.b };
}

template <bool IS_MAX>
__device__ __forceinline__ Acc2 block_reduce(Acc2 v, Acc2* smem)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		Acc2 o = { __shfl_xor_sync(SPGPU_FULL_MASK, v.a, m), __shfl_xor_sync(SPGPU_FULL_MASK, v.b, m) };
		v = combine<IS_MAX>(v, o);
	}
	if (lane == 0)
		smem[warp] = v;
	__syncthreads();
	if (warp == 0) {
		const int nwarps = blockDim.x >> 5;
		v = lane < nwarps ? smem[lane] : Acc2{ 0.0, 0.0 };
#pragma unroll
		for (int m = 16; m > 0; m >>= 1) {
			Acc2 o = { __shfl_xor_sync(SPGPU_FULL_MASK, v.a, m), __shfl_xor_sync(SPGPU_FULL_MASK, v.b, m) };
			v = combine<IS_MAX>(v, o);
		}
	}
	__syncthreads();
	return v;          /* valid in warp 0 */
}

/*
 * Publishes this CTA's partial and, in the last CTA to arrive, folds all partials.
 * Returns true in thread 0 of that last CTA with the total in `total`; the ticket is
 * reset for the next launch.  `smem` = 32 Acc2, `flag` = one shared bool.
 */
template <bool IS_MAX>
__device__ __forceinline__ bool reduce_finish(Acc2 mine, Acc2* partials, unsigned* ticket,
	Acc2* smem, bool* flag, Acc2& total)
{
	if (threadIdx.x == 0) {
		partials[blockIdx.x] = mine;
		__threadfence();
		*flag = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	}
	__syncthreads();
	if (!*flag)
		return false;
	__threadfence();
	Acc2 t = { 0.0, 0.0 };
	for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
		const double2 pv = __ldcg(reinterpret_cast<const double2*>(partials) + b);
		t = combine<IS_MAX>(t, Acc2{ pv.x, pv.y });
	}
	t = block_reduce<IS_MAX>(t, smem);
	if (threadIdx.x == 0) {
		*ticket = 0u;
		total = t;
		return true;
	}
	return false;
}

#endif
