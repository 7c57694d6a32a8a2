/* Small host-side helpers shared by the .cu launchers. */
#ifndef SPGPU_LAUNCH_CUH_
#define SPGPU_LAUNCH_CUH_

#include <cstdio>
#include <cuda_runtime.h>
#include "spgpu_internal.h"

/* Called after every kernel launch.  With SPGPU_DEBUG set in the environment (read once per
 * handle) the stream is synchronised and any CUDA error is reported on stderr -- the
 * reference's -DDEBUG build prints and calls exit(0) (reference kernels/cudadebug.h:12-25);
 * here the process keeps running and the error stays queryable with cudaGetLastError. */
static inline void spgpu_count_launch(spgpuHandle_t handle)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (h->magic != SPGPU_PRIV_MAGIC)
		return;
	++h->launches;
	if (h->debug) {
		cudaError_t e = cudaStreamSynchronize(handle->currentStream);
		if (e == cudaSuccess)
			e = cudaPeekAtLastError();
		if (e != cudaSuccess)
			fprintf(stderr, "spgpu: CUDA error after launch %llu: %s\n", h->launches, cudaGetErrorString(e));
	}
}

static inline const SpgpuTuning* spgpu_tuning(spgpuHandle_t handle)
{
	static const SpgpuTuning fallback = { 0, 0, 4, 0, 0, 0, 0, 8, 8, 20000, 0, 0, 0, 0, 0, 0 };
	SpgpuHandlePriv* h = spgpuPriv(handle);
	return h->magic == SPGPU_PRIV_MAGIC ? &h->tune : &fallback;
}

static inline unsigned spgpu_ceil_div(long long a, long long b)
{
	return (unsigned)((a + b - 1) / b);
}

/* how many hacks ahead a HELL warp prefetches its hackOffsets entry (spmv_hell_body.cuh): hellPrefetch tuning key,
 * in waves of resident CTAs; 0 = default (2 waves), < 0 = off */
static inline int spgpu_hell_prefetch(spgpuHandle_t handle, const SpgpuTuning* t)
{
	const int waves = t->hellPrefetch == 0 ? 2 : t->hellPrefetch;
	if (waves < 0)
		return 0;
	return waves * 10 * 4 * (handle->multiProcessorCount > 0 ? handle->multiProcessorCount : 148);
}

/* rows deeper than this many slots count as spikes (spmv_slots.cuh decides per warp what to do with them) */
static inline int spgpu_long_cut(const SpgpuTuning* t, int avgNnzPerRow)
{
	long long cut = (long long)(t->hellLongFactor > 0 ? t->hellLongFactor : 4) * (avgNnzPerRow > 0 ? avgNnzPerRow : 1);
	if (cut < 32) cut = 32;
	if (cut > (1 << 30)) cut = 1 << 30;
	return (int)cut;
}

#endif
