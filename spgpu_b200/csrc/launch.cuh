/* Small host-side helpers shared by the .cu launchers. */
#ifndef SPGPU_LAUNCH_CUH_
#define SPGPU_LAUNCH_CUH_

#include <cstdio>
#include <cstring>
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
	static const SpgpuTuning fallback = { 0, 0, 2, 0, 0, 0, 0, 8, 8, 20000, 0, 0, 0, 0, 0, 0, 1 };
	SpgpuHandlePriv* h = spgpuPriv(handle);
	return h->magic == SPGPU_PRIV_MAGIC ? &h->tune : &fallback;
}

/*
 * Programmatic dependent launch.  A kernel that begins with grid_dependency_wait() (griddepcontrol.wait: returns once
 * every kernel before it in the stream has completed and its writes are visible) and then calls
 * grid_launch_dependents() may be launched with cudaLaunchAttributeProgrammaticStreamSerialization: when the LAST wave of
 * the kernel before it has started, its CTAs are already placed in the slots that wave leaves free and wait there, so
 * the launch latency and the ramp-up of one kernel overlap the tail of the other -- microseconds that matter when a
 * partitioned SpMV takes 0.28 ms and a CG iteration is five such kernels.  Without the attribute both instructions are
 * no-ops.  ONLY kernels that execute grid_dependency_wait() in every CTA before touching memory may go through here.
 * No reference counterpart (the reference launches with <<< >>> on handle->currentStream, e.g. reference
 * kernels/hell_spmv_base_template.cuh:150-190).
 */
#ifdef __CUDACC__
__device__ __forceinline__ void grid_dependency_wait()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void grid_launch_dependents()
{
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline void spgpu_launch_dep(spgpuHandle_t handle, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args)
{
	cudaLaunchConfig_t cfg;
	cudaLaunchAttribute at[1];
	memset(&cfg, 0, sizeof(cfg));
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.stream = handle->currentStream;
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = spgpu_tuning(handle)->pdl > 0 ? 1u : 0u;
	cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

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

/* rows deeper than this many slots count as spikes (spmv_slots.cuh decides per warp what to do with them).  The factor:
 * on Pareto row lengths (mean 8 .. 64, caps 64 .. 4096, float and double, bench/longfactor_probe.py,
 * profiles/r2_longfactor.jsonl) 2 x the average beats 3, 4 and 6 by 1 - 14 %, and 1 x loses 11 % once the average
 * reaches 32; uniform lengths do not care. */
static inline int spgpu_long_cut(const SpgpuTuning* t, int avgNnzPerRow)
{
	long long cut = (long long)(t->hellLongFactor > 0 ? t->hellLongFactor : 2) * (avgNnzPerRow > 0 ? avgNnzPerRow : 1);
	if (cut < 32) cut = 32;
	if (cut > (1 << 30)) cut = 1 << 30;
	return (int)cut;
}

#endif
