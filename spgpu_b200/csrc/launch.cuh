/* Small host-side helpers shared by the .cu launchers. */
#ifndef SPGPU_LAUNCH_CUH_
#define SPGPU_LAUNCH_CUH_

#include <cuda_runtime.h>
#include "spgpu_internal.h"

static inline void spgpu_count_launch(spgpuHandle_t handle)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (h->magic == SPGPU_PRIV_MAGIC)
		++h->launches;
}

static inline const SpgpuTuning* spgpu_tuning(spgpuHandle_t handle)
{
	static const SpgpuTuning fallback = { 0, 0, 4, 0, 0, 128, 1, 4, 8 };
	SpgpuHandlePriv* h = spgpuPriv(handle);
	return h->magic == SPGPU_PRIV_MAGIC ? &h->tune : &fallback;
}

static inline int spgpu_block(int requested)
{
	if (requested < 32) return 32;
	if (requested > 1024) return 1024;
	return requested & ~31;
}

static inline unsigned spgpu_ceil_div(long long a, long long b)
{
	return (unsigned)((a + b - 1) / b);
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
