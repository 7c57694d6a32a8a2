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
	static const SpgpuTuning fallback = { 0, 0, 8, 0, 0, 128, 1, 4, 8 };
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

/* phase 1 of spmv_slots.cuh continues while at least this many rows of the warp are active */
static inline int spgpu_long_cut(const SpgpuTuning* t, int avgNnzPerRow)
{
	(void)avgNnzPerRow;
	int k = t->hellLongFactor > 0 ? t->hellLongFactor : 8;
	if (k > 32) k = 32;
	return k;
}

#endif
