/*
 * Device-side building blocks of the multi-GPU layer (no reference counterpart: the
 * reference is one handle per device with no communication, reference core.h:88-93):
 * system-scope acquire / release on flag words that live in a peer GPU's memory, a
 * bounded spin that reports instead of hanging, and the one-value sum all-reduce over
 * NVLink peer memory that the last CTA of a reduction can run in place.
 */
#ifndef SPGPU_PEER_SYNC_CUH_
#define SPGPU_PEER_SYNC_CUH_

#include <cstdio>
#include "reduce_common.cuh"
#include "spgpu_ext.h"

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

/* how long a spin may last and where a failure is recorded (the handle's sticky status word
 * in mapped pinned memory: the host reads it with spgpuGetDeviceStatus) */
struct SpinCtl {
	unsigned long long timeoutNs;       /* 0 = wait for ever */
	unsigned* status;                   /* may be NULL */
};

/*
 * Spin until *flag >= value (sequence numbers, wrap-safe).  Returns false -- after setting the
 * handle's sticky status to SPGPU_DEVSTATUS_TIMEOUT -- when the wait exceeds ctl.timeoutNs, or
 * at once when an earlier wait of this handle has already failed (so that one lost peer costs
 * one timeout, not one per kernel).  The caller then finishes with whatever data is there; the
 * host sees the status word, never a hung GPU.
 */
__device__ __forceinline__ bool spin_until(const unsigned* flag, unsigned value, const SpinCtl& ctl)
{
	unsigned v = ld_acquire_sys(flag);
	if ((int)(v - value) >= 0)
		return true;
	const unsigned long long t0 = global_timer_ns();
	for (unsigned it = 1;; ++it) {
		__nanosleep(it < 64 ? 20 : 200);
		v = ld_acquire_sys(flag);
		if ((int)(v - value) >= 0)
			return true;
		if ((it & 255u) == 0u) {
			if (ctl.status && *reinterpret_cast<volatile unsigned*>(ctl.status) != 0u)
				return false;
			if (ctl.timeoutNs && global_timer_ns() - t0 > ctl.timeoutNs) {
				if (ctl.status)
					*reinterpret_cast<volatile unsigned*>(ctl.status) = SPGPU_DEVSTATUS_TIMEOUT;
				printf("spgpu: timed out waiting for flag value %u (saw %u)\n", value, v);
				return false;
			}
		}
	}
}

/* ---- one-value sum all-reduce over NVLink peer memory --------------------------------- */

#define SPGPU_MAX_RANKS 16
struct PeerTables { unsigned char* t[SPGPU_MAX_RANKS]; };
/* one slot of a rank's table: a value of any of the four types as two doubles + its sequence number */
struct alignas(16) ArSlot { double a, b; unsigned seq; unsigned pad[3]; };

struct ArArgs {
	int world;                 /* <= 1: no all-reduce */
	int myRank;
	PeerTables tables;         /* tables.t[r]: rank r's 2 * world slots (peer pointer for r != myRank) */
	unsigned seq;              /* sequence number of this all-reduce, or ...                         */
	unsigned* seqPtr;          /* ... (seq == 0) device counter of COMPLETED all-reduces (advanced here) */
	SpinCtl spin;
};

/*
 * Called by ONE warp (all 32 lanes).  Every rank stores (value, seq) into slot [parity][myRank]
 * of EVERY rank's table (remote 32-byte stores over NVLink, value first, then a release store of
 * seq), then polls its own table until all `world` slots of this parity carry seq, and adds the
 * values in rank order -- the same order on every rank, so all ranks get the same bits.  Two
 * parities because a fast rank can be at most one all-reduce ahead of the slowest.
 */
__device__ __forceinline__ Acc2 peer_allreduce_sum_warp(Acc2 mine, const ArArgs& ar)
{
	const unsigned seq = ar.seqPtr ? *reinterpret_cast<volatile unsigned*>(ar.seqPtr) + 1u : ar.seq;
	__syncwarp();
	const int r = threadIdx.x & 31;
	const unsigned parity = seq & 1u;
	Acc2 v = { 0.0, 0.0 };
	if (r < ar.world) {
		ArSlot* dst = reinterpret_cast<ArSlot*>(ar.tables.t[r]) + parity * ar.world + ar.myRank;
		*reinterpret_cast<volatile double*>(&dst->a) = mine.a;
		*reinterpret_cast<volatile double*>(&dst->b) = mine.b;
		__threadfence_system();
		st_release_sys(&dst->seq, seq);
		const ArSlot* src = reinterpret_cast<const ArSlot*>(ar.tables.t[ar.myRank]) + parity * ar.world + r;
		spin_until(&src->seq, seq, ar.spin);
		v.a = *reinterpret_cast<const volatile double*>(&src->a);
		v.b = *reinterpret_cast<const volatile double*>(&src->b);
	}
	Acc2 total = { 0.0, 0.0 };
	for (int k = 0; k < ar.world; ++k) {
		total.a += __shfl_sync(SPGPU_FULL_MASK, v.a, k);
		total.b += __shfl_sync(SPGPU_FULL_MASK, v.b, k);
	}
	if (r == 0 && ar.seqPtr)
		*ar.seqPtr = seq;
	return total;
}

/* partial sum a WARP of a fused SpMV + dot kernel leaves behind: one double for the real types, two for the complex ones */
template <typename T> struct DotPartial { typedef double type; };
template <> struct DotPartial<cuFloatComplex> { typedef Acc2 type; };
template <> struct DotPartial<cuDoubleComplex> { typedef Acc2 type; };
__device__ __forceinline__ Acc2 ld_partial(const double* p) { return { __ldcs(p), 0.0 }; }
__device__ __forceinline__ Acc2 ld_partial(const Acc2* p)
{
	const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
	return { v.x, v.y };
}
__device__ __forceinline__ void acc2_to_partial(Acc2 v, double& out) { out = v.a; }
__device__ __forceinline__ void acc2_to_partial(Acc2 v, Acc2& out) { out = v; }

/* value of type T <-> the two doubles of a slot / partial */
template <typename T> __device__ __forceinline__ Acc2 to_acc2(T v);
template <> __device__ __forceinline__ Acc2 to_acc2<float>(float v) { return { (double)v, 0.0 }; }
template <> __device__ __forceinline__ Acc2 to_acc2<double>(double v) { return { v, 0.0 }; }
template <> __device__ __forceinline__ Acc2 to_acc2<cuFloatComplex>(cuFloatComplex v) { return { (double)v.x, (double)v.y }; }
template <> __device__ __forceinline__ Acc2 to_acc2<cuDoubleComplex>(cuDoubleComplex v) { return { v.x, v.y }; }

template <typename T> __device__ __forceinline__ T from_acc2(Acc2 v);
template <> __device__ __forceinline__ float from_acc2<float>(Acc2 v) { return (float)v.a; }
template <> __device__ __forceinline__ double from_acc2<double>(Acc2 v) { return v.a; }
template <> __device__ __forceinline__ cuFloatComplex from_acc2<cuFloatComplex>(Acc2 v) { return make_cuFloatComplex((float)v.a, (float)v.b); }
template <> __device__ __forceinline__ cuDoubleComplex from_acc2<cuDoubleComplex>(Acc2 v) { return make_cuDoubleComplex(v.a, v.b); }

#endif
