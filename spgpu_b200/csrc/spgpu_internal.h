/*
 * Private side of a spGPU handle (shared by the C host layer and the .cu files).
 *
 * The reference handle is a plain public struct (reference core.h:60-82) and its
 * reductions keep their partial sums in file-static __device__ arrays (e.g.
 * reference kernels/ddot.cu:35), so two handles on one device race.  Here each
 * handle owns its scratch; the public struct is the first member, so a
 * spgpuHandle_t can be cast to SpgpuHandlePriv* inside the library.
 */
#ifndef SPGPU_INTERNAL_H_
#define SPGPU_INTERNAL_H_

#include "spgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SPGPU_PRIV_MAGIC 0x53504755u /* "SPGU" */

/* Partial results of one reduction launch: up to SPGPU_RED_MAX_BLOCKS blocks,
 * each writing up to 2 doubles (complex), plus a ticket counter. */
#define SPGPU_RED_MAX_BLOCKS 2048
#define SPGPU_RED_SLOT_BYTES 16

/* dTicket: a block of device counters, all zero between calls (every user resets what it used):
 *   [0] reductions, [4] halo push, [8] halo exchange / fused push, [12..13] fused-halo done tickets,
 *   [16..19] HELL split-mode header (items queued, items taken, units queued, tail warps done) */
#define SPGPU_TICKET_WORDS 32
#define SPGPU_TICKET_SPLIT 16

/* Tunables that steer kernel selection; settable with spgpuSetTuning (ext). */
typedef struct SpgpuTuning {
	int hellVariant;     /* 0 auto, 1 direct loads predicated on rS, 2 direct loads with unpredicated slab reads, 3 bulk-async (TMA) pipeline */
	int hellBlock;       /* occupancy knob: 0 per-type default, <=64 force 32 warps/SM, 192 force 40, >=256 force 48 */
	int hellLongFactor;  /* a row deeper than factor x avgNnzPerRow (>= 32) slots counts as a spike (default 2: bench/longfactor_probe.py) */
	int hellSplit;       /* long-hack split mode: 0 auto (on when rIdx is given), 1 on, -1 off, >1 on with that queue capacity */
	int hdiaVariant;     /* 0/1 direct (unpredicated cell loads), 2 x windows staged in shared memory, 3 direct with predicated cell loads, 4 bulk-async (TMA) pipeline, 5 persistent with metadata prefetch, 6/7 per-warp slab by bulk copy with 16/32 x gathers in flight (hackSize 32) */
	int hdiaBlock;       /* occupancy knob: 8 -> rounds of 8 diagonals instead of 9, 64 -> 64-thread CTAs, 160 -> 24 warps/SM with twice the unroll, 176 -> 36, 192 -> 40, 224 -> 48 warps with UNROLL 4, >=256 -> 48 (default 32); variants 6/7: 1..32 = diagonals per warp slice */
	int ellRows;         /* ELL, short regular rows: 0/1 = exact-slot-count kernel (default), 2 = the same with 2 rows per lane, -1 = general kernel */
	int redBlocksPerSm;  /* CTAs per SM for the reductions                      */
	int vecBlocksPerSm;  /* CTAs per SM for grid-stride vector kernels          */
	int spinTimeoutMs;   /* how long a device-side wait on a peer's flag may last before it gives up and sets the
	                      * handle's sticky device status (default 20000; 0 = wait for ever) */
	int haloTrace;       /* 1: the fused SpMV + halo kernels record per-exchange timestamps (spgpuHaloTraceRead) */
	int l2Fetch;         /* > 0: cudaLimitMaxL2FetchGranularity of the handle's device is set to this many bytes
	                      * (32 / 64 / 128) when the key is set -- a DEVICE-wide limit, an experiment knob */
	int redInflight;     /* reductions: 16-byte packs a thread keeps in flight per input (2, 4 or 8; 0 = default 4) */
	int ellShortMinB;    /* (reserved) */
	int hdiaPrefetch;    /* HDIA: the same look-ahead for hackOffsets and the hack's slice of offsets[] (0 = default 2 waves, < 0 = off) */
	int hellPrefetch;    /* HELL: waves of resident CTAs ahead of which a warp prefetches its hackOffsets entry into L2 (0 = default 2, < 0 = off) */
	int pdl;             /* > 0: the kernels that begin with grid_dependency_wait() are launched with programmatic stream
	                      * serialization, so the next kernel's CTAs take the slots the previous kernel's tail leaves (launch.cuh) */
} SpgpuTuning;

typedef struct SpgpuHandlePriv {
	SpgpuHandleStruct pub;         /* MUST stay first                            */
	unsigned magic;
	void* dPartials;               /* device: SPGPU_RED_MAX_BLOCKS slots          */
	unsigned* dTicket;             /* device: SPGPU_TICKET_WORDS counters (kept 0)  */
	void* hResult;                 /* pinned, mapped host: final reduction value  */
	void* dResult;                 /* device alias of hResult                     */
	int l2Bytes;
	int smemPerBlockOptin;
	unsigned long long launches;   /* kernels launched through this handle        */
	int debug;                     /* SPGPU_DEBUG set: check for CUDA errors after every launch */
	void* dBig;                    /* device: grow-only scratch (per-CTA partials of fused kernels) */
	size_t bigBytes;
	unsigned* dHaloSeq;            /* device counters registered with spgpuSetSeqCounters (ext): completed halo */
	unsigned* dArSeq;              /* exchanges / all-reduces; used by calls that pass seq == 0                 */
	unsigned* hStatus;             /* sticky device status word (SPGPU_DEVSTATUS_*), mapped pinned: byte 32 of hResult */
	unsigned* dStatus;             /* its device alias                                                            */
	cudaEvent_t switchEvent;       /* orders work across a spgpuSetStream (the handle's scratch is reused in stream order) */
	unsigned long long* dTrace;    /* device: SPGPU_TRACE_SLOTS x 8 words, allocated when haloTrace is set        */
	unsigned* dTicketsMany;        /* device: one ticket word per vector of a multi-vector reduction (kept 0)     */
	size_t ticketsMany;
	SpgpuTuning tune;
} SpgpuHandlePriv;

/* grow-only device scratch owned by the handle; NULL on failure (core.c) */
void* spgpuScratch(spgpuHandle_t handle, size_t bytes);
/* `count` zero-initialised ticket words (their users leave them zero); NULL on failure (core.c) */
unsigned* spgpuTickets(spgpuHandle_t handle, size_t count);

static inline SpgpuHandlePriv* spgpuPriv(spgpuHandle_t h)
{
	return (SpgpuHandlePriv*)h;
}

#ifdef __cplusplus
}
#endif

#endif
