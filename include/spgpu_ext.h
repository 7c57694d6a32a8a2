/*
 * spgpu_ext.h -- ADDITIVE entry points of the B200-native build.
 *
 * Nothing here exists in the reference library and nothing here changes a
 * reference symbol; a consumer that only includes spgpu.h never sees them.
 * They exist for (1) tuning / introspection used by bench.py and the tests,
 * (2) Krylov loops that must not block the host: reductions that leave their
 * result in device memory, and vector updates that take their scalars from
 * device memory, so a whole CG iteration is stream-ordered and CUDA-graph
 * capturable, (3) the row-partitioned multi-GPU layer: halo packing / pushing
 * over NVLink peer pointers and CUDA IPC helpers to obtain those pointers
 * across the one-process-per-GPU ranks.
 */
#ifndef SPGPU_EXT_H_
#define SPGPU_EXT_H_

#include "spgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- introspection / tuning ---------------------------------------------- */

/* Version string of this build. */
const char* spgpuB200Version(void);

/* Kernels launched through this handle since spgpuCreate. */
unsigned long long spgpuGetLaunchCount(spgpuHandle_t handle);

/*
 * Kernel-selection knobs (per handle).  Keys: hellVariant, hellBlock,
 * hellLongFactor, hellSplit, hdiaVariant, hdiaBlock, ellRows, ellShortMinB,
 * redBlocksPerSm, vecBlocksPerSm, redInflight, hellPrefetch, hdiaPrefetch,
 * spinTimeoutMs, haloTrace, l2Fetch, pdl (meanings: csrc/spgpu_internal.h).
 * Returns 0, or -1 for an unknown key.
 *
 * pdl (default 1; the environment variable SPGPU_PDL=0/1 overrides the default at spgpuCreate):
 * the SpMV, BLAS-1 and Krylov kernels are launched with programmatic stream serialization and
 * begin with griddepcontrol.wait, so the CTAs of a kernel are placed while the last wave of the
 * kernel before it in the stream drains and start the moment it has completed -- same results
 * (tests/test_pdl_gpu.py), launch latency and ramp-up hidden.  Calls stay ordered on
 * handle->currentStream exactly as the reference's <<< >>> launches are.
 */
int spgpuSetTuning(spgpuHandle_t handle, const char* key, int value);
int spgpuGetTuning(spgpuHandle_t handle, const char* key);

/*
 * Sticky device-side status of the handle.  The kernels that wait for a peer GPU (halo flags,
 * all-reduce slots, spgpuWaitFlag) never hang: a wait longer than the spinTimeoutMs tuning key
 * (default 20000 ms, 0 = for ever) gives up, stores SPGPU_DEVSTATUS_TIMEOUT in a word of mapped
 * pinned memory and finishes with the data it has; later waits of the handle then return at once.
 * Returns the current status (clearing it when `clear` is non-zero), or -1 for a foreign handle.
 * A non-zero status means results since the last clear must be discarded.
 */
#define SPGPU_DEVSTATUS_OK       0
#define SPGPU_DEVSTATUS_TIMEOUT  1
int spgpuGetDeviceStatus(spgpuHandle_t handle, int clear);

/*
 * Sizes the handle's grow-only device scratch (per-warp partials of the fused SpMV + dot
 * kernels: 8 or 16 bytes per 32 rows; the split-mode queue of sorted HELL matrices; multi-vector
 * reduction results).  Growing synchronises the stream and allocates, which a CUDA-graph capture
 * does not allow -- reserve before capturing (or run the same call once eagerly).  0, or -1.
 */
int spgpuReserveScratch(spgpuHandle_t handle, size_t bytes);

/* ---- non-blocking reductions: result left in DEVICE memory ---------------- */

/*
 * dRes[0] = sum a_i*b_i (unconjugated, like spgpu?dot), written on the handle's
 * stream; no host synchronisation.  dRes must not alias a or b.
 */
#define SPGPU_DECL_DOT_DEV(S, T, R)                                          \
	void spgpu##S##dotDev(spgpuHandle_t handle, int n, const __device T* a,  \
		const __device T* b, __device T* dRes);                              \
	void spgpu##S##nrm2sqDev(spgpuHandle_t handle, int n,                    \
		const __device T* x, __device R* dRes);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_DOT_DEV)

/* dRes[0] = sum x_i, left in device memory (deterministic: per-CTA partials folded in index order). */
#define SPGPU_DECL_SUM_DEV(S, T, R)                                          \
	void spgpu##S##sumDev(spgpuHandle_t handle, int n, const __device T* x, __device T* dRes);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_SUM_DEV)

/*
 * One-value sum all-reduce across the ranks of a partition over NVLink peer memory (latency-bound
 * payload: one round of remote 32-byte stores + local polling instead of an NCCL launch).
 * tables[r] = pointer (a PEER pointer for r != myRank) to rank r's zero-initialised table of
 * 2 * world slots of SPGPU_AR_SLOT_BYTES bytes; world <= 16; seq = 1, 2, 3, ... identical on every
 * rank (0: taken from the device counter registered with spgpuSetSeqCounters).  All ranks obtain the
 * same bits (values are added in rank order, in double).  The entry points below that take a
 * `const spgpuPeerAllreduce* ar` run this exchange in the LAST CTA of the kernel that produces the
 * scalar -- no separate launch in the dependent chain of an iteration; ar == NULL (or world <= 1)
 * leaves the local value.
 */
#define SPGPU_AR_SLOT_BYTES 32
typedef struct spgpuPeerAllreduce {
	int world;
	int myRank;
	void* const* tables;
	unsigned seq;
} spgpuPeerAllreduce;
#define SPGPU_DECL_ALLREDUCE(S, T, R)                                        \
	void spgpu##S##allreduceSumDev(spgpuHandle_t handle, __device T* dValue, \
		const spgpuPeerAllreduce* ar);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_ALLREDUCE)

/*
 * z = (*dBetaNum / *dBetaDen) * betaSign * y + (*dAlphaNum / *dAlphaDen) *
 * alphaSign * x with the four scalars read from DEVICE memory at kernel time
 * (pass NULL for a numerator/denominator pair to mean 1).  This is the shape of
 * every CG update: x += (rr/pAp) p ; r -= (rr/pAp) Ap ; p = r + (rr'/rr) p.
 */
#define SPGPU_DECL_AXPBY_DEV(S, T, R)                                        \
	void spgpu##S##axpbyDev(spgpuHandle_t handle, __device T* z, int n,      \
		const __device T* dBetaNum, const __device T* dBetaDen,              \
		double betaSign, const __device T* y,                                \
		const __device T* dAlphaNum, const __device T* dAlphaDen,            \
		double alphaSign, const __device T* x);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_AXPBY_DEV)

/* ---- fused Krylov kernels --------------------- */

/*
 * HELL SpMV z = A*x fused with dRes[0] = sum_i x[xOffset+i]*z[i] (the p.Ap of CG; unconjugated
 * for C/Z like spgpu?dot) in the SpMV epilogue.  xOffset = position of row 0's own entry inside x
 * (0 on one GPU).  Two launches: the SpMV leaves one partial per 128-row block in handle-owned
 * scratch, a small kernel folds them (deterministic).
 */
#define SPGPU_DECL_HELLSPMV_DOT(S, T, R)                                     \
	void spgpu##S##hellspmvDot(spgpuHandle_t handle, __device T* z,          \
		const __device T* cM, const __device int* rP, int hackSize,          \
		const __device int* hackOffsets, const __device int* rS, int rows,   \
		const __device T* x, int baseIndex, int xOffset, __device T* dRes);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_HELLSPMV_DOT)

/*
 * Fused CG update: x += a*p ; r -= a*Ap ; dRrNew[0] = r.r (all-reduced when ar is given), with
 * a = *dRr / *dPAp read from device memory.  One pass over four vectors instead of three kernels.
 */
#define SPGPU_DECL_CGUPDATE_DEV(S, T, R)                                     \
	void spgpu##S##cgUpdateDev(spgpuHandle_t handle, __device T* x, __device T* r, \
		const __device T* p, const __device T* ap, int n,                    \
		const __device T* dRr, const __device T* dPAp, __device T* dRrNew,   \
		const spgpuPeerAllreduce* ar);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_CGUPDATE_DEV)

/* ---- device-side format construction -------------------- */

/*
 * CSR (device, rowPtr with rows+1 entries, csrBase-based) -> HELL (device), bit-identical
 * to the host route cooToEll + ellToHell on the same entries.  Step 1 (blocking) fills
 * dRs and dHackOffsets and returns the number of elements to allocate for the HELL value
 * and index arrays; step 2 (asynchronous) places the entries (padding slots untouched).
 */
int spgpuCsrToHellLayoutDevice(spgpuHandle_t handle, int rows, const __device int* dRowPtr,
	int hackSize, __device int* dRs, __device int* dHackOffsets, long long* totalElements);
#define SPGPU_DECL_CSR2HELL(S, T, R)                                                 \
	void spgpu##S##csrToHellDevice(spgpuHandle_t handle, int rows,                    \
		const __device int* dRowPtr, const __device int* dCols, const __device T* dVals, \
		int csrBase, int hackSize, const __device int* dHackOffsets, int hellBase,    \
		__device T* dHellValues, __device int* dHellIndices);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_CSR2HELL)

/*
 * CSR -> OHELL on the device: rows stored by descending length in exactly the order the
 * reference's ellToOell produces (reference ell.c:84-202: its mergesort takes the right run
 * on ties, i.e. equal lengths end up by descending original row index).  Step 1 (blocking)
 * fills dRidx (dRidx[i] = CSR row stored as HELL row i -- the rIdx argument of
 * spgpu?hellspmv), dRs (lengths in the new order) and dHackOffsets; step 2 (asynchronous)
 * places the entries.
 */
int spgpuCsrToOhellLayoutDevice(spgpuHandle_t handle, int rows, const __device int* dRowPtr,
	int hackSize, __device int* dRidx, __device int* dRs, __device int* dHackOffsets,
	long long* totalElements);
#define SPGPU_DECL_CSR2OHELL(S, T, R)                                                \
	void spgpu##S##csrToOhellDevice(spgpuHandle_t handle, int rows,                   \
		const __device int* dRowPtr, const __device int* dCols, const __device T* dVals, \
		int csrBase, int hackSize, const __device int* dHackOffsets,                  \
		const __device int* dRidx, int hellBase, __device T* dHellValues,             \
		__device int* dHellIndices);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_CSR2OHELL)

/*
 * COO -> HDIA on the device: the twins of computeHdiaHackOffsetsFromCoo and cooToHdia
 * (reference hdia_conv.h:52-70, hdia.cpp:161-349) with the same argument order, device
 * pointers instead of host pointers, and a status return.  Same results bit for bit
 * (hackOffsets, offsets in ascending order per hack, cell placement); the COO entries may
 * be in any order; duplicate (row, col) entries are unspecified (the reference keeps the
 * last).  spgpuHdiaHackOffsetsFromCooDevice blocks (it returns the allocation height);
 * spgpu?cooToHdiaDevice is asynchronous and, like the reference, expects dHdiaValues
 * zero-filled by the caller.
 */
int spgpuHdiaHackOffsetsFromCooDevice(spgpuHandle_t handle, int* allocationHeight,
	__device int* dHackOffsets, int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const __device int* dCooRowIndices, const __device int* dCooColsIndices, int cooBaseIndex);
#define SPGPU_DECL_COO2HDIA(S, T, R)                                                 \
	int spgpu##S##cooToHdiaDevice(spgpuHandle_t handle, __device T* dHdiaValues,      \
		__device int* dHdiaOffsets, const __device int* dHackOffsets, int hackSize,   \
		int rowsCount, int columnsCount, int nonZerosCount,                           \
		const __device int* dCooRowIndices, const __device int* dCooColsIndices,      \
		const __device T* dCooValues, int cooBaseIndex);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_COO2HDIA)

/* ---- multi-GPU helpers ---------------------------------------------------- */

/* 64-byte CUDA IPC handle of a cudaMalloc'ed pointer / open it in a peer process. */
int spgpuIpcGetHandle(void* devPtr, void* handle64);
int spgpuIpcOpenHandle(const void* handle64, void** devPtr);
int spgpuIpcCloseHandle(void* devPtr);
/* raw cudaMalloc / cudaFree (IPC needs allocations that are not sub-allocated) */
int spgpuDeviceAlloc(void** devPtr, size_t bytes);
int spgpuDeviceFree(void* devPtr);

/*
 * Flag block of a rank: SPGPU_HALO_FLAG_WORDS zero-initialised unsigned words in device memory that the
 * neighbours can write (peer pointer / CUDA IPC).
 *   separate-kernel protocol (spgpuHaloExchange / spgpuHaloAck / spgpuHaloPush / spgpuWaitFlag):
 *     [0] ready-from-below  [1] ready-from-above  [2] ack-from-below  [3] ack-from-above
 *   fused protocol (spgpu?{hell,hdia}spmvHalo[Dot]):
 *     [4] ready-from-below  [5] ready-from-above  [6] ack-from-below  [7] ack-from-above
 * The two protocols number their exchanges independently; both write the halo zones inside xExt, so
 * the neighbouring ranks must be synchronised (any barrier) when a caller switches between them.
 */
#define SPGPU_HALO_FLAG_WORDS 16

/*
 * Halo push over NVLink: copy `bytes` bytes of this GPU into a PEER pointer (CUDA IPC or peer
 * access) with 128-bit stores, then release-store `flagValue` to the peer's flag word so the
 * consumer can wait on it (spgpuWaitFlag).  One kernel, on the handle's stream.  peerDst == NULL
 * with bytes == 0 only sets the flag.  spgpuDhaloPush: the same for n doubles.
 */
void spgpuHaloPush(spgpuHandle_t handle, void* peerDst, const void* src, size_t bytes,
	unsigned* peerFlag, unsigned flagValue);
void spgpuDhaloPush(spgpuHandle_t handle, double* peerDst, const double* src,
	int n, unsigned* peerFlag, unsigned flagValue);
/* Stream-ordered wait until *flag >= value (bounded spin, see spgpuGetDeviceStatus). */
void spgpuWaitFlag(spgpuHandle_t handle, const unsigned* flag, unsigned value);

/*
 * Halo exchange of a 1-D chain of ranks as ONE kernel in front of the SpMV: waits until
 * the neighbours have acknowledged the previous halo (ackLo/ackHi, local flags,
 * value seq-1), pushes srcLo[0..bytes) into the lower neighbour's upper halo zone and
 * srcHi[0..bytes) into the upper neighbour's lower halo zone (peer pointers), release-
 * stores seq into the neighbours' ready flags and returns when this rank's own
 * ready flags (myReadyLo/myReadyHi) have reached seq.  NULL pointers = no
 * neighbour on that side.  spgpuHaloAck (after the SpMV has consumed the halos)
 * stores seq into the neighbours' ack flags.  spgpuDhaloExchange: the same for n doubles.
 */
void spgpuHaloExchange(spgpuHandle_t handle, void* peerDstLo, const void* srcLo,
	void* peerDstHi, const void* srcHi, size_t bytes, const unsigned* ackLo,
	const unsigned* ackHi, unsigned* peerReadyLo, unsigned* peerReadyHi,
	const unsigned* myReadyLo, const unsigned* myReadyHi, unsigned seq);
void spgpuDhaloExchange(spgpuHandle_t handle, double* peerDstLo, const double* srcLo,
	double* peerDstHi, const double* srcHi, int n, const unsigned* ackLo,
	const unsigned* ackHi, unsigned* peerReadyLo, unsigned* peerReadyHi,
	const unsigned* myReadyLo, const unsigned* myReadyHi, unsigned seq);
void spgpuHaloAck(spgpuHandle_t handle, unsigned* peerAckLo, unsigned* peerAckHi, unsigned seq);

/*
 * Where the fused SpMV + halo kernels of a rank find its neighbours.  xExt = [lower zone (haloN) |
 * owned (rows) | upper zone (haloN)].  NULL peer flags = no neighbour on that side.
 */
typedef struct spgpuHaloLinks {
	void* peerLoUpperZone;      /* PEER pointer: the LOWER neighbour's upper zone (this rank's first haloN owned entries go there) */
	void* peerHiLowerZone;      /* PEER pointer: the UPPER neighbour's lower zone (this rank's last haloN owned entries go there)  */
	unsigned* myFlags;          /* this rank's flag block */
	unsigned* peerFlagsLo;      /* PEER pointers to the neighbours' flag blocks */
	unsigned* peerFlagsHi;
} spgpuHaloLinks;

/*
 * HELL / HDIA SpMV of one row block FUSED with its halo exchange -- one kernel per partitioned SpMV,
 * transfer and multiply overlapped inside the launch, for all four value types.  Arguments up to
 * baseIndex as spgpu?hellspmv (reference hell.h:45-169) without rIdx, columns indexing xExt; the HDIA
 * form takes the row block with its diagonal offsets addressing xExt, i.e. global offset + haloN,
 * and cols = rows + 2*haloN (mg.split_hdia / device_build.hdia_row_block / spgpuMg*Create).
 * seq = 1, 2, 3, ... (the same on every rank) numbers the exchanges; 0 = taken from the device counter
 * registered with spgpuSetSeqCounters.
 *
 * The first CTAs tell the neighbours that this rank's previous exchange is over (the kernel runs, so the rows that
 * read its zones have finished: the acknowledgement costs nothing), wait for the neighbours' acknowledgement of the
 * same, push this rank's two boundary runs into the neighbours' zones over NVLink and publish seq; most interior
 * row blocks are scheduled first, then the row blocks that read a zone (they wait on the local ready word), then a
 * few waves of interior blocks.  Every block runs the same row walk with coherent loads of x.  Neighbours may
 * drift most of a kernel apart without anybody waiting.
 * Requirements: every row references only columns within haloN of its own position (rows [0, haloN) may
 * read the lower zone, rows [rows - haloN, rows) the upper one -- the host layers check this when they
 * split a matrix); a rank with a neighbour owns at least haloN rows; consecutive fused calls of a rank
 * are ordered on one stream.
 *
 * ...HaloDot: alpha = 1, beta = 0, plus dRes[0] = sum_i xExt[haloN+i]*z[i] (this rank's share of p.Ap,
 * summed over all ranks when `ar` is given): per-warp partials in handle scratch + one fold kernel
 * whose last CTA runs the all-reduce.  links == NULL: the single-GPU fused SpMV + dot.
 */
#define SPGPU_DECL_SPMV_HALO(S, T, R)                                                                  \
	void spgpu##S##hellspmvHalo(spgpuHandle_t handle, __device T* z, const __device T* y, T alpha,     \
		const __device T* cM, const __device int* rP, int hackSize, const __device int* hackOffsets,   \
		const __device int* rS, int avgNnzPerRow, int rows, __device T* xExt, T beta, int baseIndex,   \
		int haloN, const spgpuHaloLinks* links, unsigned seq);                                         \
	void spgpu##S##hellspmvHaloDot(spgpuHandle_t handle, __device T* z, const __device T* cM,          \
		const __device int* rP, int hackSize, const __device int* hackOffsets, const __device int* rS, \
		int avgNnzPerRow, int rows, __device T* xExt, int baseIndex, int haloN,                        \
		const spgpuHaloLinks* links, unsigned seq, __device T* dRes, const spgpuPeerAllreduce* ar);    \
	void spgpu##S##hdiaspmvHalo(spgpuHandle_t handle, __device T* z, const __device T* y, T alpha,     \
		const __device T* dM, const __device int* offsets, int hackSize,                               \
		const __device int* hackOffsets, int rows, int cols, __device T* xExt, T beta, int haloN,      \
		const spgpuHaloLinks* links, unsigned seq);                                                    \
	void spgpu##S##hdiaspmvHaloDot(spgpuHandle_t handle, __device T* z, const __device T* dM,          \
		const __device int* offsets, int hackSize, const __device int* hackOffsets, int rows,          \
		int cols, __device T* xExt, int haloN, const spgpuHaloLinks* links, unsigned seq,              \
		__device T* dRes, const spgpuPeerAllreduce* ar);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_SPMV_HALO)

/*
 * The block schedule of the fused kernels, host arithmetic only (no device is touched; the CPU tests sweep it):
 * plan[0] = headBlocks (128-row blocks [0, headBlocks) hold a row of [0, haloN): they wait for the lower zone),
 * plan[1] = firstHiBlock (blocks from there on hold a row of [rows - haloN, rows): they wait for the upper zone;
 * 0xffffffff without an upper neighbour), plan[2..5] = early, nLo, nHi, hiStart: CTA c (counted behind the push CTAs)
 * multiplies block spgpuHaloBlockOf(plan, c) -- `early` interior blocks first, then the nLo lower-boundary blocks, the
 * nHi upper-boundary blocks from hiStart, then the remaining interior blocks.  Returns 0, or -1 for bad arguments.
 */
int spgpuHaloBlockPlan(int rows, int haloN, int hasLowerNeighbour, int hasUpperNeighbour, int multiProcessorCount,
	unsigned* plan /* 6 words */);
unsigned spgpuHaloBlockOf(const unsigned* plan, unsigned cta);

/*
 * Sequence numbers in device memory, so that a partitioned iteration can be captured ONCE in a CUDA
 * graph and replayed (a by-value seq would be frozen into the graph).  After spgpuSetSeqCounters the
 * fused halo kernels and the all-reduces, when called with seq == 0, take their sequence number from
 * the counters: *dHaloSeq / *dAllreduceSeq hold the number of COMPLETED exchanges / all-reduces (start
 * them at the host-side count, or 0).  An all-reduce advances its counter itself; the halo counter is
 * advanced by spgpuHaloSeqAdvance, a one-thread kernel the caller puts after each fused SpMV (the SpMV
 * kernel's CTAs must all read the same value however late they are scheduled).  Returns 0, or -1 for a
 * foreign handle.
 */
int spgpuSetSeqCounters(spgpuHandle_t handle, __device unsigned* dHaloSeq, __device unsigned* dAllreduceSeq);
void spgpuHaloSeqAdvance(spgpuHandle_t handle);

/*
 * Loads every kernel that can wait for a peer GPU (the fused SpMV + halo kernels, the exchange kernels, the Krylov
 * kernels whose last CTA all-reduces) into the CURRENT device now.  CUDA otherwise loads a kernel at its first launch,
 * and that load waits for the kernels already running on the device -- a process that drives several ranks from one
 * thread would deadlock if the first launch of one rank's kernel met another rank's kernel spinning on it.
 * spgpuMgCreate does this for its devices; one-process-per-GPU callers do not need it.  0, or -1.
 */
int spgpuPreloadHaloKernels(void);
int spgpuPreloadKrylovKernels(void);

/*
 * Per-exchange trace of the fused kernels (tuning key haloTrace = 1): 8 words for each of `count`
 * exchanges starting at sequence number firstSeq (a ring of 1024): [0] first push CTA started,
 * [1] ready flags published (device nanoseconds, %globaltimer), [2] / [3] nanoseconds the row blocks
 * spent waiting for the lower / upper ready word (summed over blocks, waits > 2 us only), [4] / [5]
 * how many blocks waited, [6] first boundary block started, [7] last boundary block past its wait.
 * Synchronises the stream.  0, or -1.
 */
int spgpuHaloTraceRead(spgpuHandle_t handle, unsigned long long* hostOut, int firstSeq, int count);

#ifdef __cplusplus
}
#endif

#endif /* SPGPU_EXT_H_ */
