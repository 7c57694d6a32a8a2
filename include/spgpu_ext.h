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
 * hellLongFactor, hellSplit, hdiaVariant, hdiaBlock, redBlocksPerSm,
 * vecBlocksPerSm (meanings: csrc/spgpu_internal.h).  Returns 0, or -1 for an
 * unknown key.
 */
int spgpuSetTuning(spgpuHandle_t handle, const char* key, int value);
int spgpuGetTuning(spgpuHandle_t handle, const char* key);

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

/* dRes[0] = sum x_i, left in device memory (deterministic; folds per-CTA partials). */
void spgpuDsumDev(spgpuHandle_t handle, int n, const __device double* x, __device double* dRes);

/*
 * z = (*dBetaNum / *dBetaDen) * betaSign * y + (*dAlphaNum / *dAlphaDen) *
 * alphaSign * x with the four scalars read from DEVICE memory at kernel time
 * (pass NULL for a numerator/denominator pair to mean 1).  This is the shape of
 * every CG update: x += (rr/pAp) p ; r -= (rr/pAp) Ap ; p = r + (rr'/rr) p.
 */
void spgpuDaxpbyDev(spgpuHandle_t handle, __device double* z, int n,
	const __device double* dBetaNum, const __device double* dBetaDen,
	double betaSign, const __device double* y,
	const __device double* dAlphaNum, const __device double* dAlphaDen,
	double alphaSign, const __device double* x);

/* ---- fused Krylov kernels --------------------- */

/*
 * HELL SpMV z = A*x fused with dRes[0] = sum_i x[xOffset+i]*z[i] (the p.Ap of CG)
 * in the SpMV epilogue.  xOffset = position of row 0's own entry inside x
 * (0 on one GPU, the lower halo width on a partition).  Two launches: the SpMV leaves one
 * partial per CTA in handle-owned scratch, spgpuDsumDev folds them (deterministic).
 */
void spgpuDhellspmvDot(spgpuHandle_t handle, __device double* z,
	const __device double* cM, const __device int* rP, int hackSize,
	const __device int* hackOffsets, const __device int* rS, int rows,
	const __device double* x, int baseIndex, int xOffset,
	__device double* dRes);

/*
 * Fused CG update: x += a*p ; r -= a*Ap ; dRrNew[0] = r.r, with a = *dRr / *dPAp read
 * from device memory.  One pass over four vectors instead of three kernels.
 */
void spgpuDcgUpdateDev(spgpuHandle_t handle, __device double* x, __device double* r,
	const __device double* p, const __device double* ap, int n,
	const __device double* dRr, const __device double* dPAp, __device double* dRrNew);

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
 * Halo push over NVLink: copy n elements src[0..n) of this GPU into a PEER
 * pointer (obtained through spgpuIpcOpenHandle) with 128-bit stores, then
 * release-store `flagValue` to the peer's flag word so the consumer can wait on
 * it (spgpuWaitFlag).  One kernel, on the handle's stream.
 */
void spgpuDhaloPush(spgpuHandle_t handle, double* peerDst, const double* src,
	int n, unsigned* peerFlag, unsigned flagValue);
/* Stream-ordered wait until *flag >= value (bounded spin, gives up after 2 s). */
void spgpuWaitFlag(spgpuHandle_t handle, const unsigned* flag, unsigned value);

/*
 * Fused halo exchange of a 1-D chain of ranks, ONE kernel per SpMV: waits until
 * the neighbours have acknowledged the previous halo (ackLo/ackHi, local flags,
 * value seq-1), pushes srcLo[0..n) into the lower neighbour's upper halo zone and
 * srcHi[0..n) into the upper neighbour's lower halo zone (peer pointers), release-
 * stores seq into the neighbours' ready flags and returns when this rank's own
 * ready flags (myReadyLo/myReadyHi) have reached seq.  NULL pointers = no
 * neighbour on that side.  spgpuHaloAck (after the SpMV has consumed the halos)
 * stores seq into the neighbours' ack flags.
 */
void spgpuDhaloExchange(spgpuHandle_t handle, double* peerDstLo, const double* srcLo,
	double* peerDstHi, const double* srcHi, int n, const unsigned* ackLo,
	const unsigned* ackHi, unsigned* peerReadyLo, unsigned* peerReadyHi,
	const unsigned* myReadyLo, const unsigned* myReadyHi, unsigned seq);
void spgpuHaloAck(spgpuHandle_t handle, unsigned* peerAckLo, unsigned* peerAckHi, unsigned seq);

/*
 * HELL SpMV of one row block FUSED with its halo exchange -- one kernel per
 * partitioned SpMV, transfer and multiply overlapped inside the launch.
 * xExt = [lower halo (haloN) | owned (rows) | upper halo (haloN)], columns of the
 * block index into it.  peerXLoUpperHalo / peerXHiLowerHalo are PEER pointers to the
 * neighbours' halo zones this rank fills (NULL = no neighbour on that side).
 * myFlags / peerFlagsLo / peerFlagsHi point at 4+ zero-initialised unsigned words per
 * rank: [0] ready-from-below, [1] ready-from-above, [2] ack-from-below,
 * [3] ack-from-above.  seq = 1, 2, 3, ... (same on every rank) numbers the exchanges.
 * The first CTAs push the two boundary planes over NVLink (after the neighbour
 * acknowledged the previous ones) and publish seq; most interior row blocks are scheduled
 * first, then the row blocks that read a halo zone (they wait on the local ready flag),
 * then a few waves of interior blocks; the last boundary CTA of each side acknowledges
 * that neighbour's halo.  A row block waits for ONE zone (the first ceil(haloN/128)
 * blocks for the lower, the last for the upper), so with neighbours on both sides the
 * block must own at least 2*haloN rows; shorter blocks use spgpuDhaloExchange + the plain
 * SpMV (spgpu_b200/mg.py does this by itself).
 */
void spgpuDhellspmvHalo(spgpuHandle_t handle, __device double* z, const __device double* y,
	double alpha, const __device double* cM, const __device int* rP, int hackSize,
	const __device int* hackOffsets, const __device int* rS, int avgNnzPerRow, int rows,
	__device double* xExt, double beta, int baseIndex, int haloN,
	double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq);

/* The same kernel with this rank's share of p.Ap = sum_i xExt[haloN+i]*z[i] left in dRes
 * (alpha = 1, beta = 0); partials per row block in handle scratch, folded by spgpuDsumDev. */
void spgpuDhellspmvHaloDot(spgpuHandle_t handle, __device double* z, const __device double* cM,
	const __device int* rP, int hackSize, const __device int* hackOffsets, const __device int* rS,
	int avgNnzPerRow, int rows, __device double* xExt, int baseIndex, int haloN,
	double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq, __device double* dRes);

/*
 * HDIA twin of spgpuDhellspmvHalo (same halo protocol, same flag words): the row block is in
 * HDIA layout with its diagonal offsets addressing xExt = [halo | owned | halo], i.e. global
 * offset + haloN, and cols = rows + 2*haloN (what mg.split_hdia / device_build.hdia_row_block
 * produce).  Arguments up to beta as spgpuDhdiaspmv (reference hdia.h:38-142).
 */
void spgpuDhdiaspmvHalo(spgpuHandle_t handle, __device double* z, const __device double* y, double alpha,
	const __device double* dM, const __device int* offsets, int hackSize, const __device int* hackOffsets,
	int rows, int cols, __device double* xExt, double beta, int haloN,
	double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq);

/* HDIA twin of spgpuDhellspmvHaloDot (alpha = 1, beta = 0; dRes[0] = sum_i xExt[haloN+i]*z[i]).  With no
 * neighbours (peer pointers NULL, haloN = 0, myFlags any 16-word device buffer) it is the single-GPU
 * fused HDIA SpMV + dot. */
void spgpuDhdiaspmvHaloDot(spgpuHandle_t handle, __device double* z, const __device double* dM,
	const __device int* offsets, int hackSize, const __device int* hackOffsets, int rows, int cols,
	__device double* xExt, int haloN, double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq, __device double* dRes);

/*
 * In-place sum all-reduce of ONE double over NVLink peer memory (latency-bound payload:
 * one round of remote 16-byte stores + local polling instead of an NCCL launch).
 * tables[r] = pointer (peer pointer for r != myRank) to rank r's zero-initialised table of
 * 2 * world 16-byte slots; world <= 16; seq = 1, 2, 3, ... identical on every rank.  All
 * ranks obtain the same bits (values are added in rank order).
 */
void spgpuAllreduceSumDev(spgpuHandle_t handle, __device double* dValue, int world, int myRank,
	void* const* tables, unsigned seq);

/*
 * Sequence numbers in device memory, so that a partitioned iteration can be captured ONCE in a CUDA
 * graph and replayed (a by-value seq would be frozen into the graph).  After spgpuSetSeqCounters the
 * fused halo kernels (spgpu?{hell,hdia}spmvHalo[Dot]) and spgpuAllreduceSumDev, when called with
 * seq == 0, take their sequence number from the counters: *dHaloSeq / *dAllreduceSeq hold the number of
 * COMPLETED exchanges / all-reduces (start them at the host-side count, or 0).  The all-reduce kernel
 * advances its counter itself; the halo counter is advanced by spgpuHaloSeqAdvance, a one-thread kernel
 * the caller puts after each fused SpMV (the SpMV kernel's CTAs must all read the same value however
 * late they are scheduled).  Returns 0, or -1 for a foreign handle.
 */
int spgpuSetSeqCounters(spgpuHandle_t handle, __device unsigned* dHaloSeq, __device unsigned* dAllreduceSeq);
void spgpuHaloSeqAdvance(spgpuHandle_t handle);

#ifdef __cplusplus
}
#endif

#endif /* SPGPU_EXT_H_ */
