/*
 * spgpu_mg.h -- row-partitioned multi-GPU SpMV and its Krylov companions as a C API.
 *
 * ADDITIVE: nothing here exists in the reference library, whose multi-GPU story is "one handle
 * per device, caller's threads, no communication" (reference core.h:88-93).  This is the C-level
 * boundary SURVEY 8(b) proposes for north_star item (3): ONE process drives N devices of one
 * NVSwitch box; the matrix is split in contiguous row blocks whose boundaries are multiples of
 * hackSize, so every block is a self-contained HELL matrix in the reference's own layout
 * (reference hell.h:45-169) and is multiplied by the library's own kernels.
 *
 * Each rank (device) keeps its slice of a vector inside
 *        x_ext = [ lower halo zone (w) | owned entries | upper halo zone (w) ]
 * with the block's column indices remapped to x_ext positions.  Before a product the zones are
 * filled with the neighbours' boundary entries:
 *   SPGPU_MG_FUSED   the exchange travels INSIDE the SpMV launch (spgpu?hellspmvHalo, spgpu_ext.h):
 *                    peer stores over NVLink + flag words, no host involvement; needs distinct devices with peer access;
 *   SPGPU_MG_EVENTS  a push kernel per rank + CUDA events between the ranks' streams; works on any
 *                    set of devices, including the same device listed several times (how the
 *                    single-GPU tests exercise the partition logic).
 * A matrix whose rows reach further than a neighbouring block (not banded) is multiplied in
 * ALL-GATHER mode instead: columns stay global and every rank gathers the whole x over NVLink
 * (peer-to-peer copies) before its product.
 *
 * All calls return spgpuStatus_t.  Calls are asynchronous on the ranks' streams unless they hand a
 * value back to the host (dot, nrm2, VectorGet, the optional residual of CgStep).  One host thread
 * at a time per context.
 */
#ifndef SPGPU_MG_H_
#define SPGPU_MG_H_

#include "spgpu_ext.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spgpuMgContext* spgpuMgHandle_t;
typedef struct spgpuMgMatrix* spgpuMgMatrix_t;
typedef struct spgpuMgVector* spgpuMgVector_t;
typedef struct spgpuMgCg* spgpuMgCg_t;

#define SPGPU_MG_AUTO    0
#define SPGPU_MG_FUSED   1
#define SPGPU_MG_EVENTS  2

/* ---- context --------------------------------------------------------------------------- */

/* One rank per entry of devices[0..n).  Enables peer access between the devices (where the
 * hardware allows) and creates one spgpu handle per rank.  n <= 16. */
spgpuStatus_t spgpuMgCreate(spgpuMgHandle_t* pMg, const int* devices, int n);
void spgpuMgDestroy(spgpuMgHandle_t mg);
int spgpuMgWorld(spgpuMgHandle_t mg);
/* the spGPU handle of a rank (its stream carries the rank's work; tuning keys can be set on it) */
spgpuHandle_t spgpuMgRankHandle(spgpuMgHandle_t mg, int rank);
/* SPGPU_MG_AUTO (default: FUSED where possible), SPGPU_MG_FUSED (SPGPU_UNSUPPORTED if the devices
 * do not allow it) or SPGPU_MG_EVENTS */
spgpuStatus_t spgpuMgSetExchange(spgpuMgHandle_t mg, int mode);
/* the mode in use: SPGPU_MG_FUSED or SPGPU_MG_EVENTS */
int spgpuMgExchange(spgpuMgHandle_t mg);
/* waits for every rank's stream; SPGPU_UNSPECIFIED if a CUDA error or a device-side wait time-out
 * (spgpuGetDeviceStatus) was recorded on any rank */
spgpuStatus_t spgpuMgSynchronize(spgpuMgHandle_t mg);

/* ---- matrices ---------------------------------------------------------------------------- */

/*
 * Splits a GLOBAL HELL matrix held in HOST memory (the arrays ellToHell produces, reference
 * hell_conv.h:29-63) over the ranks and uploads the blocks.  cols = length of x.  The halo width
 * is the furthest any row reaches outside its own block; if that exceeds a neighbouring block,
 * or a row that is not among the first / last `halo` rows of its block reads a halo column, the
 * matrix is kept in all-gather mode.
 */
#define SPGPU_DECL_MG_HELLCREATE(S, T, R)                                                    \
	spgpuStatus_t spgpuMg##S##hellCreate(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA,              \
		const __host T* cM, const __host int* rP, int hackSize, const __host int* hackOffsets, \
		const __host int* rS, int avgNnzPerRow, int rows, int cols, int baseIndex);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_MG_HELLCREATE)

/*
 * The partition spgpuMg?hellCreate would choose, without touching a device: bounds[0..world] = first row of every
 * block (multiples of hackSize), *halo = halo width in elements (the furthest reach of any row outside its block, rounded
 * up to 32), *allGather = 1 when the matrix has to be multiplied in all-gather mode (then *halo = 0).
 */
spgpuStatus_t spgpuMgHellPlan(int world, const __host int* rP, int hackSize, const __host int* hackOffsets,
	const __host int* rS, int rows, int baseIndex, int* bounds, int* halo, int* allGather);

/*
 * The same for a GLOBAL HDIA matrix in HOST memory (the arrays cooToHdia / diaToHdia produce, reference
 * hdia_conv.h:29-96; layout reference hdia.h:29-130).  HDIA addresses x relative to the row, so a row block is
 * the same hacks with hackOffsets re-based and every diagonal offset raised by the halo width (local row i reads
 * x_ext[i + offset + halo]); cells outside the matrix are stored as 0.  The halo is the furthest a NON-ZERO cell
 * reaches outside its block.  SPGPU_UNSUPPORTED when that exceeds a neighbouring block or the non-zero cells are not
 * banded within it (HDIA has no all-gather form: its columns are relative to the row -- convert such a matrix to HELL).
 */
#define SPGPU_DECL_MG_HDIACREATE(S, T, R)                                                    \
	spgpuStatus_t spgpuMg##S##hdiaCreate(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA,              \
		const __host T* dM, const __host int* offsets, int hackSize,                           \
		const __host int* hackOffsets, int rows, int cols);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_MG_HDIACREATE)

/* the partition spgpuMg?hdiaCreate would choose, without touching a device; *fits = 0 where it would refuse */
spgpuStatus_t spgpuMgHdiaPlan(int world, spgpuType_t type, const __host void* dM, const __host int* offsets,
	int hackSize, const __host int* hackOffsets, int rows, int cols, int* bounds, int* halo, int* fits);

/*
 * The same from per-rank blocks that are ALREADY partitioned: rank r owns blockRows[r] consecutive
 * rows (a multiple of hackSize for every rank but the last), its HELL arrays cM[r], rP[r] (elements[r]
 * entries each), hackOffsets[r], rS[r] hold LOCAL column indices into x_ext = [haloN | owned | haloN]
 * (global column - first owned row + haloN + baseIndex).  onDevice == 0: host arrays, uploaded here;
 * onDevice != 0: device arrays on the rank's device, used in place (the caller keeps ownership).
 * This is how a matrix too large to assemble on one host (BASELINE configs[4]: 512^3) is handed over.
 */
spgpuStatus_t spgpuMgHellCreateFromBlocks(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA, spgpuType_t type,
	int hackSize, int haloN, int baseIndex, int avgNnzPerRow, const int* blockRows,
	const void* const* cM, const int* const* rP, const int* const* hackOffsets, const int* const* rS,
	const long long* elements, int onDevice);

void spgpuMgMatrixDestroy(spgpuMgMatrix_t A);
/* halo width in elements (0 on one rank), or -1 for a matrix in all-gather mode */
int spgpuMgMatrixHalo(spgpuMgMatrix_t A);
int spgpuMgMatrixRows(spgpuMgMatrix_t A);
/* rows [*lo, *hi) of the global matrix live on `rank` */
void spgpuMgMatrixRowBlock(spgpuMgMatrix_t A, int rank, int* lo, int* hi);

/* ---- vectors (partitioned like the rows of a matrix) ---------------------------------------- */

spgpuStatus_t spgpuMgVectorCreate(spgpuMgMatrix_t A, spgpuMgVector_t* pV);
void spgpuMgVectorDestroy(spgpuMgVector_t v);
/* scatter a global host vector (rows entries of the matrix's value type) to its owners / gather it back (blocking) */
spgpuStatus_t spgpuMgVectorSet(spgpuMgVector_t v, const __host void* globalValues);
spgpuStatus_t spgpuMgVectorGet(spgpuMgVector_t v, __host void* globalValues);
/* device pointer to a rank's owned entries (on that rank's device) */
void* spgpuMgVectorLocal(spgpuMgVector_t v, int rank);

/* ---- operations ------------------------------------------------------------------------------- */

/*
 * z = alpha * A * x + beta * y over the partition (y may be NULL when beta == 0; z may be y; z must
 * not be x).  One kernel per rank in FUSED mode.  spgpuMg?hellspmv / spgpuMg?hdiaspmv accept a matrix of their
 * own format only (SPGPU_UNSUPPORTED otherwise); spgpuMg?spmv takes either.  dot / nrm2 block and return the GLOBAL value
 * (partials added in rank order: every run gives the same bits); dot is unconjugated like spgpu?dot.
 */
#define SPGPU_DECL_MG_OPS(S, T, R)                                                            \
	spgpuStatus_t spgpuMg##S##hellspmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y, \
		T alpha, spgpuMgMatrix_t A, spgpuMgVector_t x, T beta);                                 \
	spgpuStatus_t spgpuMg##S##hdiaspmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y, \
		T alpha, spgpuMgMatrix_t A, spgpuMgVector_t x, T beta);                                 \
	spgpuStatus_t spgpuMg##S##spmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y,    \
		T alpha, spgpuMgMatrix_t A, spgpuMgVector_t x, T beta);                                 \
	spgpuStatus_t spgpuMg##S##dot(spgpuMgHandle_t mg, __host T* result, spgpuMgVector_t a,      \
		spgpuMgVector_t b);                                                                     \
	spgpuStatus_t spgpuMg##S##nrm2(spgpuMgHandle_t mg, __host R* result, spgpuMgVector_t x);    \
	spgpuStatus_t spgpuMg##S##axpby(spgpuMgHandle_t mg, spgpuMgVector_t z, T beta,              \
		spgpuMgVector_t y, T alpha, spgpuMgVector_t x);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_MG_OPS)

/* ---- conjugate gradients on a double matrix, HELL or HDIA (BASELINE configs[4]: "plus CG step") --- */

/*
 * x = 0, r = p = b; *rr0 (may be NULL) = b.b.  Each CgStep is `iterations` iterations of
 *   Ap = A p, a = rr / p.Ap, x += a p, r -= a Ap, rr' = r.r, p = r + (rr'/rr) p.
 * FUSED mode: 4 launches per rank and iteration (SpMV + halo + p.Ap, fold + all-reduce, x/r update +
 * r.r + all-reduce, p update), every scalar stays on the devices, no host synchronisation; *rr (may
 * be NULL: then the call does not block) receives r.r after the last iteration.  EVENTS mode: the
 * blocking recurrence through spgpuMgDdot.
 */
spgpuStatus_t spgpuMgDcgCreate(spgpuMgMatrix_t A, spgpuMgCg_t* pCg);
spgpuStatus_t spgpuMgDcgStart(spgpuMgCg_t cg, spgpuMgVector_t b, __host double* rr0);
spgpuStatus_t spgpuMgDcgStep(spgpuMgCg_t cg, int iterations, __host double* rr);
spgpuMgVector_t spgpuMgDcgSolution(spgpuMgCg_t cg);
void spgpuMgDcgDestroy(spgpuMgCg_t cg);

#ifdef __cplusplus
}
#endif

#endif /* SPGPU_MG_H_ */
