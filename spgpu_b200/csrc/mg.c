/*
 * Row-partitioned multi-GPU layer, host side in C (include/spgpu_mg.h).  No reference counterpart:
 * the reference is one handle per device with no communication (reference core.h:88-93).
 *
 * One process, N ranks = N (device, spgpu handle, stream) triples.  A matrix is N self-contained HELL
 * blocks in the reference's own layout (reference hell.h:45-169) with column indices remapped into
 * each rank's x_ext = [halo | owned | halo] -- or N HDIA blocks (reference hdia.h:29-130) whose diagonal
 * offsets are shifted by the halo width; the products are the library's own kernels (spgpu?hellspmv /
 * spgpu?hdiaspmv, or spgpu?{hell,hdia}spmvHalo[Dot] of spgpu_ext.h when the halo exchange travels inside
 * the launch).  What this file adds is bookkeeping: the split, the remap, the vectors with their
 * zones, sequence numbers, and -- in EVENTS mode -- the CUDA events between the ranks' streams.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>

#include "spgpu_internal.h"
#include "spgpu_mg.h"

#define MG_MAX_RANKS 16
#define MG_MODE_HALO 0
#define MG_MODE_ALLGATHER 1
#define MG_KIND_HELL 0
#define MG_KIND_HDIA 1

struct spgpuMgContext {
	int world;
	int dev[MG_MAX_RANKS];
	spgpuHandle_t h[MG_MAX_RANKS];
	int fusedCapable;                 /* distinct devices, peer access between every pair */
	int direct[MG_MAX_RANKS][MG_MAX_RANKS];   /* rank r's kernels can store into rank q's memory (same device or peer access) */
	int exchange;                     /* SPGPU_MG_FUSED or SPGPU_MG_EVENTS */
	unsigned* flags[MG_MAX_RANKS];    /* device: SPGPU_HALO_FLAG_WORDS words per rank */
	void* arTable[MG_MAX_RANKS];      /* device: 2 * world slots of SPGPU_AR_SLOT_BYTES */
	void* dScalar[MG_MAX_RANKS];      /* device: 64 bytes of scratch for reduction results */
	char* hostSlots;                  /* pinned: 64 bytes per rank */
	cudaEvent_t evPush[MG_MAX_RANKS], evDone[MG_MAX_RANKS];
	int doneValid[MG_MAX_RANKS];
	unsigned haloSeq, arSeq;
};

struct spgpuMgMatrix {
	spgpuMgHandle_t mg;
	spgpuType_t type;
	size_t esize;
	int kind;                         /* MG_KIND_HELL / MG_KIND_HDIA */
	int rows, hackSize, baseIndex, avg;
	int mode;                         /* MG_MODE_HALO / MG_MODE_ALLGATHER */
	int halo;
	int lo[MG_MAX_RANKS], hi[MG_MAX_RANKS];
	void* cM[MG_MAX_RANKS];           /* HELL values / HDIA dM */
	int* rP[MG_MAX_RANKS];            /* HELL column indices / HDIA diagonal offsets (already + halo) */
	int* hackOffsets[MG_MAX_RANKS];   /* HELL: element offsets, one per hack / HDIA: diagonal prefix, hacks + 1 */
	int* rS[MG_MAX_RANKS];            /* HELL row sizes / HDIA: NULL */
	int owned;                        /* device arrays allocated here (freed by MatrixDestroy) */
};

struct spgpuMgVector {
	spgpuMgMatrix_t A;
	void* ext[MG_MAX_RANKS];          /* device: [halo | owned | halo] */
	void* full[MG_MAX_RANKS];         /* device: the whole vector (all-gather mode, allocated on first use) */
	spgpuHaloLinks links[MG_MAX_RANKS];
};

struct spgpuMgCg {
	spgpuMgMatrix_t A;
	spgpuMgVector_t x, r, p, ap;
	double* s[MG_MAX_RANKS];          /* device scalars: [rr, pAp, rr', -]; slots 0 and 2 swap roles every iteration */
	unsigned iter;                    /* iterations done on the devices: r.r of the current iterate is in slot 2 * (iter & 1) */
	double rr;                        /* host copy (blocking recurrence) */
};

static void use_rank(spgpuMgHandle_t mg, int r)
{
	cudaSetDevice(mg->dev[r]);
}

static cudaStream_t rank_stream(spgpuMgHandle_t mg, int r)
{
	return mg->h[r]->currentStream;
}

/* ---- context ---------------------------------------------------------------------------------- */

spgpuStatus_t spgpuMgCreate(spgpuMgHandle_t* pMg, const int* devices, int n)
{
	int previous = 0, r, q, distinct = 1, peers = 1;
	spgpuMgHandle_t mg;
	*pMg = NULL;
	if (n < 1 || n > MG_MAX_RANKS || !devices)
		return SPGPU_UNSUPPORTED;
	mg = (spgpuMgHandle_t)calloc(1, sizeof(*mg));
	if (!mg)
		return SPGPU_OUTOFMEMORY;
	mg->world = n;
	cudaGetDevice(&previous);
	for (r = 0; r < n; ++r) {
		mg->dev[r] = devices[r];
		for (q = 0; q < r; ++q)
			if (devices[q] == devices[r])
				distinct = 0;
	}
	for (r = 0; r < n; ++r) {
		spgpuStatus_t st = spgpuCreate(&mg->h[r], mg->dev[r]);
		if (st != SPGPU_SUCCESS) {
			cudaSetDevice(previous);
			spgpuMgDestroy(mg);
			return st;
		}
	}
	for (r = 0; r < n; ++r) {
		use_rank(mg, r);
		for (q = 0; q < n; ++q) {
			int can = 0;
			cudaError_t e;
			mg->direct[r][q] = 1;
			if (mg->dev[q] == mg->dev[r])
				continue;
			if (cudaDeviceCanAccessPeer(&can, mg->dev[r], mg->dev[q]) != cudaSuccess || !can) {
				mg->direct[r][q] = 0;
				peers = 0;
				continue;
			}
			e = cudaDeviceEnablePeerAccess(mg->dev[q], 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
				mg->direct[r][q] = 0;
				peers = 0;
			}
			(void)cudaGetLastError();
		}
	}
	mg->fusedCapable = distinct && peers;
	mg->exchange = (mg->fusedCapable || n == 1) ? SPGPU_MG_FUSED : SPGPU_MG_EVENTS;
	if (cudaHostAlloc((void**)&mg->hostSlots, (size_t)64 * n, cudaHostAllocDefault) != cudaSuccess) {
		cudaSetDevice(previous);
		spgpuMgDestroy(mg);
		return SPGPU_OUTOFMEMORY;
	}
	for (r = 0; r < n; ++r) {
		const size_t tableBytes = (size_t)2 * n * SPGPU_AR_SLOT_BYTES;
		cudaError_t e;
		use_rank(mg, r);
		e = cudaMalloc((void**)&mg->flags[r], SPGPU_HALO_FLAG_WORDS * sizeof(unsigned));
		if (e == cudaSuccess) e = cudaMemset(mg->flags[r], 0, SPGPU_HALO_FLAG_WORDS * sizeof(unsigned));
		if (e == cudaSuccess) e = cudaMalloc(&mg->arTable[r], tableBytes);
		if (e == cudaSuccess) e = cudaMemset(mg->arTable[r], 0, tableBytes);
		if (e == cudaSuccess) e = cudaMalloc(&mg->dScalar[r], 64);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&mg->evPush[r], cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&mg->evDone[r], cudaEventDisableTiming);
		/* every kernel that may wait for a peer is loaded NOW: a first launch (lazy loading) waits for the kernels
		 * running on the device, and one host thread launches all the ranks (spgpu_ext.h: spgpuPreloadHaloKernels) */
		if (e == cudaSuccess && (spgpuPreloadHaloKernels() != 0 || spgpuPreloadKrylovKernels() != 0))
			e = cudaErrorUnknown;
		if (e == cudaSuccess) {
			/* ... and the kernels that run between them in an iteration, by running them once */
			double* d = (double*)mg->dScalar[r];
			e = cudaMemset(d, 0, 64);
			spgpuDscal(mg->h[r], d, 1, 1.0, d + 1);
			spgpuDdotDev(mg->h[r], 1, d, d + 1, d + 2);
			spgpuDaxpby(mg->h[r], d, 1, 1.0, d + 1, 1.0, d + 2);
		}
		if (e == cudaSuccess) e = cudaDeviceSynchronize();
		if (e != cudaSuccess) {
			cudaSetDevice(previous);
			spgpuMgDestroy(mg);
			return e == cudaErrorMemoryAllocation ? SPGPU_OUTOFMEMORY : SPGPU_UNSPECIFIED;
		}
	}
	cudaSetDevice(previous);
	*pMg = mg;
	return SPGPU_SUCCESS;
}

void spgpuMgDestroy(spgpuMgHandle_t mg)
{
	int previous = 0, r;
	if (!mg)
		return;
	cudaGetDevice(&previous);
	for (r = 0; r < mg->world; ++r) {
		if (!mg->h[r])
			continue;
		use_rank(mg, r);
		cudaStreamSynchronize(rank_stream(mg, r));
		if (mg->flags[r]) cudaFree(mg->flags[r]);
		if (mg->arTable[r]) cudaFree(mg->arTable[r]);
		if (mg->dScalar[r]) cudaFree(mg->dScalar[r]);
		if (mg->evPush[r]) cudaEventDestroy(mg->evPush[r]);
		if (mg->evDone[r]) cudaEventDestroy(mg->evDone[r]);
		spgpuDestroy(mg->h[r]);
	}
	if (mg->hostSlots)
		cudaFreeHost(mg->hostSlots);
	cudaSetDevice(previous);
	free(mg);
}

int spgpuMgWorld(spgpuMgHandle_t mg)
{
	return mg ? mg->world : 0;
}

spgpuHandle_t spgpuMgRankHandle(spgpuMgHandle_t mg, int rank)
{
	return (mg && rank >= 0 && rank < mg->world) ? mg->h[rank] : NULL;
}

spgpuStatus_t spgpuMgSetExchange(spgpuMgHandle_t mg, int mode)
{
	if (!mg)
		return SPGPU_UNSPECIFIED;
	if (mode == SPGPU_MG_AUTO)
		mode = (mg->fusedCapable || mg->world == 1) ? SPGPU_MG_FUSED : SPGPU_MG_EVENTS;
	if (mode == SPGPU_MG_FUSED && !(mg->fusedCapable || mg->world == 1))
		return SPGPU_UNSUPPORTED;
	if (mode != SPGPU_MG_FUSED && mode != SPGPU_MG_EVENTS)
		return SPGPU_UNSUPPORTED;
	if (mode != mg->exchange)
		spgpuMgSynchronize(mg);           /* the two protocols must not overlap on the same zones */
	mg->exchange = mode;
	return SPGPU_SUCCESS;
}

int spgpuMgExchange(spgpuMgHandle_t mg)
{
	return mg ? mg->exchange : 0;
}

spgpuStatus_t spgpuMgSynchronize(spgpuMgHandle_t mg)
{
	int previous = 0, r, bad = 0;
	if (!mg)
		return SPGPU_UNSPECIFIED;
	cudaGetDevice(&previous);
	for (r = 0; r < mg->world; ++r) {
		use_rank(mg, r);
		if (cudaStreamSynchronize(rank_stream(mg, r)) != cudaSuccess)
			bad = 1;
	}
	for (r = 0; r < mg->world; ++r)
		if (spgpuGetDeviceStatus(mg->h[r], 0) > 0)
			bad = 1;
	cudaSetDevice(previous);
	return bad ? SPGPU_UNSPECIFIED : SPGPU_SUCCESS;
}

/* ---- matrices ----------------------------------------------------------------------------------- */

static spgpuMgMatrix_t matrix_new(spgpuMgHandle_t mg, spgpuType_t type, int hackSize, int baseIndex, int avg)
{
	spgpuMgMatrix_t A = (spgpuMgMatrix_t)calloc(1, sizeof(*A));
	if (!A)
		return NULL;
	A->mg = mg;
	A->type = type;
	A->esize = spgpuSizeOf(type);
	A->hackSize = hackSize;
	A->baseIndex = baseIndex;
	A->avg = avg > 0 ? avg : 1;
	return A;
}

static cudaError_t upload(void** dst, const void* src, size_t bytes)
{
	cudaError_t e = cudaMalloc(dst, bytes ? bytes : 1);
	if (e == cudaSuccess && bytes)
		e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
	return e;
}

spgpuStatus_t spgpuMgHellCreateFromBlocks(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA, spgpuType_t type,
	int hackSize, int haloN, int baseIndex, int avgNnzPerRow, const int* blockRows,
	const void* const* cM, const int* const* rP, const int* const* hackOffsets, const int* const* rS,
	const long long* elements, int onDevice)
{
	int previous = 0, r, at = 0;
	spgpuMgMatrix_t A;
	*pA = NULL;
	if (!mg || hackSize <= 0 || hackSize % 32 != 0 || haloN < 0 || type < SPGPU_TYPE_FLOAT || type > SPGPU_TYPE_COMPLEX_DOUBLE)
		return SPGPU_UNSUPPORTED;
	for (r = 0; r < mg->world; ++r) {
		if (blockRows[r] < 0 || (r < mg->world - 1 && blockRows[r] % hackSize != 0))
			return SPGPU_UNSUPPORTED;
		if (mg->world > 1 && blockRows[r] < haloN)      /* a rank sends its first / last haloN OWNED entries */
			return SPGPU_UNSUPPORTED;
	}
	A = matrix_new(mg, type, hackSize, baseIndex, avgNnzPerRow);
	if (!A)
		return SPGPU_OUTOFMEMORY;
	A->mode = MG_MODE_HALO;
	A->halo = haloN;                                     /* also on one rank: the blocks' columns count from x_ext's lower zone */
	A->owned = !onDevice;
	cudaGetDevice(&previous);
	for (r = 0; r < mg->world; ++r) {
		const int n = blockRows[r];
		const int hacks = (n + hackSize - 1) / hackSize;
		A->lo[r] = at;
		A->hi[r] = at + n;
		at += n;
		if (onDevice) {
			A->cM[r] = (void*)cM[r];
			A->rP[r] = (int*)rP[r];
			A->hackOffsets[r] = (int*)hackOffsets[r];
			A->rS[r] = (int*)rS[r];
		} else {
			cudaError_t e;
			use_rank(mg, r);
			e = upload(&A->cM[r], cM[r], (size_t)elements[r] * A->esize);
			if (e == cudaSuccess) e = upload((void**)&A->rP[r], rP[r], (size_t)elements[r] * sizeof(int));
			if (e == cudaSuccess) e = upload((void**)&A->hackOffsets[r], hackOffsets[r], (size_t)hacks * sizeof(int));
			if (e == cudaSuccess) e = upload((void**)&A->rS[r], rS[r], (size_t)n * sizeof(int));
			if (e != cudaSuccess) {
				cudaSetDevice(previous);
				spgpuMgMatrixDestroy(A);
				return e == cudaErrorMemoryAllocation ? SPGPU_OUTOFMEMORY : SPGPU_UNSPECIFIED;
			}
		}
		/* per-row-block partials of the fused SpMV + dot: size the handle's scratch now, not inside an iteration */
		use_rank(mg, r);
		spgpuReserveScratch(mg->h[r], ((size_t)(n + 127) / 128 + 1) * 4 * 16);
	}
	A->rows = at;
	cudaSetDevice(previous);
	*pA = A;
	return SPGPU_SUCCESS;
}

/* entry k of row i of a HELL matrix: hackOffsets[i / hackSize] + k * hackSize + i % hackSize (reference hell.h:45-59) */
static long long hell_at(const int* hackOffsets, int hackSize, int i, int k)
{
	return (long long)hackOffsets[i / hackSize] + (long long)k * hackSize + i % hackSize;
}

/* contiguous, near-equal row blocks whose boundaries are multiples of hackSize: bounds[0..world] */
static void block_bounds(int world, int rows, int hackSize, int* bounds)
{
	const long long units = ((long long)rows + hackSize - 1) / hackSize;
	int r;
	for (r = 0; r < world; ++r) {
		const long long b = (units * r / world) * hackSize;
		bounds[r] = (int)(b < rows ? b : rows);
	}
	bounds[world] = rows;
}

/*
 * The partition of a global HELL matrix over `world` ranks, host arithmetic only (no device is touched, so it can be
 * checked without a GPU): bounds[0..world] = first row of every block (multiples of hackSize), *halo = the furthest any
 * row reaches outside its own block rounded up to 32 (0 on one rank), *allGather = 1 when halo mode is impossible --
 * a block owns fewer than halo rows (a zone would need entries of a rank two hops away), or a row that is not among
 * the first / last halo rows of its block reads a halo column (interior rows are multiplied before the zones arrive).
 */
spgpuStatus_t spgpuMgHellPlan(int world, const int* rP, int hackSize, const int* hackOffsets, const int* rS, int rows,
	int baseIndex, int* bounds, int* halo, int* allGather)
{
	int r, i, k, banded = 1;
	long long reach = 0;
	if (world < 1 || world > MG_MAX_RANKS || rows < 0 || hackSize <= 0 || hackSize % 32 != 0)
		return SPGPU_UNSUPPORTED;
	block_bounds(world, rows, hackSize, bounds);
	*halo = 0;
	*allGather = 0;
	if (world == 1)
		return SPGPU_SUCCESS;
	/* how far does any row reach outside its own block? */
	for (r = 0; r < world; ++r)
		for (i = bounds[r]; i < bounds[r + 1]; ++i)
			for (k = 0; k < rS[i]; ++k) {
				const long long g = (long long)rP[hell_at(hackOffsets, hackSize, i, k)] - baseIndex;
				if (g < bounds[r] && bounds[r] - g > reach) reach = bounds[r] - g;
				if (g >= bounds[r + 1] && g - bounds[r + 1] + 1 > reach) reach = g - bounds[r + 1] + 1;
			}
	if (reach > 0x7fffffe0ll)
		reach = 0x7fffffe0ll;
	*halo = (int)(((reach + 31) / 32) * 32);
	if (*halo == 0)
		*halo = 32;
	for (r = 0; r < world && banded; ++r) {
		const int lo = bounds[r], hi = bounds[r + 1];
		if (hi - lo < *halo)
			banded = 0;
		for (i = lo; i < hi && banded; ++i)
			for (k = 0; k < rS[i]; ++k) {
				const long long g = (long long)rP[hell_at(hackOffsets, hackSize, i, k)] - baseIndex;
				if ((g < lo && i - lo >= *halo) || (g >= hi && i - lo < (hi - lo) - *halo)) {
					banded = 0;
					break;
				}
			}
	}
	if (!banded) {
		*allGather = 1;
		*halo = 0;
	}
	return SPGPU_SUCCESS;
}

static spgpuStatus_t hell_create(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA, spgpuType_t type, const void* cM,
	const int* rP, int hackSize, const int* hackOffsets, const int* rS, int avg, int rows, int cols, int baseIndex)
{
	const int world = mg ? mg->world : 0;
	const size_t esize = spgpuSizeOf(type);
	int bound[MG_MAX_RANKS + 1], r, i, k, banded = 1, previous = 0, allGather = 0;
	long long total = 0;
	int halo = 0;
	spgpuMgMatrix_t A;
	*pA = NULL;
	if (!mg || rows < 0 || cols != rows || hackSize <= 0 || hackSize % 32 != 0)
		return SPGPU_UNSUPPORTED;        /* vectors are partitioned like the rows: square matrices */
	if (spgpuMgHellPlan(world, rP, hackSize, hackOffsets, rS, rows, baseIndex, bound, &halo, &allGather) != SPGPU_SUCCESS)
		return SPGPU_UNSUPPORTED;
	banded = !allGather;
	if (rows > 0) {
		const int lastHack = (rows + hackSize - 1) / hackSize - 1;
		int deepest = 0;
		for (i = lastHack * hackSize; i < rows; ++i)
			if (rS[i] > deepest)
				deepest = rS[i];
		total = (long long)hackOffsets[lastHack] + (long long)deepest * hackSize;
	}
	A = matrix_new(mg, type, hackSize, baseIndex, avg);
	if (!A)
		return SPGPU_OUTOFMEMORY;
	A->mode = banded ? MG_MODE_HALO : MG_MODE_ALLGATHER;
	A->halo = banded ? halo : 0;
	A->rows = rows;
	A->owned = 1;
	cudaGetDevice(&previous);
	for (r = 0; r < world; ++r) {
		const int lo = bound[r], hi = bound[r + 1], n = hi - lo;
		const int h0 = lo / hackSize, h1 = (hi + hackSize - 1) / hackSize, hacks = h1 - h0;
		const long long e0 = n > 0 ? hackOffsets[h0] : 0;
		const long long e1 = n > 0 ? (hi < rows ? hackOffsets[h1] : total) : 0;
		const size_t count = (size_t)(e1 - e0);
		int* idx = (int*)malloc((count ? count : 1) * sizeof(int));
		int* hoff = (int*)malloc((hacks > 0 ? hacks : 1) * sizeof(int));
		cudaError_t e = cudaSuccess;
		A->lo[r] = lo;
		A->hi[r] = hi;
		if (!idx || !hoff) {
			free(idx); free(hoff);
			cudaSetDevice(previous);
			spgpuMgMatrixDestroy(A);
			return SPGPU_OUTOFMEMORY;
		}
		memcpy(idx, rP + e0, count * sizeof(int));
		for (i = 0; i < hacks; ++i)
			hoff[i] = (int)(hackOffsets[h0 + i] - e0);
		if (A->mode == MG_MODE_HALO) {
			/* remap the slots that exist (padding is undefined and stays so): x_ext position = g - (lo - halo) */
			const long long shift = (long long)lo - A->halo;
			for (i = lo; i < hi; ++i)
				for (k = 0; k < rS[i]; ++k) {
					const long long at = hell_at(hackOffsets, hackSize, i, k) - e0;
					idx[at] = (int)(idx[at] - shift);
				}
		}
		use_rank(mg, r);
		e = upload(&A->cM[r], (const char*)cM + (size_t)e0 * esize, count * esize);
		if (e == cudaSuccess) e = upload((void**)&A->rP[r], idx, count * sizeof(int));
		if (e == cudaSuccess) e = upload((void**)&A->hackOffsets[r], hoff, (size_t)hacks * sizeof(int));
		if (e == cudaSuccess) e = upload((void**)&A->rS[r], rS + lo, (size_t)n * sizeof(int));
		free(idx);
		free(hoff);
		if (e != cudaSuccess) {
			cudaSetDevice(previous);
			spgpuMgMatrixDestroy(A);
			return e == cudaErrorMemoryAllocation ? SPGPU_OUTOFMEMORY : SPGPU_UNSPECIFIED;
		}
		spgpuReserveScratch(mg->h[r], ((size_t)(n + 127) / 128 + 1) * 4 * 16);
	}
	cudaSetDevice(previous);
	*pA = A;
	return SPGPU_SUCCESS;
}

/* ---- HDIA ---------------------------------------------------------------------------------------- */

static int cell_nonzero(spgpuType_t type, const void* p)
{
	switch (type) {
	case SPGPU_TYPE_FLOAT: return *(const float*)p != 0.0f;
	case SPGPU_TYPE_DOUBLE: return *(const double*)p != 0.0;
	case SPGPU_TYPE_COMPLEX_FLOAT: return ((const float*)p)[0] != 0.0f || ((const float*)p)[1] != 0.0f;
	case SPGPU_TYPE_COMPLEX_DOUBLE: return ((const double*)p)[0] != 0.0 || ((const double*)p)[1] != 0.0;
	default: return 1;
	}
}

/*
 * The partition of a global HDIA matrix (reference hdia.h:29-130: hack h holds diagonals
 * [hackOffsets[h], hackOffsets[h+1]), diagonal d is hackSize values dM[d*hackSize + j] of rows h*hackSize + j
 * at column row + offsets[d]) over `world` ranks, host arithmetic only.  HDIA addresses x relative to the row,
 * so a row block is the same hacks with hackOffsets re-based; the halo is the furthest a NON-ZERO cell reaches
 * outside its block, rounded up to 32.  *fits = 0 when a block owns fewer than halo rows or a non-zero cell of a
 * row outside the first / last halo rows reads a halo column: such a matrix is not partitioned (there is no
 * all-gather form of HDIA: its columns are relative to the row).
 */
spgpuStatus_t spgpuMgHdiaPlan(int world, spgpuType_t type, const void* dM, const int* offsets, int hackSize,
	const int* hackOffsets, int rows, int cols, int* bounds, int* halo, int* fits)
{
	const size_t es = spgpuSizeOf(type);
	int r, pass;
	long long reach = 0;
	if (world < 1 || world > MG_MAX_RANKS || rows < 0 || cols < 0 || hackSize <= 0 || hackSize % 32 != 0 || es == 0)
		return SPGPU_UNSUPPORTED;
	block_bounds(world, rows, hackSize, bounds);
	*halo = 0;
	*fits = 1;
	if (world == 1)
		return SPGPU_SUCCESS;
	/* pass 0 measures the reach, pass 1 checks the band against the halo it gives */
	for (pass = 0; pass < 2 && *fits; ++pass) {
		for (r = 0; r < world && *fits; ++r) {
			const int lo = bounds[r], hi = bounds[r + 1];
			int h;
			if (pass == 1 && hi - lo < *halo) {
				*fits = 0;
				break;
			}
			for (h = lo / hackSize; h < (hi + hackSize - 1) / hackSize && *fits; ++h) {
				int d, j;
				for (d = hackOffsets[h]; d < hackOffsets[h + 1] && *fits; ++d)
					for (j = 0; j < hackSize; ++j) {
						const long long row = (long long)h * hackSize + j, g = row + offsets[d];
						if (row >= hi)
							break;
						if (g < 0 || g >= cols || (g >= lo && g < hi))
							continue;
						if (!cell_nonzero(type, (const char*)dM + ((size_t)d * hackSize + j) * es))
							continue;
						if (pass == 0) {
							if (g < lo && lo - g > reach) reach = lo - g;
							if (g >= hi && g - hi + 1 > reach) reach = g - hi + 1;
						} else if ((g < lo && row - lo >= *halo) || (g >= hi && row - lo < (hi - lo) - *halo)) {
							*fits = 0;
							break;
						}
					}
			}
		}
		if (pass == 0) {
			if (reach > 0x7fffffe0ll)
				reach = 0x7fffffe0ll;
			*halo = (int)(((reach + 31) / 32) * 32);
			if (*halo == 0)
				*halo = 32;
		}
	}
	return SPGPU_SUCCESS;
}

static spgpuStatus_t hdia_create(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA, spgpuType_t type, const void* dM,
	const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols)
{
	const int world = mg ? mg->world : 0;
	const size_t es = spgpuSizeOf(type);
	int bound[MG_MAX_RANKS + 1], r, halo = 0, fits = 1, previous = 0;
	spgpuMgMatrix_t A;
	*pA = NULL;
	if (!mg || rows < 0 || cols != rows)
		return SPGPU_UNSUPPORTED;        /* vectors are partitioned like the rows: square matrices */
	if (spgpuMgHdiaPlan(world, type, dM, offsets, hackSize, hackOffsets, rows, cols, bound, &halo, &fits) != SPGPU_SUCCESS || !fits)
		return SPGPU_UNSUPPORTED;
	A = matrix_new(mg, type, hackSize, 0, 1);
	if (!A)
		return SPGPU_OUTOFMEMORY;
	A->kind = MG_KIND_HDIA;
	A->mode = MG_MODE_HALO;
	A->halo = halo;
	A->rows = rows;
	A->owned = 1;
	cudaGetDevice(&previous);
	for (r = 0; r < world; ++r) {
		const int lo = bound[r], hi = bound[r + 1], n = hi - lo;
		const int h0 = lo / hackSize, h1 = (hi + hackSize - 1) / hackSize, hacks = h1 - h0;
		const int d0 = n > 0 ? hackOffsets[h0] : 0, d1 = n > 0 ? hackOffsets[h1] : 0, diags = d1 - d0;
		const size_t cells = (size_t)diags * hackSize;
		char* vals = (char*)malloc(cells > 0 ? cells * es : 1);
		int* offs = (int*)malloc((diags > 0 ? diags : 1) * sizeof(int));
		int* hoff = (int*)malloc((size_t)(hacks + 1) * sizeof(int));
		cudaError_t e = cudaSuccess;
		int h, d, j;
		A->lo[r] = lo;
		A->hi[r] = hi;
		if (!vals || !offs || !hoff) {
			free(vals); free(offs); free(hoff);
			cudaSetDevice(previous);
			spgpuMgMatrixDestroy(A);
			return SPGPU_OUTOFMEMORY;
		}
		memcpy(vals, (const char*)dM + (size_t)d0 * hackSize * es, cells * es);
		for (h = 0; h <= hacks; ++h)
			hoff[h] = n > 0 ? hackOffsets[h0 + h] - d0 : 0;
		for (d = 0; d < diags; ++d)
			offs[d] = offsets[d0 + d] + halo;     /* local row i reads x_ext[i + offset + halo] */
		/* cells outside the MATRIX now fall inside the window [0, n + 2*halo) the kernel tests: store 0 there
		 * (the conversions do, reference hdia_conv.h:60-96, but nothing obliges a caller's array to) */
		for (h = 0; h < hacks; ++h)
			for (d = hoff[h]; d < hoff[h + 1]; ++d)
				for (j = 0; j < hackSize; ++j) {
					const long long row = (long long)lo + (long long)h * hackSize + j, g = row + offsets[d0 + d];
					if (row >= hi || g < 0 || g >= cols)
						memset(vals + ((size_t)d * hackSize + j) * es, 0, es);
				}
		use_rank(mg, r);
		e = upload(&A->cM[r], vals, cells * es);
		if (e == cudaSuccess) e = upload((void**)&A->rP[r], offs, (size_t)diags * sizeof(int));
		if (e == cudaSuccess) e = upload((void**)&A->hackOffsets[r], hoff, (size_t)(hacks + 1) * sizeof(int));
		free(vals);
		free(offs);
		free(hoff);
		if (e != cudaSuccess) {
			cudaSetDevice(previous);
			spgpuMgMatrixDestroy(A);
			return e == cudaErrorMemoryAllocation ? SPGPU_OUTOFMEMORY : SPGPU_UNSPECIFIED;
		}
		spgpuReserveScratch(mg->h[r], ((size_t)(n + 127) / 128 + 1) * 4 * 16);
	}
	cudaSetDevice(previous);
	*pA = A;
	return SPGPU_SUCCESS;
}

void spgpuMgMatrixDestroy(spgpuMgMatrix_t A)
{
	int previous = 0, r;
	if (!A)
		return;
	cudaGetDevice(&previous);
	for (r = 0; r < A->mg->world; ++r) {
		use_rank(A->mg, r);
		cudaStreamSynchronize(rank_stream(A->mg, r));
		if (A->owned) {
			if (A->cM[r]) cudaFree(A->cM[r]);
			if (A->rP[r]) cudaFree(A->rP[r]);
			if (A->hackOffsets[r]) cudaFree(A->hackOffsets[r]);
			if (A->rS[r]) cudaFree(A->rS[r]);
		}
	}
	cudaSetDevice(previous);
	free(A);
}

int spgpuMgMatrixHalo(spgpuMgMatrix_t A)
{
	return A->mode == MG_MODE_ALLGATHER ? -1 : A->halo;
}

int spgpuMgMatrixRows(spgpuMgMatrix_t A)
{
	return A->rows;
}

void spgpuMgMatrixRowBlock(spgpuMgMatrix_t A, int rank, int* lo, int* hi)
{
	*lo = A->lo[rank];
	*hi = A->hi[rank];
}

/* ---- vectors ------------------------------------------------------------------------------------- */

static void* owned_ptr(spgpuMgVector_t v, int r)
{
	return (char*)v->ext[r] + (size_t)v->A->halo * v->A->esize;
}

spgpuStatus_t spgpuMgVectorCreate(spgpuMgMatrix_t A, spgpuMgVector_t* pV)
{
	spgpuMgHandle_t mg = A->mg;
	const size_t es = A->esize, w = (size_t)A->halo;
	int previous = 0, r;
	spgpuMgVector_t v = (spgpuMgVector_t)calloc(1, sizeof(*v));
	*pV = NULL;
	if (!v)
		return SPGPU_OUTOFMEMORY;
	v->A = A;
	cudaGetDevice(&previous);
	for (r = 0; r < mg->world; ++r) {
		const size_t n = (size_t)(A->hi[r] - A->lo[r]);
		const size_t extBytes = (n + 2 * w) * es;
		cudaError_t e;
		use_rank(mg, r);
		e = cudaMalloc(&v->ext[r], extBytes ? extBytes : 16);
		if (e == cudaSuccess) e = cudaMemset(v->ext[r], 0, extBytes);
		if (e != cudaSuccess) {
			cudaSetDevice(previous);
			spgpuMgVectorDestroy(v);
			return e == cudaErrorMemoryAllocation ? SPGPU_OUTOFMEMORY : SPGPU_UNSPECIFIED;
		}
	}
	/* where each rank's fused kernels find the neighbours' zones of THIS vector */
	for (r = 0; r < mg->world; ++r) {
		spgpuHaloLinks* k = &v->links[r];
		memset(k, 0, sizeof(*k));
		if (w == 0 || mg->world == 1)
			continue;
		if (r > 0) {                       /* lower neighbour's UPPER zone */
			const size_t nb = (size_t)(A->hi[r - 1] - A->lo[r - 1]);
			k->peerLoUpperZone = (char*)v->ext[r - 1] + (w + nb) * es;
			k->peerFlagsLo = mg->flags[r - 1];
		}
		if (r < mg->world - 1) {           /* upper neighbour's LOWER zone */
			k->peerHiLowerZone = v->ext[r + 1];
			k->peerFlagsHi = mg->flags[r + 1];
		}
		k->myFlags = mg->flags[r];
	}
	for (r = 0; r < mg->world; ++r) {
		use_rank(mg, r);
		cudaDeviceSynchronize();
	}
	cudaSetDevice(previous);
	*pV = v;
	return SPGPU_SUCCESS;
}

void spgpuMgVectorDestroy(spgpuMgVector_t v)
{
	int previous = 0, r;
	if (!v)
		return;
	cudaGetDevice(&previous);
	for (r = 0; r < v->A->mg->world; ++r) {
		use_rank(v->A->mg, r);
		cudaStreamSynchronize(rank_stream(v->A->mg, r));
		if (v->ext[r]) cudaFree(v->ext[r]);
		if (v->full[r]) cudaFree(v->full[r]);
	}
	cudaSetDevice(previous);
	free(v);
}

static spgpuStatus_t vector_copy(spgpuMgVector_t v, void* host, int toDevice)
{
	spgpuMgMatrix_t A = v->A;
	int previous = 0, r, bad = 0;
	cudaGetDevice(&previous);
	for (r = 0; r < A->mg->world; ++r) {
		const size_t bytes = (size_t)(A->hi[r] - A->lo[r]) * A->esize;
		char* hp = (char*)host + (size_t)A->lo[r] * A->esize;
		use_rank(A->mg, r);
		if (bytes && cudaMemcpyAsync(toDevice ? owned_ptr(v, r) : (void*)hp, toDevice ? (const void*)hp : owned_ptr(v, r), bytes,
				toDevice ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, rank_stream(A->mg, r)) != cudaSuccess)
			bad = 1;
	}
	for (r = 0; r < A->mg->world; ++r) {
		use_rank(A->mg, r);
		if (cudaStreamSynchronize(rank_stream(A->mg, r)) != cudaSuccess)
			bad = 1;
	}
	cudaSetDevice(previous);
	return bad ? SPGPU_UNSPECIFIED : SPGPU_SUCCESS;
}

spgpuStatus_t spgpuMgVectorSet(spgpuMgVector_t v, const void* globalValues)
{
	return vector_copy(v, (void*)globalValues, 1);
}

spgpuStatus_t spgpuMgVectorGet(spgpuMgVector_t v, void* globalValues)
{
	return vector_copy(v, globalValues, 0);
}

void* spgpuMgVectorLocal(spgpuMgVector_t v, int rank)
{
	return owned_ptr(v, rank);
}

/* ---- exchanges that are not inside the SpMV kernel ------------------------------------------------ */

/* EVENTS mode: every rank pushes its boundary entries into the neighbours' (even) zones; the products then wait
 * for the neighbours' pushes.  A push waits until the neighbours have finished the product that read the zones. */
static void exchange_with_events(spgpuMgHandle_t mg, spgpuMgVector_t x)
{
	spgpuMgMatrix_t A = x->A;
	const size_t es = A->esize, w = (size_t)A->halo;
	int r;
	for (r = 0; r < mg->world; ++r) {
		const size_t n = (size_t)(A->hi[r] - A->lo[r]);
		cudaStream_t s = rank_stream(mg, r);
		use_rank(mg, r);
		if (r > 0 && mg->doneValid[r - 1]) cudaStreamWaitEvent(s, mg->evDone[r - 1], 0);
		if (r < mg->world - 1 && mg->doneValid[r + 1]) cudaStreamWaitEvent(s, mg->evDone[r + 1], 0);
		if (r > 0) {
			const size_t nb = (size_t)(A->hi[r - 1] - A->lo[r - 1]);
			void* dst = (char*)x->ext[r - 1] + (w + nb) * es;
			if (mg->direct[r][r - 1])
				spgpuHaloPush(mg->h[r], dst, (char*)x->ext[r] + w * es, w * es, NULL, 0);
			else
				cudaMemcpyPeerAsync(dst, mg->dev[r - 1], (char*)x->ext[r] + w * es, mg->dev[r], w * es, s);
		}
		if (r < mg->world - 1) {
			if (mg->direct[r][r + 1])
				spgpuHaloPush(mg->h[r], x->ext[r + 1], (char*)x->ext[r] + n * es, w * es, NULL, 0);
			else
				cudaMemcpyPeerAsync(x->ext[r + 1], mg->dev[r + 1], (char*)x->ext[r] + n * es, mg->dev[r], w * es, s);
		}
		cudaEventRecord(mg->evPush[r], s);
	}
	for (r = 0; r < mg->world; ++r) {
		cudaStream_t s = rank_stream(mg, r);
		use_rank(mg, r);
		if (r > 0) cudaStreamWaitEvent(s, mg->evPush[r - 1], 0);
		if (r < mg->world - 1) cudaStreamWaitEvent(s, mg->evPush[r + 1], 0);
	}
}

/* all-gather mode: every rank copies its owned entries into every rank's full-length x over NVLink */
static spgpuStatus_t allgather(spgpuMgHandle_t mg, spgpuMgVector_t x)
{
	spgpuMgMatrix_t A = x->A;
	const size_t es = A->esize;
	int r, q;
	for (r = 0; r < mg->world; ++r)
		if (!x->full[r]) {
			use_rank(mg, r);
			if (cudaMalloc(&x->full[r], (size_t)(A->rows > 0 ? A->rows : 1) * es) != cudaSuccess)
				return SPGPU_OUTOFMEMORY;
		}
	for (r = 0; r < mg->world; ++r) {
		const size_t bytes = (size_t)(A->hi[r] - A->lo[r]) * es;
		cudaStream_t s = rank_stream(mg, r);
		use_rank(mg, r);
		for (q = 0; q < mg->world; ++q)
			if (mg->doneValid[q] && q != r)
				cudaStreamWaitEvent(s, mg->evDone[q], 0);
		for (q = 0; q < mg->world && bytes; ++q) {
			void* dst = (char*)x->full[q] + (size_t)A->lo[r] * es;
			if (mg->dev[q] == mg->dev[r])
				cudaMemcpyAsync(dst, owned_ptr(x, r), bytes, cudaMemcpyDeviceToDevice, s);
			else
				cudaMemcpyPeerAsync(dst, mg->dev[q], owned_ptr(x, r), mg->dev[r], bytes, s);
		}
		cudaEventRecord(mg->evPush[r], s);
	}
	for (r = 0; r < mg->world; ++r) {
		cudaStream_t s = rank_stream(mg, r);
		use_rank(mg, r);
		for (q = 0; q < mg->world; ++q)
			if (q != r)
				cudaStreamWaitEvent(s, mg->evPush[q], 0);
	}
	return SPGPU_SUCCESS;
}

static void mark_products_done(spgpuMgHandle_t mg)
{
	int r;
	for (r = 0; r < mg->world; ++r) {
		use_rank(mg, r);
		cudaEventRecord(mg->evDone[r], rank_stream(mg, r));
		mg->doneValid[r] = 1;
	}
}

static spgpuStatus_t check_launches(void)
{
	return cudaGetLastError() == cudaSuccess ? SPGPU_SUCCESS : SPGPU_UNSPECIFIED;
}

/* ---- typed operations ----------------------------------------------------------------------------- */

#define MG_ROWS(A, r) ((A)->hi[r] - (A)->lo[r])

#define SPGPU_DEFINE_MG(S, T, R, TYPECODE, IS_NONZERO, ADD, SQRT)                                        \
	spgpuStatus_t spgpuMg##S##hellCreate(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA, const T* cM,            \
		const int* rP, int hackSize, const int* hackOffsets, const int* rS, int avgNnzPerRow, int rows,   \
		int cols, int baseIndex)                                                                          \
	{                                                                                                     \
		return hell_create(mg, pA, TYPECODE, cM, rP, hackSize, hackOffsets, rS, avgNnzPerRow, rows, cols, \
			baseIndex);                                                                                   \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##hdiaCreate(spgpuMgHandle_t mg, spgpuMgMatrix_t* pA, const T* dM,            \
		const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols)                     \
	{                                                                                                     \
		return hdia_create(mg, pA, TYPECODE, dM, offsets, hackSize, hackOffsets, rows, cols);             \
	}                                                                                                     \
	/* one rank's product on a filled x_ext, with the plain entry point of the matrix's format */         \
	static void mg_##S##_block_spmv(spgpuMgMatrix_t A, int r, T* z, const T* y, T alpha, const T* x, T beta) \
	{                                                                                                     \
		if (A->kind == MG_KIND_HDIA)                                                                      \
			spgpu##S##hdiaspmv(A->mg->h[r], z, y, alpha, (const T*)A->cM[r], A->rP[r], A->hackSize,       \
				A->hackOffsets[r], MG_ROWS(A, r), MG_ROWS(A, r) + 2 * A->halo, x, beta);                  \
		else                                                                                              \
			spgpu##S##hellspmv(A->mg->h[r], z, y, alpha, (const T*)A->cM[r], A->rP[r], A->hackSize,       \
				A->hackOffsets[r], A->rS[r], NULL, A->avg, MG_ROWS(A, r), x, beta, A->baseIndex);         \
	}                                                                                                     \
	static spgpuStatus_t mg_##S##_spmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y, T alpha, \
		spgpuMgMatrix_t A, spgpuMgVector_t x, T beta)                                                     \
	{                                                                                                     \
		int previous = 0, r;                                                                              \
		spgpuStatus_t st = SPGPU_SUCCESS;                                                                 \
		if (!mg || !A || !x || !z || A->type != TYPECODE || x->A != A || z->A != A || (y && y->A != A)    \
				|| z == x || (!y && IS_NONZERO(beta)))                                                    \
			return SPGPU_UNSUPPORTED;                                                                     \
		cudaGetDevice(&previous);                                                                         \
		if (A->mode == MG_MODE_ALLGATHER && mg->world > 1) {                                              \
			st = allgather(mg, x);                                                                        \
			for (r = 0; r < mg->world && st == SPGPU_SUCCESS; ++r) {                                      \
				use_rank(mg, r);                                                                          \
				mg_##S##_block_spmv(A, r, (T*)owned_ptr(z, r), y ? (const T*)owned_ptr(y, r) : NULL,      \
					alpha, (const T*)x->full[r], beta);                                                   \
			}                                                                                             \
			mark_products_done(mg);                                                                       \
		} else if (mg->exchange == SPGPU_MG_FUSED || mg->world == 1 || A->halo == 0) {                    \
			const unsigned seq = ++mg->haloSeq;                                                           \
			for (r = 0; r < mg->world; ++r) {                                                             \
				const spgpuHaloLinks* links = mg->world > 1 ? &x->links[r] : NULL;                        \
				T* zr = (T*)owned_ptr(z, r);                                                              \
				const T* yr = y ? (const T*)owned_ptr(y, r) : NULL;                                       \
				use_rank(mg, r);                                                                          \
				if (A->kind == MG_KIND_HDIA)                                                              \
					spgpu##S##hdiaspmvHalo(mg->h[r], zr, yr, alpha, (const T*)A->cM[r], A->rP[r],         \
						A->hackSize, A->hackOffsets[r], MG_ROWS(A, r), MG_ROWS(A, r) + 2 * A->halo,       \
						(T*)x->ext[r], beta, A->halo, links, seq);                                        \
				else                                                                                      \
					spgpu##S##hellspmvHalo(mg->h[r], zr, yr, alpha, (const T*)A->cM[r], A->rP[r],         \
						A->hackSize, A->hackOffsets[r], A->rS[r], A->avg, MG_ROWS(A, r), (T*)x->ext[r],   \
						beta, A->baseIndex, A->halo, links, seq);                                         \
			}                                                                                             \
		} else {                                                                                          \
			exchange_with_events(mg, x);                                                                  \
			for (r = 0; r < mg->world; ++r) {                                                             \
				use_rank(mg, r);                                                                          \
				mg_##S##_block_spmv(A, r, (T*)owned_ptr(z, r), y ? (const T*)owned_ptr(y, r) : NULL,      \
					alpha, (const T*)x->ext[r], beta);                                                    \
			}                                                                                             \
			mark_products_done(mg);                                                                       \
		}                                                                                                 \
		if (st == SPGPU_SUCCESS)                                                                          \
			st = check_launches();                                                                        \
		cudaSetDevice(previous);                                                                          \
		return st;                                                                                        \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##hellspmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y, T alpha, \
		spgpuMgMatrix_t A, spgpuMgVector_t x, T beta)                                                     \
	{                                                                                                     \
		return (A && A->kind == MG_KIND_HELL) ? mg_##S##_spmv(mg, z, y, alpha, A, x, beta) : SPGPU_UNSUPPORTED; \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##hdiaspmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y, T alpha, \
		spgpuMgMatrix_t A, spgpuMgVector_t x, T beta)                                                     \
	{                                                                                                     \
		return (A && A->kind == MG_KIND_HDIA) ? mg_##S##_spmv(mg, z, y, alpha, A, x, beta) : SPGPU_UNSUPPORTED; \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##spmv(spgpuMgHandle_t mg, spgpuMgVector_t z, spgpuMgVector_t y, T alpha,     \
		spgpuMgMatrix_t A, spgpuMgVector_t x, T beta)                                                     \
	{                                                                                                     \
		return mg_##S##_spmv(mg, z, y, alpha, A, x, beta);                                                \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##dot(spgpuMgHandle_t mg, T* result, spgpuMgVector_t a, spgpuMgVector_t b)    \
	{                                                                                                     \
		int previous = 0, r, bad = 0;                                                                     \
		T total;                                                                                          \
		memset(&total, 0, sizeof(total));                                                                 \
		if (!mg || !a || !b || a->A != b->A || a->A->type != TYPECODE)                                    \
			return SPGPU_UNSUPPORTED;                                                                     \
		cudaGetDevice(&previous);                                                                         \
		for (r = 0; r < mg->world; ++r) {                                                                 \
			use_rank(mg, r);                                                                              \
			spgpu##S##dotDev(mg->h[r], MG_ROWS(a->A, r), (const T*)owned_ptr(a, r),                       \
				(const T*)owned_ptr(b, r), (T*)mg->dScalar[r]);                                           \
			if (cudaMemcpyAsync(mg->hostSlots + 64 * r, mg->dScalar[r], sizeof(T), cudaMemcpyDeviceToHost, \
					rank_stream(mg, r)) != cudaSuccess)                                                   \
				bad = 1;                                                                                  \
		}                                                                                                 \
		for (r = 0; r < mg->world; ++r) {                                                                 \
			use_rank(mg, r);                                                                              \
			if (cudaStreamSynchronize(rank_stream(mg, r)) != cudaSuccess)                                 \
				bad = 1;                                                                                  \
		}                                                                                                 \
		for (r = 0; r < mg->world; ++r) {            /* rank order: the same bits every run */            \
			T part;                                                                                       \
			memcpy(&part, mg->hostSlots + 64 * r, sizeof(T));                                             \
			total = ADD(total, part);                                                                     \
		}                                                                                                 \
		*result = total;                                                                                  \
		cudaSetDevice(previous);                                                                          \
		return bad ? SPGPU_UNSPECIFIED : SPGPU_SUCCESS;                                                   \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##nrm2(spgpuMgHandle_t mg, R* result, spgpuMgVector_t x)                      \
	{                                                                                                     \
		int previous = 0, r, bad = 0;                                                                     \
		R total = 0;                                                                                      \
		if (!mg || !x || x->A->type != TYPECODE)                                                          \
			return SPGPU_UNSUPPORTED;                                                                     \
		cudaGetDevice(&previous);                                                                         \
		for (r = 0; r < mg->world; ++r) {                                                                 \
			use_rank(mg, r);                                                                              \
			spgpu##S##nrm2sqDev(mg->h[r], MG_ROWS(x->A, r), (const T*)owned_ptr(x, r), (R*)mg->dScalar[r]); \
			if (cudaMemcpyAsync(mg->hostSlots + 64 * r, mg->dScalar[r], sizeof(R), cudaMemcpyDeviceToHost, \
					rank_stream(mg, r)) != cudaSuccess)                                                   \
				bad = 1;                                                                                  \
		}                                                                                                 \
		for (r = 0; r < mg->world; ++r) {                                                                 \
			use_rank(mg, r);                                                                              \
			if (cudaStreamSynchronize(rank_stream(mg, r)) != cudaSuccess)                                 \
				bad = 1;                                                                                  \
		}                                                                                                 \
		for (r = 0; r < mg->world; ++r) {                                                                 \
			R part;                                                                                       \
			memcpy(&part, mg->hostSlots + 64 * r, sizeof(R));                                             \
			total += part;                                                                                \
		}                                                                                                 \
		*result = SQRT(total);                                                                            \
		cudaSetDevice(previous);                                                                          \
		return bad ? SPGPU_UNSPECIFIED : SPGPU_SUCCESS;                                                   \
	}                                                                                                     \
	spgpuStatus_t spgpuMg##S##axpby(spgpuMgHandle_t mg, spgpuMgVector_t z, T beta, spgpuMgVector_t y,     \
		T alpha, spgpuMgVector_t x)                                                                       \
	{                                                                                                     \
		int previous = 0, r;                                                                              \
		spgpuStatus_t st;                                                                                 \
		if (!mg || !z || !y || !x || z->A != y->A || z->A != x->A || z->A->type != TYPECODE)              \
			return SPGPU_UNSUPPORTED;                                                                     \
		cudaGetDevice(&previous);                                                                         \
		for (r = 0; r < mg->world; ++r) {                                                                 \
			use_rank(mg, r);                                                                              \
			spgpu##S##axpby(mg->h[r], (T*)owned_ptr(z, r), MG_ROWS(z->A, r), beta, (T*)owned_ptr(y, r),   \
				alpha, (T*)owned_ptr(x, r));                                                              \
		}                                                                                                 \
		st = check_launches();                                                                            \
		cudaSetDevice(previous);                                                                          \
		return st;                                                                                        \
	}

#define NZ_REAL(v) ((v) != 0)
#define NZ_CPLX(v) ((v).x != 0 || (v).y != 0)
#define ADD_REAL(a, b) ((a) + (b))
SPGPU_DEFINE_MG(S, float, float, SPGPU_TYPE_FLOAT, NZ_REAL, ADD_REAL, sqrtf)
SPGPU_DEFINE_MG(D, double, double, SPGPU_TYPE_DOUBLE, NZ_REAL, ADD_REAL, sqrt)
SPGPU_DEFINE_MG(C, cuFloatComplex, float, SPGPU_TYPE_COMPLEX_FLOAT, NZ_CPLX, cuCaddf, sqrtf)
SPGPU_DEFINE_MG(Z, cuDoubleComplex, double, SPGPU_TYPE_COMPLEX_DOUBLE, NZ_CPLX, cuCadd, sqrt)

/* ---- conjugate gradients (double) --------------------------------------------------------------------- */

spgpuStatus_t spgpuMgDcgCreate(spgpuMgMatrix_t A, spgpuMgCg_t* pCg)
{
	spgpuMgCg_t cg;
	int previous = 0, r;
	spgpuStatus_t st;
	*pCg = NULL;
	if (!A || A->type != SPGPU_TYPE_DOUBLE)
		return SPGPU_UNSUPPORTED;
	cg = (spgpuMgCg_t)calloc(1, sizeof(*cg));
	if (!cg)
		return SPGPU_OUTOFMEMORY;
	cg->A = A;
	st = spgpuMgVectorCreate(A, &cg->x);
	if (st == SPGPU_SUCCESS) st = spgpuMgVectorCreate(A, &cg->r);
	if (st == SPGPU_SUCCESS) st = spgpuMgVectorCreate(A, &cg->p);
	if (st == SPGPU_SUCCESS) st = spgpuMgVectorCreate(A, &cg->ap);
	cudaGetDevice(&previous);
	for (r = 0; r < A->mg->world && st == SPGPU_SUCCESS; ++r) {
		use_rank(A->mg, r);
		if (cudaMalloc((void**)&cg->s[r], 4 * sizeof(double)) != cudaSuccess || cudaMemset(cg->s[r], 0, 4 * sizeof(double)) != cudaSuccess)
			st = SPGPU_OUTOFMEMORY;
	}
	cudaSetDevice(previous);
	if (st != SPGPU_SUCCESS) {
		spgpuMgDcgDestroy(cg);
		return st;
	}
	*pCg = cg;
	return SPGPU_SUCCESS;
}

void spgpuMgDcgDestroy(spgpuMgCg_t cg)
{
	int previous = 0, r;
	if (!cg)
		return;
	spgpuMgVectorDestroy(cg->x);
	spgpuMgVectorDestroy(cg->r);
	spgpuMgVectorDestroy(cg->p);
	spgpuMgVectorDestroy(cg->ap);
	cudaGetDevice(&previous);
	for (r = 0; r < cg->A->mg->world; ++r)
		if (cg->s[r]) {
			use_rank(cg->A->mg, r);
			cudaFree(cg->s[r]);
		}
	cudaSetDevice(previous);
	free(cg);
}

spgpuMgVector_t spgpuMgDcgSolution(spgpuMgCg_t cg)
{
	return cg ? cg->x : NULL;
}

static int cg_on_device(spgpuMgCg_t cg)
{
	spgpuMgHandle_t mg = cg->A->mg;
	return cg->A->mode == MG_MODE_HALO && (mg->exchange == SPGPU_MG_FUSED || mg->world == 1);
}

spgpuStatus_t spgpuMgDcgStart(spgpuMgCg_t cg, spgpuMgVector_t b, double* rr0)
{
	spgpuMgMatrix_t A;
	spgpuMgHandle_t mg;
	int previous = 0, r;
	spgpuStatus_t st;
	if (!cg || !b || b->A != cg->A)
		return SPGPU_UNSUPPORTED;
	A = cg->A;
	mg = A->mg;
	cudaGetDevice(&previous);
	for (r = 0; r < mg->world; ++r) {
		const size_t bytes = (size_t)MG_ROWS(A, r) * sizeof(double);
		cudaStream_t s = rank_stream(mg, r);
		use_rank(mg, r);
		cudaMemsetAsync(owned_ptr(cg->x, r), 0, bytes, s);
		cudaMemcpyAsync(owned_ptr(cg->r, r), owned_ptr(b, r), bytes, cudaMemcpyDeviceToDevice, s);
		cudaMemcpyAsync(owned_ptr(cg->p, r), owned_ptr(b, r), bytes, cudaMemcpyDeviceToDevice, s);
	}
	st = spgpuMgDdot(mg, &cg->rr, cg->r, cg->r);
	cg->iter = 0;
	for (r = 0; r < mg->world && st == SPGPU_SUCCESS; ++r) {
		use_rank(mg, r);
		if (cudaMemcpy(cg->s[r], &cg->rr, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
			st = SPGPU_UNSPECIFIED;
	}
	if (rr0)
		*rr0 = cg->rr;
	cudaSetDevice(previous);
	return st;
}

spgpuStatus_t spgpuMgDcgStep(spgpuMgCg_t cg, int iterations, double* rr)
{
	spgpuMgMatrix_t A;
	spgpuMgHandle_t mg;
	int previous = 0, r, it;
	spgpuStatus_t st = SPGPU_SUCCESS;
	if (!cg)
		return SPGPU_UNSUPPORTED;
	A = cg->A;
	mg = A->mg;
	cudaGetDevice(&previous);
	if (cg_on_device(cg)) {
		void* tables[MG_MAX_RANKS];
		for (r = 0; r < mg->world; ++r)
			tables[r] = mg->arTable[r];
		for (it = 0; it < iterations; ++it) {
			const unsigned seq = ++mg->haloSeq;
			const int cur = 2 * (int)(cg->iter & 1u), nxt = 2 - cur;   /* "rr <- rr'" is a swap of roles, not a kernel */
			spgpuPeerAllreduce ar;
			ar.world = mg->world;
			ar.tables = tables;
			ar.seq = ++mg->arSeq;
			for (r = 0; r < mg->world; ++r) {            /* Ap = A p (+ halo of p) and p.Ap, all-reduced */
				use_rank(mg, r);
				ar.myRank = r;
				if (A->kind == MG_KIND_HDIA)
					spgpuDhdiaspmvHaloDot(mg->h[r], (double*)owned_ptr(cg->ap, r), (const double*)A->cM[r], A->rP[r],
						A->hackSize, A->hackOffsets[r], MG_ROWS(A, r), MG_ROWS(A, r) + 2 * A->halo, (double*)cg->p->ext[r],
						A->halo, mg->world > 1 ? &cg->p->links[r] : NULL, seq, cg->s[r] + 1, mg->world > 1 ? &ar : NULL);
				else
					spgpuDhellspmvHaloDot(mg->h[r], (double*)owned_ptr(cg->ap, r), (const double*)A->cM[r], A->rP[r],
						A->hackSize, A->hackOffsets[r], A->rS[r], A->avg, MG_ROWS(A, r), (double*)cg->p->ext[r],
						A->baseIndex, A->halo, mg->world > 1 ? &cg->p->links[r] : NULL, seq, cg->s[r] + 1,
						mg->world > 1 ? &ar : NULL);
			}
			ar.seq = ++mg->arSeq;
			for (r = 0; r < mg->world; ++r) {            /* x += a p ; r -= a Ap ; rr' = r.r, all-reduced */
				use_rank(mg, r);
				ar.myRank = r;
				spgpuDcgUpdateDev(mg->h[r], (double*)owned_ptr(cg->x, r), (double*)owned_ptr(cg->r, r),
					(const double*)owned_ptr(cg->p, r), (const double*)owned_ptr(cg->ap, r), MG_ROWS(A, r),
					cg->s[r] + cur, cg->s[r] + 1, cg->s[r] + nxt, mg->world > 1 ? &ar : NULL);
			}
			for (r = 0; r < mg->world; ++r) {            /* p = r + (rr'/rr) p */
				use_rank(mg, r);
				spgpuDaxpbyDev(mg->h[r], (double*)owned_ptr(cg->p, r), MG_ROWS(A, r), cg->s[r] + nxt, cg->s[r] + cur, 1.0,
					(const double*)owned_ptr(cg->p, r), NULL, NULL, 1.0, (const double*)owned_ptr(cg->r, r));
			}
			++cg->iter;
		}
		st = check_launches();
		if (rr && st == SPGPU_SUCCESS) {
			use_rank(mg, 0);
			if (cudaMemcpyAsync(mg->hostSlots, cg->s[0] + 2 * (cg->iter & 1u), sizeof(double), cudaMemcpyDeviceToHost,
					rank_stream(mg, 0)) != cudaSuccess)
				st = SPGPU_UNSPECIFIED;
			if (st == SPGPU_SUCCESS)
				st = spgpuMgSynchronize(mg);
			memcpy(&cg->rr, mg->hostSlots, sizeof(double));
			*rr = cg->rr;
		}
	} else {
		for (it = 0; it < iterations && st == SPGPU_SUCCESS; ++it) {
			double pap = 0.0, rrNew = 0.0, a;
			st = spgpuMgDspmv(mg, cg->ap, NULL, 1.0, A, cg->p, 0.0);
			if (st == SPGPU_SUCCESS) st = spgpuMgDdot(mg, &pap, cg->p, cg->ap);
			a = cg->rr / pap;
			if (st == SPGPU_SUCCESS) st = spgpuMgDaxpby(mg, cg->x, 1.0, cg->x, a, cg->p);
			if (st == SPGPU_SUCCESS) st = spgpuMgDaxpby(mg, cg->r, 1.0, cg->r, -a, cg->ap);
			if (st == SPGPU_SUCCESS) st = spgpuMgDdot(mg, &rrNew, cg->r, cg->r);
			if (st == SPGPU_SUCCESS) st = spgpuMgDaxpby(mg, cg->p, rrNew / cg->rr, cg->p, 1.0, cg->r);
			cg->rr = rrNew;
		}
		if (rr)
			*rr = cg->rr;
	}
	cudaSetDevice(previous);
	return st;
}
