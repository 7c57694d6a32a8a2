/*
 * spgpu.h -- umbrella header of the B200-native spGPU drop-in (libspgpu.so).
 *
 * This file is the C ABI of the SpMV hot path.  Every declaration below binds the
 * same symbol, with the same argument order and meaning, as the reference
 * library's headers (cited per block as <reference file>:<line>); the reference
 * has no single "spgpu.h", its ABI is the union of core.h, ell.h, hell.h, dia.h,
 * hdia.h, vector.h and the four *_conv.h headers.  Thin headers with those names
 * live in include/core/ and simply include this file, so a consumer that does
 * `#include "core/hell.h"` keeps compiling.
 *
 * Conventions (same as the reference):
 *   - all vector/matrix arguments are DEVICE pointers unless the name says host;
 *   - alpha/beta scalars are passed BY VALUE from the host (complex ones as the
 *     cuComplex.h structs);
 *   - sizes and pitches are `int`, pitches are in ELEMENTS;
 *   - compute calls are asynchronous on handle->currentStream and return void;
 *     the reductions (dot, nrm2, amax, asum) block and return the value;
 *   - the caller must have cudaSetDevice(handle->device) current.
 *
 * The declarations are generated per value type with the SPGPU_FOR_* macros so
 * that the four precisions cannot drift apart.
 */
#ifndef SPGPU_H_
#define SPGPU_H_

#include <stddef.h>
#include "driver_types.h"
#include "cuComplex.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Address-space markers, documentation only (reference core.h:37-40). */
#ifndef __host
#define __host
#endif
#ifndef __device
#define __device
#endif

/* ------------------------------------------------------------------------- */
/* Core: status codes, type codes, the handle   (reference core.h:43-138)     */
/* ------------------------------------------------------------------------- */

typedef int spgpuStatus_t;
#define SPGPU_SUCCESS      0
#define SPGPU_UNSUPPORTED  1
#define SPGPU_UNSPECIFIED  2
#define SPGPU_OUTOFMEMORY  3

typedef int spgpuType_t;
#define SPGPU_TYPE_INT             0
#define SPGPU_TYPE_FLOAT           1
#define SPGPU_TYPE_DOUBLE          2
#define SPGPU_TYPE_COMPLEX_FLOAT   3
#define SPGPU_TYPE_COMPLEX_DOUBLE  4

/*
 * Public part of a handle.  The struct is visible to callers in the reference
 * (core.h:60-82), so field order and types are ABI and are kept verbatim.  The
 * library allocates a larger private record whose first member is this struct
 * (per-handle reduction scratch lives behind it, see csrc/spgpu_internal.h).
 */
typedef struct spgpuHandleStruct {
	cudaStream_t currentStream;   /* stream every call on this handle launches on  */
	cudaStream_t defaultStream;   /* created by spgpuCreate, owned by the handle   */
	int device;
	int warpSize;
	int maxThreadsPerBlock;
	int maxGridSizeX;
	int maxGridSizeY;
	int maxGridSizeZ;
	int multiProcessorCount;
	int capabilityMajor;
	int capabilityMinor;
} SpgpuHandleStruct;

typedef const SpgpuHandleStruct* spgpuHandle_t;

/* core.h:94  -- SPGPU_SUCCESS, or SPGPU_UNSPECIFIED when the device query fails. */
spgpuStatus_t spgpuCreate(spgpuHandle_t* pHandle, int device);
/* core.h:101 */
void spgpuDestroy(spgpuHandle_t pHandle);
/* core.h:109 */
void spgpuStreamCreate(spgpuHandle_t pHandle, cudaStream_t* stream);
/* core.h:116 */
void spgpuStreamDestroy(cudaStream_t stream);
/* core.h:124 -- stream == 0 selects the handle's default stream again. */
void spgpuSetStream(spgpuHandle_t pHandle, cudaStream_t stream);
/* core.h:131 */
cudaStream_t spgpuGetStream(spgpuHandle_t pHandle);
/* core.h:138 -- bytes of one element of typeCode, 0 for an unknown code. */
size_t spgpuSizeOf(spgpuType_t typeCode);

/* core.h:151-154 */
#define cuFloatComplex_isZero(a)      (a.x == 0.0f && a.y == 0.0f)
#define cuDoubleComplex_isZero(a)     (a.x == 0.0 && a.y == 0.0)
#define cuFloatComplex_isNotZero(a)   (a.x != 0.0f || a.y != 0.0f)
#define cuDoubleComplex_isNotZero(a)  (a.x != 0.0 || a.y != 0.0)

/* Value types: X(symbol letter, element type, real type of norms). */
#define SPGPU_FOR_FLOAT_TYPES(X) \
	X(S, float, float)              \
	X(D, double, double)            \
	X(C, cuFloatComplex, float)     \
	X(Z, cuDoubleComplex, double)

/* ------------------------------------------------------------------------- */
/* SpMV:  z = alpha * A * x + beta * y                                        */
/* ------------------------------------------------------------------------- */

#define ELL_PITCH_ALIGN_BYTE   128   /* ell.h:24  */
#define HELL_PITCH_ALIGN_BYTE  128   /* hell.h:24 */
#define DIA_PITCH_ALIGN_BYTE   128   /* dia.h:24  */

/*
 * ELL (ell.h:46-173).  A is pitched column-major ELLPACK: slot k of row i is
 * cM[i + k*cMPitch] / rP[i + k*rPPitch] (column index, baseIndex-based).
 * rS[i] = slots used by row i, or NULL to use maxNnzPerRow for every row (the
 * padding must then hold value 0 with a valid index).  rIdx (or NULL) sends row
 * i's result to z[rIdx[i]].  beta == 0 means y is never read.  z may alias y.
 */
#define SPGPU_DECL_ELLSPMV(S, T, R)                                          \
	void spgpu##S##ellspmv(spgpuHandle_t handle, __device T* z,              \
		const __device T* y, T alpha, const __device T* cM,                  \
		const __device int* rP, int cMPitch, int rPPitch,                    \
		const __device int* rS, const __device int* rIdx, int avgNnzPerRow,  \
		int maxNnzPerRow, int rows, const __device T* x, T beta,             \
		int baseIndex);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_ELLSPMV)

/*
 * HELL (hell.h:45-169).  Rows are grouped in hacks of hackSize rows (a multiple
 * of 32); hack h is its own column-major ELL slab starting at ELEMENT offset
 * hackOffsets[h] (ceil(rows/hackSize) entries, no terminator): slot k of row i
 * is at hackOffsets[i/hackSize] + k*hackSize + i%hackSize.  rS is mandatory;
 * slots >= rS[i] are never read (their content is undefined).
 */
#define SPGPU_DECL_HELLSPMV(S, T, R)                                         \
	void spgpu##S##hellspmv(spgpuHandle_t handle, __device T* z,             \
		const __device T* y, T alpha, const __device T* cM,                  \
		const __device int* rP, int hackSize,                                \
		const __device int* hackOffsets, const __device int* rS,             \
		const __device int* rIdx, int avgNnzPerRow, int rows,                \
		const __device T* x, T beta, int baseIndex);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_HELLSPMV)

/*
 * DIA (dia.h:42-143).  Diagonal j holds A(i, i+offsets[j]) at dM[i + j*dMPitch];
 * cells whose column falls outside [0, cols) are skipped, all others are read.
 */
#define SPGPU_DECL_DIASPMV(S, T, R)                                          \
	void spgpu##S##diaspmv(spgpuHandle_t handle, __device T* z,              \
		const __device T* y, T alpha, const __device T* dM,                  \
		const __device int* offsets, int dMPitch, int rows, int cols,        \
		int diags, const __device T* x, T beta);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_DIASPMV)

/*
 * HDIA (hdia.h:38-142).  Hack h owns the diagonals [hackOffsets[h],
 * hackOffsets[h+1]) of offsets[] (hackOffsets has hacks+1 entries and counts
 * DIAGONALS, unlike HELL's); cell (d, r) is dM[d*hackSize + r].
 */
#define SPGPU_DECL_HDIASPMV(S, T, R)                                         \
	void spgpu##S##hdiaspmv(spgpuHandle_t handle, T* z, const T* y, T alpha, \
		const T* dM, const int* offsets, int hackSize,                       \
		const int* hackOffsets, int rows, int cols, const T* x, T beta);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_HDIASPMV)

/*
 * ELL coefficient update (ell.h:194-302).  Out of the SpMV path; exported for
 * link compatibility.  For each (aI, aJ, aVal) triple the slot of row aI whose
 * column equals aJ is overwritten with aVal (rows' columns must be sorted).
 */
#define SPGPU_DECL_ELLCSPUT(S, T, R)                                         \
	void spgpu##S##ellcsput(spgpuHandle_t handle, T alpha, __device T* cM,   \
		__device const int* rP, int cMPitch, int rPPitch,                    \
		__device const int* rS, int nnz, __device int* aI, __device int* aJ, \
		__device T* aVal, int baseIndex);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_ELLCSPUT)

/* ------------------------------------------------------------------------- */
/* BLAS-1 companions (vector.h)                                               */
/* ------------------------------------------------------------------------- */

/*
 * Per value type, lines of reference vector.h for S / D / C / Z:
 *   dot    69/366/630/925   sum a_i*b_i, NOT conjugated for C/Z
 *   mdot   85/397/646/941   `count` dots over vectors `pitch` elements apart
 *   nrm2   117/414/678/972  sqrt(sum |x_i|^2), real result
 *   mnrm2  131/429/692/987
 *   scal   148/351/709/910  y = alpha*x
 *   axpby  165/447/726/1005 z = beta*y + alpha*x   (note the argument order)
 *   maxpby 187/469/748/1027
 *   abs    103/380/664/958  y = alpha*|x|
 *   axy    206/488/767/1046 z = alpha*x*y (element-wise)
 *   axypbz 225/506/786/1064 w = beta*z + alpha*x*y
 *   maxy, maxypbz           multi-vector forms of the two above
 *   asum   319/599/879/1158 sum |x_i|;  amax 323/603/883/1162  max |x_i|
 *   masum, mamax
 * In-place use (output aliasing an input exactly) is allowed everywhere.
 */
#define SPGPU_DECL_BLAS1(S, T, R)                                            \
	T    spgpu##S##dot(spgpuHandle_t handle, int n, __device T* a,           \
		__device T* b);                                                      \
	void spgpu##S##mdot(spgpuHandle_t handle, T* y, int n, __device T* a,    \
		__device T* b, int count, int pitch);                                \
	R    spgpu##S##nrm2(spgpuHandle_t handle, int n, __device T* x);         \
	void spgpu##S##mnrm2(spgpuHandle_t handle, R* y, int n, __device T* x,   \
		int count, int pitch);                                               \
	void spgpu##S##scal(spgpuHandle_t handle, __device T* y, int n, T alpha, \
		__device T* x);                                                      \
	void spgpu##S##axpby(spgpuHandle_t handle, __device T* z, int n, T beta, \
		__device T* y, T alpha, __device T* x);                              \
	void spgpu##S##maxpby(spgpuHandle_t handle, __device T* z, int n,        \
		T beta, __device T* y, T alpha, __device T* x, int count,            \
		int pitch);                                                          \
	void spgpu##S##abs(spgpuHandle_t handle, __device T* y, int n, T alpha,  \
		__device T* x);                                                      \
	void spgpu##S##axy(spgpuHandle_t handle, __device T* z, int n, T alpha,  \
		__device T* x, __device T* y);                                       \
	void spgpu##S##axypbz(spgpuHandle_t handle, __device T* w, int n,        \
		T beta, __device T* z, T alpha, __device T* x, __device T* y);       \
	void spgpu##S##maxy(spgpuHandle_t handle, __device T* z, int n, T alpha, \
		__device T* x, __device T* y, int count, int pitch);                 \
	void spgpu##S##maxypbz(spgpuHandle_t handle, __device T* w, int n,       \
		T beta, __device T* z, T alpha, __device T* x, __device T* y,        \
		int count, int pitch);                                               \
	R    spgpu##S##asum(spgpuHandle_t handle, int n, T* x);                  \
	R    spgpu##S##amax(spgpuHandle_t handle, int n, T* x);                  \
	void spgpu##S##masum(spgpuHandle_t handle, R* y, int n, T* x, int count, \
		int pitch);                                                          \
	void spgpu##S##mamax(spgpuHandle_t handle, R* y, int n, T* x, int count, \
		int pitch);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_BLAS1)

/*
 * Sparse-vector companions, also for int (vector.h: gath 30/291/572/852/1130,
 * scat 50/311/592/872/1150, setscal 1182-1215):
 *   gath:    xValues[i] = y[xIndices[i] - xBaseIndex]
 *   scat:    y[p] = beta != 0 ? beta*y[p] + xValues[i] : xValues[i],
 *            p = xIndices[i] - xBaseIndex
 *   entries with p < 0 are skipped by both.
 *   setscal: y[first-baseIndex .. last-baseIndex] = val
 */
#define SPGPU_FOR_ALL_TYPES(X) \
	X(I, int)                  \
	X(S, float)                \
	X(D, double)               \
	X(C, cuFloatComplex)       \
	X(Z, cuDoubleComplex)

#define SPGPU_DECL_SPVEC(S, T)                                               \
	void spgpu##S##gath(spgpuHandle_t handle, __device T* xValues, int xNnz, \
		const __device int* xIndices, int xBaseIndex, const __device T* y);  \
	void spgpu##S##scat(spgpuHandle_t handle, __device T* y, int xNnz,       \
		const __device T* xValues, const __device int* xIndices,             \
		int xBaseIndex, T beta);                                             \
	void spgpu##S##setscal(spgpuHandle_t handle, int first, int last,        \
		int baseIndex, T val, __device T* y);
SPGPU_FOR_ALL_TYPES(SPGPU_DECL_SPVEC)

/* ------------------------------------------------------------------------- */
/* Host-side format construction (all pointers are HOST memory).              */
/* The integer metadata these produce is bit-exact with the reference.        */
/* ------------------------------------------------------------------------- */

/* ell_conv.h:30 (the misspelling is the reference's symbol name). */
void computeEllRowLenghts(int* ellRowLengths, int* ellMaxRowSize, int rowsCount,
	int nonZerosCount, const int* cooRowIndices, int cooBaseIndex);
/* ell_conv.h:44 -- rowsCount rounded up to a multiple of 32 elements. */
int computeEllAllocPitch(int rowsCount);
/* ell_conv.h:64 -- slots of a row fill in COO order; caller zero-fills first. */
void cooToEll(void* ellValues, int* ellIndices, int ellValuesPitch,
	int ellIndicesPitch, int ellMaxRowSize, int ellBaseIndex, int rowsCount,
	int nonZerosCount, const int* cooRowIndices, const int* cooColsIndices,
	const void* cooValues, int cooBaseIndex, spgpuType_t valuesType);
/* ell_conv.h:80 -- rows sorted by descending length; rIdx[i] = source row. */
void ellToOell(int* rIdx, void* dstEllValues, int* dstEllIndices, int* dstRs,
	const void* srcEllValues, const int* srcEllIndices, const int* srcRs,
	int ellValuesPitch, int ellIndicesPitch, int rowsCount,
	spgpuType_t valuesType);

/* hell_conv.h:29 -- allocation is allocationHeight*hackSize elements. */
void computeHellAllocSize(int* allocationHeight, int hackSize, int rowsCount,
	const int* ellRowLengths);
/* hell_conv.h:49 */
void ellToHell(void* hellValues, int* hellIndices, int* hackOffsets,
	int hackSize, const void* ellValues, const int* ellIndices,
	int ellValuesPitch, int ellIndicesPitch, int* ellRowLengths, int rowsCount,
	spgpuType_t valuesType);

/* dia_conv.h:20 */
int computeDiaDiagonalsCount(int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices);
/* dia_conv.h:28 -- offsets come out ascending; caller zero-fills values. */
void coo2dia(void* values, int* offsets, int valuesPitch, int diagonals,
	int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, const void* cooValues,
	int cooBaseIndex, spgpuType_t valuesType);
/* dia_conv.h:43 */
int computeDiaAllocPitch(int rowsCount);

/* hdia_conv.h:20 */
int getHdiaHacksCount(int hackSize, int rowsCount);
/* hdia_conv.h:22 -- hackOffsets gets hacks+1 entries (prefix of diagonal counts). */
void computeHdiaHackOffsets(int* allocationHeight, int* hackOffsets,
	int hackSize, const void* diaValues, int diaValuesPitch, int diagonals,
	int rowsCount, spgpuType_t valuesType);
/* hdia_conv.h:32 */
void diaToHdia(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, const void* diaValues, const int* diaOffsets,
	int diaValuesPitch, int diagonals, int rowsCount, spgpuType_t valuesType);
/* hdia_conv.h:45 */
void computeHdiaHackOffsetsFromCoo(int* allocationHeight, int* hackOffsets,
	int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, int cooBaseIndex);
/* hdia_conv.h:57 -- caller zero-fills hdiaValues first. */
void cooToHdia(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, const void* cooValues,
	int cooBaseIndex, spgpuType_t valuesType);

/* hdia_conv.h:72 -- cooToHdia over blocks: every COO value is a block of blockSize elements. */
void bcooToBhdia(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, const void* cooValues,
	int cooBaseIndex, spgpuType_t valuesType, int blockSize);

/* coo_conv.h:20 -- number of non-zero blockRows x blockCols blocks of a 0-based COO matrix. */
int computeBcooSize(int blockRows, int blockCols, const int* rows, const int* cols, int nonZeros);
/* coo_conv.h:23 -- blocks in first-seen order, column-major inside a block, zero-filled. */
void cooToBcoo(int* bRows, int* bCols, void* blockValues, int blockRows, int blockCols,
	const int* rows, const int* cols, const void* values, int nonZeros, spgpuType_t valuesType);

#ifdef __cplusplus
}
#endif

#endif /* SPGPU_H_ */
