/*
 * MatrixMarket input for the SpMV path: what the reference's perf drivers
 * do before they convert and multiply -- read a coordinate .mtx file into 0-based COO arrays
 * and unfold a symmetric matrix (reference src/utils/mmread.hpp:36-99, mmread.cpp:16-216,
 * mmutils.hpp:11-62; the banner rules are those of the NIST mmio.c the reference vendors
 * under src/external/).  The reference keeps this in a C++ utility next to its drivers, not
 * in libspgpu; here it is plain C with path arguments instead of FILE*, so a binding can
 * call it.  Status and storage / symmetry codes are the reference's.
 */
#ifndef SPGPU_MM_H_
#define SPGPU_MM_H_

#include "spgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference mmread.hpp:12-33 */
#define MATRIX_READ_SUCCESS        0
#define MATRIX_READ_UNSUPPORTED    1
#define MATRIX_READ_INVALID_INPUT  2

#define MATRIX_STORAGE_INTEGER     0
#define MATRIX_STORAGE_REAL        1
#define MATRIX_STORAGE_COMPLEX     2
#define MATRIX_STORAGE_PATTERN     3

#define MATRIX_TYPE_GENERAL        0
#define MATRIX_TYPE_SYMMETRIC      1
#define MATRIX_TYPE_SKEW           2
#define MATRIX_TYPE_HERMITIAN      3

typedef struct spgpuMmProperties {
	int rowsCount;
	int columnsCount;
	int nonZerosCount;      /* entries stored in the file (before any unfolding) */
	int isStoredSparse;     /* coordinate format */
	int matrixStorage;      /* MATRIX_STORAGE_* */
	int matrixType;         /* MATRIX_TYPE_*    */
} spgpuMmProperties;

/* loadMmProperties (mmread.cpp:16-62): banner + size line.  Returns 1 when the file is a valid
 * MatrixMarket matrix, 0 otherwise (the reference returns bool). */
int spgpuMmLoadProperties(const char* path, spgpuMmProperties* props);

/*
 * loadMmMatrixToCoo (mmread.cpp:137-216): the nonZerosCount entries of a coordinate file as
 * 0-based COO in file order.  valuesType selects the overload: SPGPU_TYPE_FLOAT accepts real
 * and integer storage, SPGPU_TYPE_DOUBLE only real, SPGPU_TYPE_INT only integer (the
 * reference's rules, kept as they are); values == NULL reads a pattern file (indices only).
 * Returns MATRIX_READ_*; a file with fewer entries than announced is INVALID_INPUT (the
 * reference prints a message and returns success with the tail uninitialised).
 */
int spgpuMmLoadMatrixToCoo(const char* path, void* values, int* rowIndices, int* columnIndices,
	spgpuType_t valuesType);

/* getUnfoldedMmSymmetricSize / unfoldMmSymmetricReal (mmutils.hpp:11-62): entries with value 0
 * are dropped, diagonal entries kept once, every other entry followed by its transpose. */
int spgpuMmUnfoldedSymmetricSize(const void* values, const int* rows, const int* cols,
	int nonZerosCount, spgpuType_t valuesType);
void spgpuMmUnfoldSymmetric(int* unfoldedRows, int* unfoldedCols, void* unfoldedValues,
	const int* rows, const int* cols, const void* values, int nonZerosCount,
	spgpuType_t valuesType);

/* loadMmVectorToDenseVector (mmread.cpp:219-270) for an array-format file holding one vector;
 * float/double need real storage, int needs integer storage (else INVALID_INPUT). */
int spgpuMmLoadDenseVector(const char* path, void* values, int vectorSize, spgpuType_t valuesType);

#ifdef __cplusplus
}
#endif

#endif
