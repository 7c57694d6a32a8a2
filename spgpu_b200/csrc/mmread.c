/*
 * MatrixMarket coordinate reader + symmetric unfolding (include/spgpu_mm.h).
 *
 * Follows reference src/utils/mmread.cpp:16-270 and mmutils.hpp:11-62 for WHAT is read and
 * how it is typed, and the banner / size-line rules of the NIST mmio.c the reference vendors
 * (src/external/mmio.c: mm_read_banner, mm_is_valid, mm_read_mtx_crd_size).  Lines are
 * parsed with strtol / strtod instead of fscanf("%d %d %lg"): the same C library conversion,
 * so the same bits, without fscanf's per-call overhead on files with 10^8 entries.
 */
#define _POSIX_C_SOURCE 200809L   /* getline */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spgpu_mm.h"

#define MM_LINE 1025      /* MM_MAX_LINE_LENGTH of mmio.h */

static void lower(char* s)
{
	for (; *s; ++s)
		*s = (char)tolower((unsigned char)*s);
}

/* banner -> (sparse, storage, type); 0 on success */
static int read_banner(FILE* f, int* sparse, int* storage, int* type)
{
	char line[MM_LINE], banner[64], mtx[64], crd[64], data[64], scheme[64];
	if (!fgets(line, sizeof line, f))
		return -1;
	if (sscanf(line, "%63s %63s %63s %63s %63s", banner, mtx, crd, data, scheme) != 5)
		return -1;
	lower(mtx); lower(crd); lower(data); lower(scheme);
	if (strncmp(banner, "%%MatrixMarket", 14) != 0 || strcmp(mtx, "matrix") != 0)
		return -1;
	if (strcmp(crd, "coordinate") == 0) *sparse = 1;
	else if (strcmp(crd, "array") == 0) *sparse = 0;
	else return -1;
	if (strcmp(data, "real") == 0) *storage = MATRIX_STORAGE_REAL;
	else if (strcmp(data, "complex") == 0) *storage = MATRIX_STORAGE_COMPLEX;
	else if (strcmp(data, "pattern") == 0) *storage = MATRIX_STORAGE_PATTERN;
	else if (strcmp(data, "integer") == 0) *storage = MATRIX_STORAGE_INTEGER;
	else return -1;
	if (strcmp(scheme, "general") == 0) *type = MATRIX_TYPE_GENERAL;
	else if (strcmp(scheme, "symmetric") == 0) *type = MATRIX_TYPE_SYMMETRIC;
	else if (strcmp(scheme, "hermitian") == 0) *type = MATRIX_TYPE_HERMITIAN;
	else if (strcmp(scheme, "skew-symmetric") == 0) *type = MATRIX_TYPE_SKEW;
	else return -1;
	/* mm_is_valid (mmio.c:86-94) */
	if (!*sparse && *storage == MATRIX_STORAGE_PATTERN) return -1;
	if (*storage == MATRIX_STORAGE_REAL && *type == MATRIX_TYPE_HERMITIAN) return -1;
	if (*storage == MATRIX_STORAGE_PATTERN && (*type == MATRIX_TYPE_HERMITIAN || *type == MATRIX_TYPE_SKEW)) return -1;
	return 0;
}

/* next line that is neither a comment nor blank; NULL at end of file */
static char* next_data_line(FILE* f, char** buf, size_t* cap)
{
	for (;;) {
		if (getline(buf, cap, f) < 0)
			return NULL;
		char* p = *buf;
		while (*p && isspace((unsigned char)*p)) ++p;
		if (*p == '\0' || *p == '%')
			continue;
		return p;
	}
}

/* size line: "M N nz" (coordinate) or "M N" (array) */
static int read_size(FILE* f, int sparse, int* m, int* n, int* nz)
{
	char* buf = NULL;
	size_t cap = 0;
	char* p = next_data_line(f, &buf, &cap);
	int ok = 0;
	if (p) {
		if (sparse)
			ok = sscanf(p, "%d %d %d", m, n, nz) == 3;
		else {
			ok = sscanf(p, "%d %d", m, n) == 2;
			*nz = ok ? *m * *n : 0;
		}
	}
	free(buf);
	return ok ? 0 : -1;
}

static FILE* open_and_skip_header(const char* path, spgpuMmProperties* props)
{
	FILE* f = fopen(path, "r");
	if (!f)
		return NULL;
	if (read_banner(f, &props->isStoredSparse, &props->matrixStorage, &props->matrixType) != 0 ||
	    read_size(f, props->isStoredSparse, &props->rowsCount, &props->columnsCount, &props->nonZerosCount) != 0 ||
	    props->rowsCount < 0 || props->columnsCount < 0 || props->nonZerosCount < 0) {
		fclose(f);
		return NULL;
	}
	return f;
}

int spgpuMmLoadProperties(const char* path, spgpuMmProperties* props)
{
	memset(props, 0, sizeof *props);
	FILE* f = open_and_skip_header(path, props);
	if (!f)
		return 0;
	fclose(f);
	return 1;
}

int spgpuMmLoadMatrixToCoo(const char* path, void* values, int* rowIndices, int* columnIndices,
	spgpuType_t valuesType)
{
	spgpuMmProperties pr;
	memset(&pr, 0, sizeof pr);
	FILE* f = open_and_skip_header(path, &pr);
	if (!f)
		return MATRIX_READ_INVALID_INPUT;
	int status = MATRIX_READ_SUCCESS;
	if (!pr.isStoredSparse)
		status = MATRIX_READ_INVALID_INPUT;
	else if (values == NULL)
		status = pr.matrixStorage == MATRIX_STORAGE_PATTERN ? MATRIX_READ_SUCCESS : MATRIX_READ_UNSUPPORTED;
	else if (valuesType == SPGPU_TYPE_FLOAT)
		status = (pr.matrixStorage == MATRIX_STORAGE_REAL || pr.matrixStorage == MATRIX_STORAGE_INTEGER)
			? MATRIX_READ_SUCCESS : MATRIX_READ_UNSUPPORTED;
	else if (valuesType == SPGPU_TYPE_DOUBLE)
		status = pr.matrixStorage == MATRIX_STORAGE_REAL ? MATRIX_READ_SUCCESS : MATRIX_READ_UNSUPPORTED;
	else if (valuesType == SPGPU_TYPE_INT)
		status = pr.matrixStorage == MATRIX_STORAGE_INTEGER ? MATRIX_READ_SUCCESS : MATRIX_READ_UNSUPPORTED;
	else
		status = MATRIX_READ_UNSUPPORTED;

	char* buf = NULL;
	size_t cap = 0;
	for (int i = 0; status == MATRIX_READ_SUCCESS && i < pr.nonZerosCount; ++i) {
		char* p = next_data_line(f, &buf, &cap);
		char* end;
		if (!p) { status = MATRIX_READ_INVALID_INPUT; break; }
		const long r = strtol(p, &end, 10);
		if (end == p) { status = MATRIX_READ_INVALID_INPUT; break; }
		p = end;
		const long c = strtol(p, &end, 10);
		if (end == p) { status = MATRIX_READ_INVALID_INPUT; break; }
		p = end;
		rowIndices[i] = (int)r - 1;              /* 1-based -> 0-based (mmread.cpp:88-90) */
		columnIndices[i] = (int)c - 1;
		if (values == NULL)
			continue;
		if (valuesType == SPGPU_TYPE_INT) {
			const long v = strtol(p, &end, 10);
			if (end == p) { status = MATRIX_READ_INVALID_INPUT; break; }
			((int*)values)[i] = (int)v;
		} else {
			const double v = strtod(p, &end);      /* the reference reads %lg and casts (mmread.cpp:79-92) */
			if (end == p) { status = MATRIX_READ_INVALID_INPUT; break; }
			if (valuesType == SPGPU_TYPE_FLOAT) ((float*)values)[i] = (float)v;
			else ((double*)values)[i] = v;
		}
	}
	free(buf);
	fclose(f);
	return status;
}

#define FOR_VALUE_TYPE(valuesType, BODY)                                \
	switch (valuesType) {                                               \
	case SPGPU_TYPE_FLOAT:  { typedef float V;  BODY } break;           \
	case SPGPU_TYPE_DOUBLE: { typedef double V; BODY } break;           \
	case SPGPU_TYPE_INT:    { typedef int V;    BODY } break;           \
	default: break;                                                     \
	}

int spgpuMmUnfoldedSymmetricSize(const void* values, const int* rows, const int* cols,
	int nonZerosCount, spgpuType_t valuesType)
{
	int count = 0;
	FOR_VALUE_TYPE(valuesType,
		const V* v = (const V*)values;
		for (int i = 0; i < nonZerosCount; ++i)
			if (v[i] != 0)
				count += rows[i] == cols[i] ? 1 : 2;
	)
	return count;
}

void spgpuMmUnfoldSymmetric(int* unfoldedRows, int* unfoldedCols, void* unfoldedValues,
	const int* rows, const int* cols, const void* values, int nonZerosCount, spgpuType_t valuesType)
{
	FOR_VALUE_TYPE(valuesType,
		const V* v = (const V*)values;
		V* u = (V*)unfoldedValues;
		int nnz = 0;
		for (int i = 0; i < nonZerosCount; ++i) {
			if (v[i] == 0)
				continue;
			unfoldedRows[nnz] = rows[i]; unfoldedCols[nnz] = cols[i]; u[nnz] = v[i]; ++nnz;
			if (rows[i] != cols[i]) {
				unfoldedRows[nnz] = cols[i]; unfoldedCols[nnz] = rows[i]; u[nnz] = v[i]; ++nnz;
			}
		}
	)
}

int spgpuMmLoadDenseVector(const char* path, void* values, int vectorSize, spgpuType_t valuesType)
{
	spgpuMmProperties pr;
	memset(&pr, 0, sizeof pr);
	FILE* f = open_and_skip_header(path, &pr);
	if (!f)
		return MATRIX_READ_INVALID_INPUT;
	const int wantStorage = valuesType == SPGPU_TYPE_INT ? MATRIX_STORAGE_INTEGER : MATRIX_STORAGE_REAL;
	int status = (pr.isStoredSparse || pr.matrixStorage != wantStorage ||
		(valuesType != SPGPU_TYPE_INT && valuesType != SPGPU_TYPE_FLOAT && valuesType != SPGPU_TYPE_DOUBLE) ||
		(long long)pr.rowsCount * pr.columnsCount < vectorSize) ? MATRIX_READ_INVALID_INPUT : MATRIX_READ_SUCCESS;
	char* buf = NULL;
	size_t cap = 0;
	for (int i = 0; status == MATRIX_READ_SUCCESS && i < vectorSize; ++i) {
		char* p = next_data_line(f, &buf, &cap);
		char* end;
		if (!p) { status = MATRIX_READ_INVALID_INPUT; break; }
		if (valuesType == SPGPU_TYPE_INT) {
			const long v = strtol(p, &end, 10);
			if (end == p) { status = MATRIX_READ_INVALID_INPUT; break; }
			((int*)values)[i] = (int)v;
		} else {
			const double v = strtod(p, &end);
			if (end == p) { status = MATRIX_READ_INVALID_INPUT; break; }
			if (valuesType == SPGPU_TYPE_FLOAT) ((float*)values)[i] = (float)v;
			else ((double*)values)[i] = v;
		}
	}
	free(buf);
	fclose(f);
	return status;
}
