/*
 * TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the reference's own MatrixMarket reader
 * (reference src/utils/mmread.cpp, mmutils.hpp, src/external/mmio.c -- compiled from where they
 * lie by oracle/Makefile into oracle/_ref/libmm_ref.so) so tests/test_mmread.py can compare
 * spgpu_b200's reader (csrc/mmread.c) with it on the same files.  The reference functions
 * are overloaded C++ taking FILE*; this file only opens the path and forwards.
 */
#include <cstdio>
#include "utils/mmread.hpp"
#include "utils/mmutils.hpp"

extern "C" {

int ref_mm_properties(const char* path, int* out6)
{
	FILE* f = fopen(path, "r");
	if (!f) return 0;
	bool sparse = false;
	int storage = -1, type = -1;
	const bool ok = loadMmProperties(&out6[0], &out6[1], &out6[2], &sparse, &storage, &type, f);
	out6[3] = sparse; out6[4] = storage; out6[5] = type;
	fclose(f);
	return ok ? 1 : 0;
}

#define REF_MM_LOAD(NAME, T)                                                               \
	int NAME(const char* path, T* values, int* rows, int* cols)                            \
	{                                                                                      \
		FILE* f = fopen(path, "r");                                                        \
		if (!f) return -1;                                                                 \
		int m, n, nz, storage, type; bool sparse;                                          \
		if (!loadMmProperties(&m, &n, &nz, &sparse, &storage, &type, f)) { fclose(f); return -1; } \
		const int r = loadMmMatrixToCoo(values, rows, cols, m, n, nz, sparse, storage, f); \
		fclose(f);                                                                         \
		return r;                                                                          \
	}
REF_MM_LOAD(ref_mm_load_float, float)
REF_MM_LOAD(ref_mm_load_double, double)
REF_MM_LOAD(ref_mm_load_int, int)

int ref_mm_load_pattern(const char* path, int* rows, int* cols)
{
	FILE* f = fopen(path, "r");
	if (!f) return -1;
	int m, n, nz, storage, type; bool sparse;
	if (!loadMmProperties(&m, &n, &nz, &sparse, &storage, &type, f)) { fclose(f); return -1; }
	const int r = loadMmMatrixToCoo(rows, cols, m, n, nz, sparse, storage, f);
	fclose(f);
	return r;
}

#define REF_MM_UNFOLD(SUFFIX, T)                                                           \
	int ref_mm_unfolded_size_##SUFFIX(T* values, int* rows, int* cols, int nnz)            \
	{                                                                                      \
		int count = 0;                                                                     \
		getUnfoldedMmSymmetricSize(&count, values, rows, cols, nnz);                       \
		return count;                                                                      \
	}                                                                                      \
	void ref_mm_unfold_##SUFFIX(int* urows, int* ucols, T* uvals, int* rows, int* cols, T* values, int nnz) \
	{                                                                                      \
		unfoldMmSymmetricReal(urows, ucols, uvals, rows, cols, values, nnz);               \
	}
REF_MM_UNFOLD(float, float)
REF_MM_UNFOLD(double, double)

int ref_mm_load_vector_double(const char* path, double* values, int n)
{
	FILE* f = fopen(path, "r");
	if (!f) return -1;
	/* banner, comments and the two-number size line of an array file are skipped by hand: the
	 * reference has no header reader for vectors (loadMmProperties wants three numbers) */
	char line[1100];
	if (!fgets(line, sizeof line, f)) { fclose(f); return -1; }
	do { if (!fgets(line, sizeof line, f)) { fclose(f); return -1; } } while (line[0] == '%');
	const int r = loadMmVectorToDenseVector(values, n, MATRIX_STORAGE_REAL, f);
	fclose(f);
	return r;
}

}
