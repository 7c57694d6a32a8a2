/*
 * CPU ORACLE for the spGPU SpMV hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of what the reference's device kernels compute, one
 * function per ABI entry point (same argument lists minus the handle, all
 * pointers are HOST memory).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (spgpu_b200/lib/libspgpu.so) never links or calls it and has no CPU
 * fallback.
 *
 * Pinning: the reference ships no golden vectors for SpMV, so
 * this restatement is pinned against the REFERENCE ITSELF: (1) oracle/Makefile
 * builds the reference's own sources into oracle/_ref/libspgpu_ref.so (host
 * conversions run here on the CPU; its kernels run on the GPU box), and
 * tests/test_parity_reference.py compares oracle == reference kernels == ours
 * on the same inputs; (2) the reference's known-answer tests that do exist
 * (testSparseVector.c:47-125 scatter/gather, ctest.c ELL == HELL) are
 * restated in tests/ against this file.
 *
 * Arithmetic: the reference writes a*b+c (mathbase.cuh / cudalang.h:10-16) and
 * nvcc contracts it to one FMA, so T_FMA is fma()/fmaf() here; complex values
 * follow cuCfma / cuCmul of cuComplex.h term by term.
 *
 * Build: gcc -O3 -fopenmp -fPIC -shared -std=c11 spmv_oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y; } orc_cfloat;
typedef struct { double x, y; } orc_cdouble;

int oracle_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0)
		omp_set_num_threads(n);
#else
	(void)n;
#endif
}

/* ---- complex helpers, term by term as cuComplex.h (cuCfma :334-346, cuCmul) -- */
static inline orc_cfloat cf_fma(orc_cfloat a, orc_cfloat b, orc_cfloat c)
{
	orc_cfloat r;
	r.x = fmaf(a.x, b.x, c.x);
	r.y = fmaf(a.x, b.y, c.y);
	r.x = fmaf(-a.y, b.y, r.x);
	r.y = fmaf(a.y, b.x, r.y);
	return r;
}
static inline orc_cfloat cf_mul(orc_cfloat a, orc_cfloat b)
{
	orc_cfloat r;
	r.x = fmaf(a.x, b.x, -(a.y * b.y));
	r.y = fmaf(a.x, b.y, a.y * b.x);
	return r;
}
static inline orc_cfloat cf_add(orc_cfloat a, orc_cfloat b)
{
	orc_cfloat r = { a.x + b.x, a.y + b.y };
	return r;
}
static inline orc_cdouble cd_fma(orc_cdouble a, orc_cdouble b, orc_cdouble c)
{
	orc_cdouble r;
	r.x = fma(a.x, b.x, c.x);
	r.y = fma(a.x, b.y, c.y);
	r.x = fma(-a.y, b.y, r.x);
	r.y = fma(a.y, b.x, r.y);
	return r;
}
static inline orc_cdouble cd_mul(orc_cdouble a, orc_cdouble b)
{
	orc_cdouble r;
	r.x = fma(a.x, b.x, -(a.y * b.y));
	r.y = fma(a.x, b.y, a.y * b.x);
	return r;
}
static inline orc_cdouble cd_add(orc_cdouble a, orc_cdouble b)
{
	orc_cdouble r = { a.x + b.x, a.y + b.y };
	return r;
}
static const orc_cfloat cf_zero = { 0.0f, 0.0f };
static const orc_cdouble cd_zero = { 0.0, 0.0 };

/* ---- float ---- */
#define SYM S
#define T float
#define R float
#define T_ZERO 0.0f
#define T_FMA(a, b, c) fmaf((a), (b), (c))
#define T_MUL(a, b) ((a) * (b))
#define T_ADD(a, b) ((a) + (b))
#define T_NONZERO(a) ((a) != 0.0f)
#define T_ABS(a) fabsf(a)
#define T_SQABS(a) ((a) * (a))
#include "spmv_oracle_body.inc"
#undef SYM
#undef T
#undef R
#undef T_ZERO
#undef T_FMA
#undef T_MUL
#undef T_ADD
#undef T_NONZERO
#undef T_ABS
#undef T_SQABS

/* ---- double ---- */
#define SYM D
#define T double
#define R double
#define T_ZERO 0.0
#define T_FMA(a, b, c) fma((a), (b), (c))
#define T_MUL(a, b) ((a) * (b))
#define T_ADD(a, b) ((a) + (b))
#define T_NONZERO(a) ((a) != 0.0)
#define T_ABS(a) fabs(a)
#define T_SQABS(a) ((a) * (a))
#include "spmv_oracle_body.inc"
#undef SYM
#undef T
#undef R
#undef T_ZERO
#undef T_FMA
#undef T_MUL
#undef T_ADD
#undef T_NONZERO
#undef T_ABS
#undef T_SQABS

/* ---- complex float ---- */
#define SYM C
#define T orc_cfloat
#define R float
#define T_ZERO cf_zero
#define T_FMA(a, b, c) cf_fma((a), (b), (c))
#define T_MUL(a, b) cf_mul((a), (b))
#define T_ADD(a, b) cf_add((a), (b))
#define T_NONZERO(a) ((a).x != 0.0f || (a).y != 0.0f)
#define T_ABS(a) hypotf((a).x, (a).y)
#define T_SQABS(a) ((a).x * (a).x + (a).y * (a).y)
#include "spmv_oracle_body.inc"
#undef SYM
#undef T
#undef R
#undef T_ZERO
#undef T_FMA
#undef T_MUL
#undef T_ADD
#undef T_NONZERO
#undef T_ABS
#undef T_SQABS

/* ---- complex double ---- */
#define SYM Z
#define T orc_cdouble
#define R double
#define T_ZERO cd_zero
#define T_FMA(a, b, c) cd_fma((a), (b), (c))
#define T_MUL(a, b) cd_mul((a), (b))
#define T_ADD(a, b) cd_add((a), (b))
#define T_NONZERO(a) ((a).x != 0.0 || (a).y != 0.0)
#define T_ABS(a) hypot((a).x, (a).y)
#define T_SQABS(a) ((a).x * (a).x + (a).y * (a).y)
#include "spmv_oracle_body.inc"
#undef SYM
#undef T
#undef R
#undef T_ZERO
#undef T_FMA
#undef T_MUL
#undef T_ADD
#undef T_NONZERO
#undef T_ABS
#undef T_SQABS

/* ---- dot: sum a_i*b_i, complex UNCONJUGATED (reference zdot.cu:54, cdot.cu),
 * accumulated in long double; results returned through pointers so the complex
 * ones need no struct-return ABI. ---- */
void oracle_Sdot(int n, const float* a, const float* b, double* out)
{
	long double s = 0.0L;
	for (long i = 0; i < n; ++i)
		s += (long double)a[i] * (long double)b[i];
	out[0] = (double)s;
}
void oracle_Ddot(int n, const double* a, const double* b, double* out)
{
	long double s = 0.0L;
	for (long i = 0; i < n; ++i)
		s += (long double)a[i] * (long double)b[i];
	out[0] = (double)s;
}
void oracle_Cdot(int n, const orc_cfloat* a, const orc_cfloat* b, double* out)
{
	long double re = 0.0L, im = 0.0L;
	for (long i = 0; i < n; ++i) {
		re += (long double)a[i].x * b[i].x - (long double)a[i].y * b[i].y;
		im += (long double)a[i].x * b[i].y + (long double)a[i].y * b[i].x;
	}
	out[0] = (double)re;
	out[1] = (double)im;
}
void oracle_Zdot(int n, const orc_cdouble* a, const orc_cdouble* b, double* out)
{
	long double re = 0.0L, im = 0.0L;
	for (long i = 0; i < n; ++i) {
		re += (long double)a[i].x * b[i].x - (long double)a[i].y * b[i].y;
		im += (long double)a[i].x * b[i].y + (long double)a[i].y * b[i].x;
	}
	out[0] = (double)re;
	out[1] = (double)im;
}

/* int gather / scatter (reference igath.cu / iscat.cu through the same templates) */
void oracle_Igath(int* xValues, int xNnz, const int* xIndices, int xBaseIndex, const int* y)
{
	for (long i = 0; i < xNnz; ++i) {
		long p = (long)xIndices[i] - xBaseIndex;
		if (p >= 0)
			xValues[i] = y[p];
	}
}
void oracle_Iscat(int* y, int xNnz, const int* xValues, const int* xIndices, int xBaseIndex, int beta)
{
	for (long i = 0; i < xNnz; ++i) {
		long p = (long)xIndices[i] - xBaseIndex;
		if (p < 0)
			continue;
		y[p] = beta != 0 ? beta * y[p] + xValues[i] : xValues[i];
	}
}
