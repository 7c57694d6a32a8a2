/*
 * Value-type traits and memory-access helpers shared by every kernel.
 *
 * Arithmetic follows what the reference's device code computes after nvcc's
 * FMA contraction (reference kernels/mathbase.cuh: a*b+c; complex through
 * cuCfma / cuCmul of cuComplex.h), so results agree with it to rounding.
 *
 * Memory policy on B200: the matrix arrays (values, column indices, row sizes)
 * are read exactly once per SpMV -> `ld_stream` (ld.global.cs: evict-first,
 * keeps them from pushing x out of the 126 MB L2); x is re-read by
 * neighbouring rows -> `ld_keep` (ld.global.nc, read-only path, normal L1/L2
 * allocation).
 */
#ifndef SPGPU_NUMERIC_CUH_
#define SPGPU_NUMERIC_CUH_

#include <cuda_runtime.h>
#include <cuComplex.h>

#define SPGPU_FULL_MASK 0xffffffffu

template <typename T> struct Num;

template <> struct Num<float> {
	typedef float real;
	static constexpr bool is_complex = false;
	static __host__ __device__ __forceinline__ float zero() { return 0.0f; }
	static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
	static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
	static __device__ __forceinline__ float div(float a, float b) { return a / b; }
	static __device__ __forceinline__ float add(float a, float b) { return a + b; }
	static __host__ __device__ __forceinline__ bool nonzero(float a) { return a != 0.0f; }
	static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
	static __device__ __forceinline__ float sqabs(float a) { return a * a; }
	static __host__ __device__ __forceinline__ float from_real(float r) { return r; }
};

template <> struct Num<double> {
	typedef double real;
	static constexpr bool is_complex = false;
	static __host__ __device__ __forceinline__ double zero() { return 0.0; }
	static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
	static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
	static __device__ __forceinline__ double div(double a, double b) { return a / b; }
	static __device__ __forceinline__ double add(double a, double b) { return a + b; }
	static __host__ __device__ __forceinline__ bool nonzero(double a) { return a != 0.0; }
	static __device__ __forceinline__ double abs(double a) { return fabs(a); }
	static __device__ __forceinline__ double sqabs(double a) { return a * a; }
	static __host__ __device__ __forceinline__ double from_real(double r) { return r; }
};

template <> struct Num<cuFloatComplex> {
	typedef float real;
	static constexpr bool is_complex = true;
	static __host__ __device__ __forceinline__ cuFloatComplex zero() { return make_cuFloatComplex(0.0f, 0.0f); }
	static __device__ __forceinline__ cuFloatComplex fma(cuFloatComplex a, cuFloatComplex b, cuFloatComplex c) { return cuCfmaf(a, b, c); }
	static __device__ __forceinline__ cuFloatComplex mul(cuFloatComplex a, cuFloatComplex b) { return cuCmulf(a, b); }
	static __device__ __forceinline__ cuFloatComplex div(cuFloatComplex a, cuFloatComplex b) { return cuCdivf(a, b); }
	static __device__ __forceinline__ cuFloatComplex add(cuFloatComplex a, cuFloatComplex b) { return cuCaddf(a, b); }
	static __host__ __device__ __forceinline__ bool nonzero(cuFloatComplex a) { return a.x != 0.0f || a.y != 0.0f; }
	static __device__ __forceinline__ float abs(cuFloatComplex a) { return cuCabsf(a); }
	static __device__ __forceinline__ float sqabs(cuFloatComplex a) { return fmaf(a.x, a.x, a.y * a.y); }
	static __host__ __device__ __forceinline__ cuFloatComplex from_real(float r) { return make_cuFloatComplex(r, 0.0f); }
};

template <> struct Num<cuDoubleComplex> {
	typedef double real;
	static constexpr bool is_complex = true;
	static __host__ __device__ __forceinline__ cuDoubleComplex zero() { return make_cuDoubleComplex(0.0, 0.0); }
	static __device__ __forceinline__ cuDoubleComplex fma(cuDoubleComplex a, cuDoubleComplex b, cuDoubleComplex c) { return cuCfma(a, b, c); }
	static __device__ __forceinline__ cuDoubleComplex mul(cuDoubleComplex a, cuDoubleComplex b) { return cuCmul(a, b); }
	static __device__ __forceinline__ cuDoubleComplex div(cuDoubleComplex a, cuDoubleComplex b) { return cuCdiv(a, b); }
	static __device__ __forceinline__ cuDoubleComplex add(cuDoubleComplex a, cuDoubleComplex b) { return cuCadd(a, b); }
	static __host__ __device__ __forceinline__ bool nonzero(cuDoubleComplex a) { return a.x != 0.0 || a.y != 0.0; }
	static __device__ __forceinline__ double abs(cuDoubleComplex a) { return cuCabs(a); }
	static __device__ __forceinline__ double sqabs(cuDoubleComplex a) { return ::fma(a.x, a.x, a.y * a.y); }
	static __host__ __device__ __forceinline__ cuDoubleComplex from_real(double r) { return make_cuDoubleComplex(r, 0.0); }
};

/* int "arithmetic" for the I gather/scatter (reference mathbase.cuh int_fma) */
template <> struct Num<int> {
	typedef int real;
	static constexpr bool is_complex = false;
	static __host__ __device__ __forceinline__ int zero() { return 0; }
	static __device__ __forceinline__ int fma(int a, int b, int c) { return a * b + c; }
	static __device__ __forceinline__ int mul(int a, int b) { return a * b; }
	static __device__ __forceinline__ int add(int a, int b) { return a + b; }
	static __host__ __device__ __forceinline__ bool nonzero(int a) { return a != 0; }
};

/* ---- loads ---------------------------------------------------------------- */

/* read-once stream: evict-first */
template <typename T> __device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
/*
 * Read-once stream whose 128-byte lines may be only PARTLY used (rows of unequal length: lanes past their row's
 * end are switched off): evict-first AND an L2 fetch granularity of 64 bytes (SASS LDG.E.EF.LTC64B).  Measured on
 * B200 (bench/probes/sector_probe.cu, profiles/r2_sector_probe.md): whatever the load flavour, an L2 miss on one
 * 32-byte sector fetches the WHOLE 128-byte line from DRAM -- cudaLimitMaxL2FetchGranularity is ignored -- except
 * with the .L2::64B qualifier, which halves it; there is no 32-byte option.  Full lines cost the same either way.
 */
template <typename T> __device__ __forceinline__ T ld_stream64(const T* p);
template <> __device__ __forceinline__ int ld_stream64<int>(const int* p)
{
	int v;
	asm("ld.global.cs.L2::64B.s32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
template <> __device__ __forceinline__ float ld_stream64<float>(const float* p)
{
	float v;
	asm("ld.global.cs.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
	return v;
}
template <> __device__ __forceinline__ double ld_stream64<double>(const double* p)
{
	double v;
	asm("ld.global.cs.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
	return v;
}
template <> __device__ __forceinline__ cuFloatComplex ld_stream64<cuFloatComplex>(const cuFloatComplex* p)
{
	cuFloatComplex v;
	asm("ld.global.cs.L2::64B.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
	return v;
}
template <> __device__ __forceinline__ cuDoubleComplex ld_stream64<cuDoubleComplex>(const cuDoubleComplex* p)
{
	cuDoubleComplex v;
	asm("ld.global.cs.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
	return v;
}

/* bring the line holding *p into L2 (no register, no wait) */
__device__ __forceinline__ void prefetch_l2(const void* p)
{
	asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

/* read-only, cache normally (x gathers) */
template <typename T> __device__ __forceinline__ T ld_keep(const T* p) { return __ldg(p); }

/* ---- how a row walk fetches x[c] ------------------------------------------- */

/* the single-GPU kernels: x is read-only for the whole launch -> read-only path (ld.global.nc) */
template <typename T> struct XPlain {
	const T* x;
	__device__ __forceinline__ T ld(int c) const { return __ldg(x + c); }
};

/* x entries that a PEER GPU may write during the launch (halo zones of the fused kernels, ext_halo.cu): plain weak
 * loads (ld.global, SASS LDG.E) -- inside the memory model, ordered after the CTA's acquire on the ready flag; .nc
 * loads are not.  (L1::evict_last on these loads measures the same as the default policy.) */
template <typename T> struct XWeak {
	const T* x;
	__device__ __forceinline__ T ld(int c) const { return x[c]; }
};


/* ---- warp helpers --------------------------------------------------------- */

template <typename T> __device__ __forceinline__ T shfl_xor(T v, int m);
template <> __device__ __forceinline__ float shfl_xor<float>(float v, int m) { return __shfl_xor_sync(SPGPU_FULL_MASK, v, m); }
template <> __device__ __forceinline__ double shfl_xor<double>(double v, int m) { return __shfl_xor_sync(SPGPU_FULL_MASK, v, m); }
template <> __device__ __forceinline__ cuFloatComplex shfl_xor<cuFloatComplex>(cuFloatComplex v, int m)
{
	return make_cuFloatComplex(__shfl_xor_sync(SPGPU_FULL_MASK, v.x, m), __shfl_xor_sync(SPGPU_FULL_MASK, v.y, m));
}
template <> __device__ __forceinline__ cuDoubleComplex shfl_xor<cuDoubleComplex>(cuDoubleComplex v, int m)
{
	return make_cuDoubleComplex(__shfl_xor_sync(SPGPU_FULL_MASK, v.x, m), __shfl_xor_sync(SPGPU_FULL_MASK, v.y, m));
}

/* butterfly sum over the 32 lanes; every lane ends with the total */
template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1)
		v = Num<T>::add(v, shfl_xor<T>(v, m));
	return v;
}

/* z = beta*y + alpha*acc, the epilogue every reference SpMV kernel ends with
 * (e.g. reference hell_spmv_base_template.cuh:219-222). */
template <typename T>
__device__ __forceinline__ T spmv_epilogue(T acc, T alpha, T beta, bool useBeta, T yv)
{
	T scaled = Num<T>::mul(alpha, acc);
	return useBeta ? Num<T>::fma(beta, yv, scaled) : scaled;
}

#endif
