/*
 * Element-wise BLAS-1 companions for sm_100a: axpby, scal, abs, axy, axypbz,
 * setscal and their multi-vector forms.
 *
 * Replaces reference kernels/{s,d,c,z}axpby.cu, scal_base.cuh, abs_base.cuh,
 * axy_base.cuh, setscal_base.cuh.  Same results and the same aliasing rule:
 * an output may be exactly one of the inputs (every thread reads the elements
 * it is about to overwrite before it writes them).
 *
 * These are pure HBM streams, so the kernels move 128 bits per thread per
 * access (2 packs in flight per input per thread), run as a grid-stride loop
 * on a grid sized from the SM count, and fall back to scalar accesses only
 * when a pointer is not 16-byte aligned.  The reference launches one thread
 * per element with 4/8-byte accesses.
 */
#include "launch.cuh"
#include "numeric.cuh"

template <typename T> struct alignas(16) Pack {
	static constexpr int N = 16 / (int)sizeof(T);
	T v[N];
};

/* Op::apply(a, b, c) combines up to three input elements into the output. */
template <typename T, int NIN, typename Op>
__global__ void __launch_bounds__(256)
ew_kernel(T* out, const T* in0, const T* in1, const T* in2, long long n, Op op, int vec)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	constexpr int N = Pack<T>::N;

	if (vec) {
		const long long npacks = n / N;
		Pack<T>* o = reinterpret_cast<Pack<T>*>(out);
		const Pack<T>* p0 = reinterpret_cast<const Pack<T>*>(in0);
		const Pack<T>* p1 = reinterpret_cast<const Pack<T>*>(in1);
		const Pack<T>* p2 = reinterpret_cast<const Pack<T>*>(in2);
		for (long long p = tid; p < npacks; p += 2 * nthreads) {
			const long long q = p + nthreads;
			const bool two = q < npacks;
			Pack<T> a0, b0, c0, a1, b1, c1, r0, r1;
			if (NIN > 0) { a0 = p0[p]; if (two) a1 = p0[q]; }
			if (NIN > 1) { b0 = p1[p]; if (two) b1 = p1[q]; }
			if (NIN > 2) { c0 = p2[p]; if (two) c1 = p2[q]; }
#pragma unroll
			for (int e = 0; e < N; ++e) {
				r0.v[e] = op.apply(a0.v[e], b0.v[e], c0.v[e]);
				r1.v[e] = op.apply(a1.v[e], b1.v[e], c1.v[e]);
			}
			o[p] = r0;
			if (two) o[q] = r1;
		}
		/* tail elements that do not fill a pack */
		const long long done = npacks * N;
		if (tid < n - done) {
			const long long e = done + tid;
			T a = NIN > 0 ? in0[e] : Num<T>::zero();
			T b = NIN > 1 ? in1[e] : Num<T>::zero();
			T c = NIN > 2 ? in2[e] : Num<T>::zero();
			out[e] = op.apply(a, b, c);
		}
	} else {
		for (long long e = tid; e < n; e += nthreads) {
			T a = NIN > 0 ? in0[e] : Num<T>::zero();
			T b = NIN > 1 ? in1[e] : Num<T>::zero();
			T c = NIN > 2 ? in2[e] : Num<T>::zero();
			out[e] = op.apply(a, b, c);
		}
	}
}

static inline int aligned16(const void* p) { return ((size_t)p & 15) == 0; }

template <typename T, int NIN, typename Op>
static void ew_launch(spgpuHandle_t handle, T* out, const T* in0, const T* in1,
	const T* in2, long long n, Op op)
{
	if (n <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int block = 256;
	const int vec = aligned16(out) && (NIN < 1 || aligned16(in0)) &&
		(NIN < 2 || aligned16(in1)) && (NIN < 3 || aligned16(in2));
	const long long items = vec ? (n / Pack<T>::N + 1) / 2 + 1 : n;
	long long want = (items + block - 1) / block;
	const long long cap = (long long)handle->multiProcessorCount * (t->vecBlocksPerSm > 0 ? t->vecBlocksPerSm : 8);
	if (cap > 0 && want > cap) want = cap;
	if (want < 1) want = 1;
	spgpu_launch_dep(handle, ew_kernel<T, NIN, Op>, (unsigned)want, block, out, in0, in1, in2, n, op, vec);
	spgpu_count_launch(handle);
}

/* ---- operations ----------------------------------------------------------- */

/* alpha*x            (reference scal_base.cuh:44; axpby with beta == 0) */
template <typename T> struct OpScale {
	T alpha;
	__device__ __forceinline__ T apply(T x, T, T) const { return Num<T>::mul(alpha, x); }
};
/* beta*y + alpha*x   (reference daxpby.cu:40-43, caxpby.cu:41-44); inputs (y, x) */
template <typename T> struct OpAxpby {
	T alpha, beta;
	__device__ __forceinline__ T apply(T y, T x, T) const { return Num<T>::fma(beta, y, Num<T>::mul(alpha, x)); }
};
/* alpha*(x*y)        (reference axy_base.cuh:49) */
template <typename T> struct OpAxy {
	T alpha;
	__device__ __forceinline__ T apply(T x, T y, T) const { return Num<T>::mul(alpha, Num<T>::mul(x, y)); }
};
/* alpha*(x*y) + beta*z   (reference axy_base.cuh:119-122); inputs (z, x, y) */
template <typename T> struct OpAxypbz {
	T alpha, beta;
	__device__ __forceinline__ T apply(T z, T x, T y) const
	{
		return Num<T>::fma(alpha, Num<T>::mul(x, y), Num<T>::mul(beta, z));
	}
};
/* alpha*|x| (result stored in T; complex gets (|x|, 0) like reference mathbase.cuh) */
template <typename T> struct OpAbs {
	T alpha;
	bool scale;
	__device__ __forceinline__ T apply(T x, T, T) const
	{
		T a = Num<T>::from_real(Num<T>::abs(x));
		return scale ? Num<T>::mul(alpha, a) : a;
	}
};
template <typename T> struct OpFill {
	T val;
	__device__ __forceinline__ T apply(T, T, T) const { return val; }
};

template <typename T> static bool is_one(T a);
template <> bool is_one<float>(float a) { return a == 1.0f; }
template <> bool is_one<double>(double a) { return a == 1.0; }
template <> bool is_one<cuFloatComplex>(cuFloatComplex a) { return a.x == 1.0f && a.y == 0.0f; }
template <> bool is_one<cuDoubleComplex>(cuDoubleComplex a) { return a.x == 1.0 && a.y == 0.0; }

/* ---- typed drivers ---------------------------------------------------------- */

template <typename T>
static void scal_impl(spgpuHandle_t h, T* y, int n, T alpha, const T* x)
{
	OpScale<T> op = { alpha };
	ew_launch<T, 1>(h, y, x, (const T*)0, (const T*)0, n, op);
}

template <typename T>
static void axpby_impl(spgpuHandle_t h, T* z, int n, T beta, const T* y, T alpha, const T* x)
{
	if (!Num<T>::nonzero(beta)) {          /* y is not read */
		scal_impl<T>(h, z, n, alpha, x);
		return;
	}
	OpAxpby<T> op = { alpha, beta };
	ew_launch<T, 2>(h, z, y, x, (const T*)0, n, op);
}

template <typename T>
static void axy_impl(spgpuHandle_t h, T* z, int n, T alpha, const T* x, const T* y)
{
	OpAxy<T> op = { alpha };
	ew_launch<T, 2>(h, z, x, y, (const T*)0, n, op);
}

template <typename T>
static void axypbz_impl(spgpuHandle_t h, T* w, int n, T beta, const T* z, T alpha, const T* x, const T* y)
{
	/* degenerate cases exactly as reference axy_base.cuh:156-164 */
	if (!Num<T>::nonzero(alpha)) {
		scal_impl<T>(h, w, n, beta, z);
	} else if (!Num<T>::nonzero(beta)) {
		axy_impl<T>(h, w, n, alpha, x, y);
	} else {
		OpAxypbz<T> op = { alpha, beta };
		ew_launch<T, 3>(h, w, z, x, y, n, op);
	}
}

template <typename T>
static void abs_impl(spgpuHandle_t h, T* y, int n, T alpha, const T* x)
{
	OpAbs<T> op = { alpha, !is_one<T>(alpha) };
	ew_launch<T, 1>(h, y, x, (const T*)0, (const T*)0, n, op);
}

template <typename T>
static void setscal_impl(spgpuHandle_t h, int first, int last, int baseIndex, T val, T* y)
{
	OpFill<T> op = { val };
	const long long n = (long long)last - first + 1;
	ew_launch<T, 0>(h, y + (first - baseIndex), (const T*)0, (const T*)0, (const T*)0, n, op);
}

/* ---- C entry points -------------------------------------------------------- */

#define SPGPU_DEFINE_STREAM(S, T)                                              \
	extern "C" void spgpu##S##scal(spgpuHandle_t h, T* y, int n, T alpha, T* x) \
	{ scal_impl<T>(h, y, n, alpha, x); }                                        \
	extern "C" void spgpu##S##axpby(spgpuHandle_t h, T* z, int n, T beta, T* y, \
		T alpha, T* x)                                                          \
	{ axpby_impl<T>(h, z, n, beta, y, alpha, x); }                              \
	extern "C" void spgpu##S##maxpby(spgpuHandle_t h, T* z, int n, T beta,      \
		T* y, T alpha, T* x, int count, int pitch)                              \
	{                                                                           \
		for (int v = 0; v < count; ++v) {                                       \
			const long long o = (long long)v * pitch;                           \
			axpby_impl<T>(h, z + o, n, beta, y + o, alpha, x + o);              \
		}                                                                       \
	}                                                                           \
	extern "C" void spgpu##S##abs(spgpuHandle_t h, T* y, int n, T alpha, T* x)  \
	{ abs_impl<T>(h, y, n, alpha, x); }                                         \
	extern "C" void spgpu##S##axy(spgpuHandle_t h, T* z, int n, T alpha, T* x,  \
		T* y)                                                                   \
	{ axy_impl<T>(h, z, n, alpha, x, y); }                                      \
	extern "C" void spgpu##S##axypbz(spgpuHandle_t h, T* w, int n, T beta,      \
		T* z, T alpha, T* x, T* y)                                              \
	{ axypbz_impl<T>(h, w, n, beta, z, alpha, x, y); }                          \
	extern "C" void spgpu##S##maxy(spgpuHandle_t h, T* z, int n, T alpha, T* x, \
		T* y, int count, int pitch)                                             \
	{                                                                           \
		for (int v = 0; v < count; ++v) {                                       \
			const long long o = (long long)v * pitch;                           \
			axy_impl<T>(h, z + o, n, alpha, x + o, y + o);                      \
		}                                                                       \
	}                                                                           \
	extern "C" void spgpu##S##maxypbz(spgpuHandle_t h, T* w, int n, T beta,     \
		T* z, T alpha, T* x, T* y, int count, int pitch)                        \
	{                                                                           \
		for (int v = 0; v < count; ++v) {                                       \
			const long long o = (long long)v * pitch;                           \
			axypbz_impl<T>(h, w + o, n, beta, z + o, alpha, x + o, y + o);      \
		}                                                                       \
	}                                                                           \
	extern "C" void spgpu##S##setscal(spgpuHandle_t h, int first, int last,     \
		int baseIndex, T val, T* y)                                             \
	{ setscal_impl<T>(h, first, last, baseIndex, val, y); }

SPGPU_DEFINE_STREAM(S, float)
SPGPU_DEFINE_STREAM(D, double)
SPGPU_DEFINE_STREAM(C, cuFloatComplex)
SPGPU_DEFINE_STREAM(Z, cuDoubleComplex)

extern "C" void spgpuIsetscal(spgpuHandle_t h, int first, int last, int baseIndex, int val, int* y)
{
	setscal_impl<int>(h, first, last, baseIndex, val, y);
}
