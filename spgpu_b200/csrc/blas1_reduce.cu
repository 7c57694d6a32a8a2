/*
 * Reductions for sm_100a: dot, nrm2, amax, asum (+ multi-vector forms and the
 * additive device-result variants of include/spgpu_ext.h).
 *
 * Replaces reference kernels/{s,d,c,z}dot.cu, {s,d,c,z}nrm2.cu, amax_base.cuh,
 * asum_base.cuh.  Results: dot is NOT conjugated for C/Z (reference zdot.cu:54);
 * nrm2 = sqrt(sum |x|^2) unscaled (dnrm2.cu:52-53,146); amax / asum are the
 * documented full reductions (the reference's last warp steps discard lanes
 * 2..31 and truncate doubles to float, SURVEY 2.3 -- not reproduced).
 *
 * One pass, one kernel: 128-bit loads in a grid-stride loop over
 * (#SM x redBlocksPerSm) CTAs, warp-shuffle + shared-memory block reduction,
 * one partial per CTA into the HANDLE's scratch (the reference uses a
 * file-static __device__ array shared by every handle), and the last CTA to
 * finish (ticket counter) folds the partials in a fixed order -- so the result
 * is deterministic for a given n -- and stores the final value straight into
 * mapped pinned host memory.  The blocking entry points then only need a
 * stream synchronise; no cudaMemcpyFromSymbol on the legacy stream.
 * Partials are combined in double for every type.
 */
#include <cstring>
#include "launch.cuh"
#include "reduce_common.cuh"

#define RED_BLOCK 256

template <typename T> struct alignas(16) RPack {
	static constexpr int N = 16 / (int)sizeof(T);
	T v[N];
};

/* ---- per-operation policy -------------------------------------------------- */
/* Each op accumulates in the element's real precision per thread (short chains:
 * n / (grid*block) terms) and hands over doubles for the cross-thread part.   */

template <typename T> struct OpDot {
	static constexpr int NIN = 2;
	static constexpr bool IS_MAX = false;
	T s;
	__device__ __forceinline__ void init() { s = Num<T>::zero(); }
	__device__ __forceinline__ void take(T x, T y) { s = Num<T>::fma(x, y, s); }
	__device__ __forceinline__ Acc2 result() const;
};
template <> __device__ __forceinline__ Acc2 OpDot<float>::result() const { return { (double)s, 0.0 }; }
template <> __device__ __forceinline__ Acc2 OpDot<double>::result() const { return { s, 0.0 }; }
template <> __device__ __forceinline__ Acc2 OpDot<cuFloatComplex>::result() const { return { (double)s.x, (double)s.y }; }
template <> __device__ __forceinline__ Acc2 OpDot<cuDoubleComplex>::result() const { return { s.x, s.y }; }

template <typename T> struct OpSqSum {
	static constexpr int NIN = 1;
	static constexpr bool IS_MAX = false;
	typename Num<T>::real s;
	__device__ __forceinline__ void init() { s = 0; }
	__device__ __forceinline__ void take(T x, T) { s += Num<T>::sqabs(x); }
	__device__ __forceinline__ Acc2 result() const { return { (double)s, 0.0 }; }
};

template <typename T> struct OpSum {
	static constexpr int NIN = 1;
	static constexpr bool IS_MAX = false;
	T s;
	__device__ __forceinline__ void init() { s = Num<T>::zero(); }
	__device__ __forceinline__ void take(T x, T) { s = Num<T>::add(s, x); }
	__device__ __forceinline__ Acc2 result() const;
};
template <> __device__ __forceinline__ Acc2 OpSum<float>::result() const { return { (double)s, 0.0 }; }
template <> __device__ __forceinline__ Acc2 OpSum<double>::result() const { return { s, 0.0 }; }
template <> __device__ __forceinline__ Acc2 OpSum<cuFloatComplex>::result() const { return { (double)s.x, (double)s.y }; }
template <> __device__ __forceinline__ Acc2 OpSum<cuDoubleComplex>::result() const { return { s.x, s.y }; }

template <typename T> struct OpAbsSum {
	static constexpr int NIN = 1;
	static constexpr bool IS_MAX = false;
	typename Num<T>::real s;
	__device__ __forceinline__ void init() { s = 0; }
	__device__ __forceinline__ void take(T x, T) { s += Num<T>::abs(x); }
	__device__ __forceinline__ Acc2 result() const { return { (double)s, 0.0 }; }
};

template <typename T> struct OpAbsMax {
	static constexpr int NIN = 1;
	static constexpr bool IS_MAX = true;
	typename Num<T>::real s;
	__device__ __forceinline__ void init() { s = 0; }
	__device__ __forceinline__ void take(T x, T) { s = fmax(s, Num<T>::abs(x)); }
	__device__ __forceinline__ Acc2 result() const { return { (double)s, 0.0 }; }
};

/*
 * finish: 0 sum as is, 1 sqrt of the sum (nrm2).  outKind: how the final value
 * is stored: 0 = as T (dot), 1 = as real of T.
 *
 * grid.y = vectors of a multi-vector call (spgpu?m{dot,nrm2,asum,amax}): vector v starts `pitch`
 * elements after vector v-1, has its own run of gridDim.x partial slots, its own ticket word and
 * its own result slot (outBytes apart) -- `count` reductions in ONE launch.
 *
 * INFL = 16-byte packs a thread keeps in flight per input (reduce_inflight below).
 */
template <typename T, typename Op, int INFL>
__global__ void __launch_bounds__(RED_BLOCK)
reduce_kernel(const T* x, const T* y, long long n, long long pitch, Acc2* partials,
	unsigned* tickets, void* out, int outBytes, int finish, int outKind)
{
	__shared__ Acc2 smem[32];
	__shared__ bool amLast;
	grid_dependency_wait();
	grid_launch_dependents();
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	constexpr int N = RPack<T>::N;

	x += (long long)blockIdx.y * pitch;
	if (Op::NIN > 1)
		y += (long long)blockIdx.y * pitch;
	partials += (size_t)blockIdx.y * gridDim.x;
	unsigned* ticket = tickets + blockIdx.y;
	out = reinterpret_cast<char*>(out) + (size_t)blockIdx.y * outBytes;
	const bool vec = (((size_t)x | (Op::NIN > 1 ? (size_t)y : (size_t)0)) & 15) == 0;

	Op acc0, acc1;
	acc0.init();
	acc1.init();

	if (vec) {
		const long long npacks = n / N;
		const RPack<T>* px = reinterpret_cast<const RPack<T>*>(x);
		const RPack<T>* py = reinterpret_cast<const RPack<T>*>(y);
		for (long long p = tid; p < npacks; p += INFL * nthreads) {
			RPack<T> a[INFL], b[INFL];
#pragma unroll
			for (int k = 0; k < INFL; ++k) {
				const long long q = p + k * nthreads;
				if (q < npacks) {
					a[k] = px[q];
					if (Op::NIN > 1) b[k] = py[q];
				}
			}
#pragma unroll
			for (int k = 0; k < INFL; ++k) {
				if (p + k * nthreads < npacks) {
#pragma unroll
					for (int e = 0; e < N; ++e) {
						if (k & 1) acc1.take(a[k].v[e], b[k].v[e]);
						else       acc0.take(a[k].v[e], b[k].v[e]);
					}
				}
			}
		}
		const long long done = npacks * N;
		if (tid < n - done)
			acc0.take(x[done + tid], Op::NIN > 1 ? y[done + tid] : Num<T>::zero());
	} else {
		for (long long e = tid; e < n; e += nthreads)
			acc0.take(x[e], Op::NIN > 1 ? y[e] : Num<T>::zero());
	}

	Acc2 v = combine<Op::IS_MAX>(acc0.result(), acc1.result());
	v = block_reduce<Op::IS_MAX>(v, smem);

	Acc2 total;
	if (!reduce_finish<Op::IS_MAX>(v, partials, ticket, smem, &amLast, total))
		return;
	if (threadIdx.x == 0) {
		double a = total.a, b = total.b;
		if (finish == 1)
			a = sqrt(a);
		typedef typename Num<T>::real R;
		if (outKind == 1 || !Num<T>::is_complex) {
			*reinterpret_cast<R*>(out) = (R)a;
		} else {
			reinterpret_cast<R*>(out)[0] = (R)a;
			reinterpret_cast<R*>(out)[1] = (R)b;
		}
		__threadfence_system();
	}
}

/* 16-byte packs a thread keeps in flight per input: redInflight tuning key, default 4.  Measured on 1 GiB double
 * vectors (profiles/r2_blas1.json, fraction of the copy peak; 2 / 4 / 8 packs at 8 CTAs per SM): dot 0.98 / 1.09 / 1.08,
 * nrm2 0.84 / 0.99 / 0.80, amax 0.91 / 0.97 / 0.68 -- 4 is the one depth that serves both shapes. */
template <typename Op>
static int reduce_inflight(const SpgpuTuning* t)
{
	const int v = t->redInflight;
	if (v == 2 || v == 4 || v == 8)
		return v;
	return 4;
}

/* CTAs of one reduction over n elements when `cap` CTAs may run */
template <typename T, typename Op>
static unsigned reduce_grid(long long n, long long cap, int INFL)
{
	const long long items = (n / RPack<T>::N + INFL - 1) / INFL + 1;
	long long want = (items + RED_BLOCK - 1) / RED_BLOCK;
	if (cap > SPGPU_RED_MAX_BLOCKS) cap = SPGPU_RED_MAX_BLOCKS;
	if (cap < 1) cap = 1;
	if (want > cap) want = cap;
	if (want < 1) want = 1;
	return (unsigned)want;
}

template <typename T, typename Op>
static void reduce_dispatch(spgpuHandle_t handle, int infl, dim3 grid, const T* x, const T* y, long long n, long long pitch,
	Acc2* partials, unsigned* tickets, void* out, int outBytes, int finish, int outKind)
{
	if (infl == 8)
		spgpu_launch_dep(handle, reduce_kernel<T, Op, 8>, grid, RED_BLOCK, x, y, n, pitch, partials, tickets, out, outBytes, finish, outKind);
	else if (infl == 4)
		spgpu_launch_dep(handle, reduce_kernel<T, Op, 4>, grid, RED_BLOCK, x, y, n, pitch, partials, tickets, out, outBytes, finish, outKind);
	else
		spgpu_launch_dep(handle, reduce_kernel<T, Op, 2>, grid, RED_BLOCK, x, y, n, pitch, partials, tickets, out, outBytes, finish, outKind);
}

template <typename T, typename Op>
static void reduce_launch(spgpuHandle_t handle, const T* x, const T* y, long long n,
	void* out, int finish, int outKind)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	const SpgpuTuning* t = spgpu_tuning(handle);
	const long long cap = (long long)handle->multiProcessorCount * (t->redBlocksPerSm > 0 ? t->redBlocksPerSm : 8);
	reduce_dispatch<T, Op>(handle, reduce_inflight<Op>(t), dim3(reduce_grid<T, Op>(n, cap, reduce_inflight<Op>(t))),
		x, y, n, 0, reinterpret_cast<Acc2*>(h->dPartials), h->dTicket, out, 0, finish, outKind);
	spgpu_count_launch(handle);
}

/* blocking form: result through the handle's mapped pinned slot */
template <typename T, typename Op, typename Ret>
static Ret reduce_blocking(spgpuHandle_t handle, const T* x, const T* y, int n, int finish, int outKind)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	Ret zero;
	memset(&zero, 0, sizeof(zero));
	if (n <= 0 || h->magic != SPGPU_PRIV_MAGIC || !h->dResult)
		return zero;
	reduce_launch<T, Op>(handle, x, y, n, h->dResult, finish, outKind);
	cudaStreamSynchronize(handle->currentStream);
	Ret r;
	memcpy(&r, h->hResult, sizeof(r));
	return r;
}

/*
 * Multi-vector forms: `count` reductions over vectors `pitch` elements apart.  The reference
 * loops the blocking scalar routine (count launches + count host synchronisations, e.g.
 * reference ddot.cu:152-160); here the whole batch is ONE launch (grid.y = vector, see
 * reduce_kernel), its results land in handle-owned device scratch, and ONE copy + ONE
 * synchronisation returns them.  Batches of more than 65535 vectors go in slices of that many.
 */
template <typename T, typename Op, typename Ret>
static void reduce_many(spgpuHandle_t handle, Ret* hostOut, const T* x, const T* y, int n,
	int count, int pitch, int finish, int outKind)
{
	if (count <= 0)
		return;
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (n <= 0 || h->magic != SPGPU_PRIV_MAGIC) {
		memset(hostOut, 0, sizeof(Ret) * (size_t)count);
		return;
	}
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int slice = count < 65535 ? count : 65535;
	/* the CTAs the device holds, shared out among the vectors of a slice (at least one each) */
	const long long cap = (long long)handle->multiProcessorCount * (t->redBlocksPerSm > 0 ? t->redBlocksPerSm : 8);
	const int infl = reduce_inflight<Op>(t);
	const unsigned gx = reduce_grid<T, Op>(n, cap / slice > 0 ? cap / slice : 1, infl);
	/* scratch: [results | partials]; tickets in their own zero-kept block */
	const size_t resBytes = ((sizeof(Ret) * (size_t)count + 255) / 256) * 256;
	char* base = (char*)spgpuScratch(handle, resBytes + (size_t)slice * gx * sizeof(Acc2));
	unsigned* tickets = spgpuTickets(handle, (size_t)slice);
	if (!base || !tickets)
		return;
	Ret* dOut = reinterpret_cast<Ret*>(base);
	for (int v0 = 0; v0 < count; v0 += slice) {
		const int nv = count - v0 < slice ? count - v0 : slice;
		const long long o = (long long)v0 * pitch;
		reduce_dispatch<T, Op>(handle, infl, dim3(gx, (unsigned)nv),
			x + o, y ? y + o : (const T*)0, n, pitch, reinterpret_cast<Acc2*>(base + resBytes), tickets,
			dOut + v0, (int)sizeof(Ret), finish, outKind);
		spgpu_count_launch(handle);
	}
	cudaMemcpyAsync(hostOut, dOut, sizeof(Ret) * (size_t)count, cudaMemcpyDeviceToHost, handle->currentStream);
	cudaStreamSynchronize(handle->currentStream);
}

/* ---- C entry points -------------------------------------------------------- */

#define SPGPU_DEFINE_REDUCE(S, T, R)                                           \
	extern "C" T spgpu##S##dot(spgpuHandle_t h, int n, T* a, T* b)              \
	{ return reduce_blocking<T, OpDot<T>, T>(h, a, b, n, 0, 0); }               \
	extern "C" void spgpu##S##mdot(spgpuHandle_t h, T* y, int n, T* a, T* b,    \
		int count, int pitch)                                                   \
	{ reduce_many<T, OpDot<T>, T>(h, y, a, b, n, count, pitch, 0, 0); }                                                                           \
	extern "C" R spgpu##S##nrm2(spgpuHandle_t h, int n, T* x)                   \
	{ return reduce_blocking<T, OpSqSum<T>, R>(h, x, (const T*)0, n, 1, 1); }   \
	extern "C" void spgpu##S##mnrm2(spgpuHandle_t h, R* y, int n, T* x,         \
		int count, int pitch)                                                   \
	{ reduce_many<T, OpSqSum<T>, R>(h, y, x, (const T*)0, n, count, pitch, 1, 1); }                                                                           \
	extern "C" R spgpu##S##asum(spgpuHandle_t h, int n, T* x)                   \
	{ return reduce_blocking<T, OpAbsSum<T>, R>(h, x, (const T*)0, n, 0, 1); }  \
	extern "C" R spgpu##S##amax(spgpuHandle_t h, int n, T* x)                   \
	{ return reduce_blocking<T, OpAbsMax<T>, R>(h, x, (const T*)0, n, 0, 1); }  \
	extern "C" void spgpu##S##masum(spgpuHandle_t h, R* y, int n, T* x,         \
		int count, int pitch)                                                   \
	{ reduce_many<T, OpAbsSum<T>, R>(h, y, x, (const T*)0, n, count, pitch, 0, 1); }                                                                           \
	extern "C" void spgpu##S##mamax(spgpuHandle_t h, R* y, int n, T* x,         \
		int count, int pitch)                                                   \
	{ reduce_many<T, OpAbsMax<T>, R>(h, y, x, (const T*)0, n, count, pitch, 0, 1); }                                                                           \
	extern "C" void spgpu##S##dotDev(spgpuHandle_t h, int n, const T* a,        \
		const T* b, T* dRes)                                                    \
	{                                                                           \
		if (n > 0) reduce_launch<T, OpDot<T> >(h, a, b, n, dRes, 0, 0);         \
		else cudaMemsetAsync(dRes, 0, sizeof(T), h->currentStream);             \
	}                                                                           \
	extern "C" void spgpu##S##nrm2sqDev(spgpuHandle_t h, int n, const T* x,     \
		R* dRes)                                                                \
	{                                                                           \
		if (n > 0) reduce_launch<T, OpSqSum<T> >(h, x, (const T*)0, n, dRes, 0, 1); \
		else cudaMemsetAsync(dRes, 0, sizeof(R), h->currentStream);             \
	}

SPGPU_DEFINE_REDUCE(S, float, float)
SPGPU_DEFINE_REDUCE(D, double, double)
SPGPU_DEFINE_REDUCE(C, cuFloatComplex, float)
SPGPU_DEFINE_REDUCE(Z, cuDoubleComplex, double)

/* dRes[0] = sum x_i, result left in device memory */
#define SPGPU_DEFINE_SUMDEV(S, T, R)                                           \
	extern "C" void spgpu##S##sumDev(spgpuHandle_t h, int n, const T* x, T* dRes) \
	{                                                                           \
		if (n > 0) reduce_launch<T, OpSum<T> >(h, x, (const T*)0, n, dRes, 0, 0); \
		else cudaMemsetAsync(dRes, 0, sizeof(T), h->currentStream);             \
	}
SPGPU_FOR_FLOAT_TYPES(SPGPU_DEFINE_SUMDEV)
