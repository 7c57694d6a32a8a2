/*
 * Krylov helpers that keep their scalars on the device (include/spgpu_ext.h), for
 * all four value types: vector updates whose coefficients are quotients of device-resident scalars,
 * the fused CG update (x += a p ; r -= a Ap ; r.r), and the fold of per-row-block partials the fused
 * SpMV + dot kernels leave behind.  Where a result is a scalar of a partitioned iteration, the last
 * CTA of the kernel that produces it also all-reduces it over NVLink peer memory (peer_sync.cuh), so
 * the dependent chain of an iteration carries no separate all-reduce launches.  No reference
 * counterpart (the reference's reductions return their value to the host, e.g. reference ddot.cu:112-150).
 *
 * Complex types use the library's own UNCONJUGATED products (reference zdot.cu:54), i.e. these are the
 * building blocks of COCG for complex symmetric matrices.
 */
#include <cstring>
#include "launch.cuh"
#include "peer_sync.cuh"

ArArgs spgpu_ar_args(spgpuHandle_t handle, const spgpuPeerAllreduce* ar);

template <typename T> struct alignas(16) KPack {
	static constexpr int N = 16 / (int)sizeof(T);
	T v[N];
};

/* sign * (*num / *den), a NULL pointer meaning 1 */
template <typename T>
__device__ __forceinline__ T dev_quotient(const T* num, const T* den, double sign)
{
	T a = Num<T>::from_real((typename Num<T>::real)sign);
	if (num) a = Num<T>::mul(a, __ldg(num));
	if (den) a = Num<T>::div(a, __ldg(den));
	return a;
}

/* ---- z = b*y + a*x with a, b formed from device-resident scalars ----------- */

template <typename T>
__global__ void __launch_bounds__(256)
axpby_dev_kernel(T* z, long long n, const T* bNum, const T* bDen, double bSign, const T* y,
	const T* aNum, const T* aDen, double aSign, const T* x, int vec)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const T a = dev_quotient<T>(aNum, aDen, aSign), b = dev_quotient<T>(bNum, bDen, bSign);
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	constexpr int N = KPack<T>::N;
	if (vec) {
		const long long np = n / N;
		KPack<T>* zo = reinterpret_cast<KPack<T>*>(z);
		const KPack<T>* yp = reinterpret_cast<const KPack<T>*>(y);
		const KPack<T>* xp = reinterpret_cast<const KPack<T>*>(x);
		for (long long p = tid; p < np; p += 2 * nthreads) {
			const long long q = p + nthreads;
			const bool two = q < np;
			const long long q2 = two ? q : p;              /* both loads always issued together */
			KPack<T> y0 = yp[p], x0 = xp[p], y1 = yp[q2], x1 = xp[q2];
#pragma unroll
			for (int e = 0; e < N; ++e) {
				y0.v[e] = Num<T>::fma(b, y0.v[e], Num<T>::mul(a, x0.v[e]));
				y1.v[e] = Num<T>::fma(b, y1.v[e], Num<T>::mul(a, x1.v[e]));
			}
			zo[p] = y0;
			if (two) zo[q] = y1;
		}
		const long long done = np * N;
		if (tid < n - done)
			z[done + tid] = Num<T>::fma(b, y[done + tid], Num<T>::mul(a, x[done + tid]));
	} else {
		for (long long e = tid; e < n; e += nthreads)
			z[e] = Num<T>::fma(b, y[e], Num<T>::mul(a, x[e]));
	}
}

template <typename T>
static void axpby_dev_launch(spgpuHandle_t handle, T* z, int n, const T* dBetaNum, const T* dBetaDen,
	double betaSign, const T* y, const T* dAlphaNum, const T* dAlphaDen, double alphaSign, const T* x)
{
	if (n <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int vec = (((size_t)z | (size_t)y | (size_t)x) & 15) == 0;
	long long want = ((vec ? n / (2 * KPack<T>::N) : n) + 255) / 256 + 1;
	const long long cap = (long long)handle->multiProcessorCount * (t->vecBlocksPerSm > 0 ? t->vecBlocksPerSm : 8);
	if (cap > 0 && want > cap) want = cap;
	spgpu_launch_dep(handle, axpby_dev_kernel<T>, (unsigned)want, 256, z, n, dBetaNum, dBetaDen,
		betaSign, y, dAlphaNum, dAlphaDen, alphaSign, x, vec);
	spgpu_count_launch(handle);
}

/* ---- last CTA: total -> (all-reduce over the ranks) -> dRes ---------------------------- */

/* called by every thread of the CTA that reduce_finish elected (total valid in thread 0) */
template <typename T>
__device__ __forceinline__ void finish_scalar(Acc2 total, T* dRes, const ArArgs& ar, Acc2* smem)
{
	if (ar.world > 1) {
		if (threadIdx.x == 0)
			smem[0] = total;
		__syncthreads();
		if (threadIdx.x < 32) {
			total = peer_allreduce_sum_warp(smem[0], ar);
			if (threadIdx.x == 0)
				*dRes = from_acc2<T>(total);
		}
	} else if (threadIdx.x == 0) {
		*dRes = from_acc2<T>(total);
	}
}

/* ---- fused CG update: x += a p ; r -= a Ap ; dRes = r.r   (a = *rr / *pAp) ---- */

template <typename T>
__global__ void __launch_bounds__(256)
cg_update_kernel(T* x, T* r, const T* p, const T* ap, long long n, const T* rr, const T* pap, int vec,
	Acc2* partials, unsigned* ticket, T* dRes, const ArArgs ar)
{
	__shared__ Acc2 smem[32];
	__shared__ bool amLast;
	grid_dependency_wait();
	grid_launch_dependents();
	const T alpha = Num<T>::div(__ldg(rr), __ldg(pap));
	const T nalpha = Num<T>::mul(Num<T>::from_real(-1), alpha);
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	constexpr int N = KPack<T>::N;
	T s0 = Num<T>::zero(), s1 = Num<T>::zero();
	if (vec) {
		KPack<T>* xq = reinterpret_cast<KPack<T>*>(x);
		KPack<T>* rq = reinterpret_cast<KPack<T>*>(r);
		const KPack<T>* pq = reinterpret_cast<const KPack<T>*>(p);
		const KPack<T>* aq = reinterpret_cast<const KPack<T>*>(ap);
		const long long np = n / N;
		for (long long q = tid; q < np; q += nthreads) {
			KPack<T> xv = xq[q], rv = rq[q];
			const KPack<T> pv = pq[q], av = aq[q];
#pragma unroll
			for (int e = 0; e < N; ++e) {
				xv.v[e] = Num<T>::fma(alpha, pv.v[e], xv.v[e]);
				rv.v[e] = Num<T>::fma(nalpha, av.v[e], rv.v[e]);
				if (e & 1) s1 = Num<T>::fma(rv.v[e], rv.v[e], s1);
				else       s0 = Num<T>::fma(rv.v[e], rv.v[e], s0);
			}
			xq[q] = xv; rq[q] = rv;
		}
		const long long done = np * N;
		if (tid < n - done) {
			const long long e = done + tid;
			x[e] = Num<T>::fma(alpha, p[e], x[e]);
			const T rv = Num<T>::fma(nalpha, ap[e], r[e]);
			r[e] = rv;
			s0 = Num<T>::fma(rv, rv, s0);
		}
	} else {
		for (long long e = tid; e < n; e += nthreads) {
			x[e] = Num<T>::fma(alpha, p[e], x[e]);
			const T rv = Num<T>::fma(nalpha, ap[e], r[e]);
			r[e] = rv;
			s0 = Num<T>::fma(rv, rv, s0);
		}
	}
	Acc2 v = block_reduce<false>(to_acc2<T>(Num<T>::add(s0, s1)), smem);
	Acc2 total = { 0.0, 0.0 };
	reduce_finish<false>(v, partials, ticket, smem, &amLast, total);
	if (!amLast)
		return;
	finish_scalar<T>(total, dRes, ar, smem);
}

template <typename T>
static void cg_update_launch(spgpuHandle_t handle, T* x, T* r, const T* p, const T* ap, int n,
	const T* dRr, const T* dPAp, T* dRrNew, const spgpuPeerAllreduce* ar)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int vec = (((size_t)x | (size_t)r | (size_t)p | (size_t)ap) & 15) == 0;
	long long want = ((vec ? (n > 0 ? n : 0) / KPack<T>::N : (n > 0 ? n : 0)) + 255) / 256;
	long long cap = (long long)handle->multiProcessorCount * (t->vecBlocksPerSm > 0 ? t->vecBlocksPerSm : 8);
	if (cap > SPGPU_RED_MAX_BLOCKS) cap = SPGPU_RED_MAX_BLOCKS;
	if (want > cap) want = cap;
	if (want < 1) want = 1;
	spgpu_launch_dep(handle, cg_update_kernel<T>, (unsigned)want, 256, x, r, p, ap, n > 0 ? n : 0, dRr, dPAp, vec,
		reinterpret_cast<Acc2*>(h->dPartials), h->dTicket, dRrNew, spgpu_ar_args(handle, ar));
	spgpu_count_launch(handle);
}

/* ---- fold of the per-warp partials a fused SpMV + dot leaves behind -------------------------- */

template <typename T>
__global__ void __launch_bounds__(256)
fold_partials_kernel(const typename DotPartial<T>::type* __restrict__ in, long long n, Acc2* partials, unsigned* ticket,
	T* dRes, const ArArgs ar)
{
	__shared__ Acc2 smem[32];
	__shared__ bool amLast;
	grid_dependency_wait();
	grid_launch_dependents();
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	Acc2 s = { 0.0, 0.0 };
	for (long long e = tid; e < n; e += 4 * nthreads) {
		Acc2 v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const long long q = e + u * nthreads;
			v[u] = q < n ? ld_partial(in + q) : Acc2{ 0.0, 0.0 };
		}
		s.a += (v[0].a + v[1].a) + (v[2].a + v[3].a);
		s.b += (v[0].b + v[1].b) + (v[2].b + v[3].b);
	}
	Acc2 v = block_reduce<false>(s, smem);
	Acc2 total = { 0.0, 0.0 };
	reduce_finish<false>(v, partials, ticket, smem, &amLast, total);
	if (!amLast)
		return;
	finish_scalar<T>(total, dRes, ar, smem);
}

template <typename T>
void spgpu_fold_partials(spgpuHandle_t handle, const typename DotPartial<T>::type* partials, long long n, T* dRes,
	const spgpuPeerAllreduce* ar)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	long long want = (n + 4 * 256 - 1) / (4 * 256);
	long long cap = (long long)handle->multiProcessorCount * 8;
	if (cap > SPGPU_RED_MAX_BLOCKS) cap = SPGPU_RED_MAX_BLOCKS;
	if (want > cap) want = cap;
	if (want < 1) want = 1;
	spgpu_launch_dep(handle, fold_partials_kernel<T>, (unsigned)want, 256, partials, n > 0 ? n : 0,
		reinterpret_cast<Acc2*>(h->dPartials), h->dTicket, dRes, spgpu_ar_args(handle, ar));
	spgpu_count_launch(handle);
}

template void spgpu_fold_partials<float>(spgpuHandle_t, const double*, long long, float*, const spgpuPeerAllreduce*);
template void spgpu_fold_partials<double>(spgpuHandle_t, const double*, long long, double*, const spgpuPeerAllreduce*);
template void spgpu_fold_partials<cuFloatComplex>(spgpuHandle_t, const Acc2*, long long, cuFloatComplex*, const spgpuPeerAllreduce*);
template void spgpu_fold_partials<cuDoubleComplex>(spgpuHandle_t, const Acc2*, long long, cuDoubleComplex*, const spgpuPeerAllreduce*);

/* ---- C entry points ------------------------------------------------------------------- */

#define SPGPU_DEFINE_KRYLOV(S, T, R)                                                            \
	extern "C" void spgpu##S##axpbyDev(spgpuHandle_t handle, T* z, int n, const T* dBetaNum,       \
		const T* dBetaDen, double betaSign, const T* y, const T* dAlphaNum, const T* dAlphaDen,    \
		double alphaSign, const T* x)                                                              \
	{                                                                                              \
		axpby_dev_launch<T>(handle, z, n, dBetaNum, dBetaDen, betaSign, y, dAlphaNum, dAlphaDen,   \
			alphaSign, x);                                                                         \
	}                                                                                              \
	extern "C" void spgpu##S##cgUpdateDev(spgpuHandle_t handle, T* x, T* r, const T* p,            \
		const T* ap, int n, const T* dRr, const T* dPAp, T* dRrNew, const spgpuPeerAllreduce* ar)  \
	{                                                                                              \
		cg_update_launch<T>(handle, x, r, p, ap, n, dRr, dPAp, dRrNew, ar);                        \
	}
SPGPU_FOR_FLOAT_TYPES(SPGPU_DEFINE_KRYLOV)

/* see spgpuPreloadHaloKernels (ext_halo.cu): the kernels of this file whose last CTA may wait for a peer */
template <typename T>
static int preload_krylov_type()
{
	cudaFuncAttributes a;
	int bad = 0;
	bad += cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(axpby_dev_kernel<T>)) != cudaSuccess;
	bad += cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(cg_update_kernel<T>)) != cudaSuccess;
	bad += cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(fold_partials_kernel<T>)) != cudaSuccess;
	return bad;
}

extern "C" int spgpuPreloadKrylovKernels(void)
{
	return (preload_krylov_type<float>() + preload_krylov_type<double>() + preload_krylov_type<cuFloatComplex>()
		+ preload_krylov_type<cuDoubleComplex>()) ? -1 : 0;
}
