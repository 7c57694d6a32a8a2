/*
 * Additive entry points (include/spgpu_ext.h): Krylov helpers that keep their
 * scalars on the device, the fused HELL SpMV + dot, and the NVLink halo push
 * used by the row-partitioned multi-GPU layer.  No reference counterpart.
 */
#include <cstdio>
#include <cstring>
#include "launch.cuh"
#include "reduce_common.cuh"
#include "spmv_hell_body.cuh"
#include "spmv_hdia_body.cuh"

/* ---- z = b*y + a*x with a, b formed from device-resident scalars ----------- */

__global__ void __launch_bounds__(256)
daxpby_dev_kernel(double* z, long long n, const double* bNum, const double* bDen,
	double bSign, const double* y, const double* aNum, const double* aDen,
	double aSign, const double* x, int vec)
{
	double a = aSign, b = bSign;
	if (aNum) a *= __ldg(aNum);
	if (aDen) a /= __ldg(aDen);
	if (bNum) b *= __ldg(bNum);
	if (bDen) b /= __ldg(bDen);
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	if (vec) {
		const long long np = n >> 1;
		double2* zo = reinterpret_cast<double2*>(z);
		const double2* y2 = reinterpret_cast<const double2*>(y);
		const double2* x2 = reinterpret_cast<const double2*>(x);
		for (long long p = tid; p < np; p += 2 * nthreads) {
			const long long q = p + nthreads;
			const bool two = q < np;
			const long long q2 = two ? q : p;              /* both loads always issued together */
			double2 y0 = y2[p], x0 = x2[p], y1 = y2[q2], x1 = x2[q2];
			zo[p] = make_double2(fma(b, y0.x, a * x0.x), fma(b, y0.y, a * x0.y));
			if (two) zo[q] = make_double2(fma(b, y1.x, a * x1.x), fma(b, y1.y, a * x1.y));
		}
		if (tid == 0 && (n & 1))
			z[n - 1] = fma(b, y[n - 1], a * x[n - 1]);
	} else {
		for (long long e = tid; e < n; e += nthreads)
			z[e] = fma(b, y[e], a * x[e]);
	}
}

extern "C" void spgpuDaxpbyDev(spgpuHandle_t handle, double* z, int n,
	const double* dBetaNum, const double* dBetaDen, double betaSign,
	const double* y, const double* dAlphaNum, const double* dAlphaDen,
	double alphaSign, const double* x)
{
	if (n <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int vec = (((size_t)z | (size_t)y | (size_t)x) & 15) == 0;
	long long want = ((vec ? n / 4 : n) + 255) / 256 + 1;
	const long long cap = (long long)handle->multiProcessorCount * (t->vecBlocksPerSm > 0 ? t->vecBlocksPerSm : 8);
	if (cap > 0 && want > cap) want = cap;
	daxpby_dev_kernel<<<(unsigned)want, 256, 0, handle->currentStream>>>(z, n, dBetaNum, dBetaDen,
		betaSign, y, dAlphaNum, dAlphaDen, alphaSign, x, vec);
	spgpu_count_launch(handle);
}

/* ---- HELL SpMV fused with p.Ap ---------------------------------------------- */

/*
 * The direct HELL kernel with one extra step per row: z_i * x[xOffset+i] is summed over
 * the CTA (shuffle + 4 shared doubles, fixed order) and stored as ONE partial per CTA
 * into handle-owned scratch -- no atomics, no tickets (a million CTAs hammering one
 * address cost more than the dot they save).  spgpuDsumDev then folds the partials
 * deterministically; that second launch reads 8 MB where a separate dot would re-read
 * two 1 GB vectors.
 */
template <int UNROLL, int HACK, int MINB>
__global__ void __launch_bounds__(128, MINB)
dhell_spmv_dot_kernel(const HellArgs<double> a, int xOffset, double* __restrict__ ctaPartials)
{
	__shared__ double ws[4];
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31;
	double zval;
	hell_warp_rows_value<double, UNROLL, HACK>(a, i - lane, zval);
	double contrib = 0.0;
	if (i < (unsigned)a.rows)
		contrib = zval * __ldg(a.x + xOffset + i);
	contrib = warp_sum<double>(contrib);
	if (lane == 0)
		ws[threadIdx.x >> 5] = contrib;
	__syncthreads();
	if (threadIdx.x == 0)
		ctaPartials[blockIdx.x] = (ws[0] + ws[1]) + (ws[2] + ws[3]);
}

extern "C" void spgpuDsumDev(spgpuHandle_t h, int n, const double* x, double* dRes);

extern "C" void spgpuDhellspmvDot(spgpuHandle_t handle, double* z, const double* cM,
	const int* rP, int hackSize, const int* hackOffsets, const int* rS, int rows,
	const double* x, int baseIndex, int xOffset, double* dRes)
{
	if (rows <= 0) {
		cudaMemsetAsync(dRes, 0, sizeof(double), handle->currentStream);
		return;
	}
	const SpgpuTuning* t = spgpu_tuning(handle);
	const unsigned grid = spgpu_ceil_div(rows, 128);
	double* partials = (double*)spgpuScratch(handle, (size_t)grid * sizeof(double));
	if (!partials)
		return;
	const HellArgs<double> a = { z, NULL, 1.0, cM, rP, hackSize, hackOffsets, rS, NULL, rows, x, 0.0,
		baseIndex, spgpu_long_cut(t, 8), t->hellVariant != 1, 0, NULL, NULL, 0 };
	if (hackSize == 32)
		dhell_spmv_dot_kernel<8, 32, 10><<<grid, 128, 0, handle->currentStream>>>(a, xOffset, partials);
	else
		dhell_spmv_dot_kernel<8, 0, 8><<<grid, 128, 0, handle->currentStream>>>(a, xOffset, partials);
	spgpu_count_launch(handle);
	spgpuDsumDev(handle, (int)grid, partials, dRes);
}

/* ---- fused CG update: x += a p ; r -= a Ap ; dRes = r.r   (a = *rr / *pAp) ---- */

__global__ void __launch_bounds__(256)
dcg_update_kernel(double* x, double* r, const double* p, const double* ap, long long n,
	const double* rr, const double* pap, int vec, Acc2* partials, unsigned* ticket, double* dRes)
{
	__shared__ Acc2 smem[32];
	__shared__ bool amLast;
	const double alpha = __ldg(rr) / __ldg(pap);
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	double s0 = 0.0, s1 = 0.0;
	if (vec) {
		double2* x2 = reinterpret_cast<double2*>(x);
		double2* r2 = reinterpret_cast<double2*>(r);
		const double2* p2 = reinterpret_cast<const double2*>(p);
		const double2* a2 = reinterpret_cast<const double2*>(ap);
		const long long np = n >> 1;
		for (long long q = tid; q < np; q += nthreads) {
			double2 xv = x2[q], rv = r2[q];
			const double2 pv = p2[q], av = a2[q];
			xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
			rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
			x2[q] = xv; r2[q] = rv;
			s0 = fma(rv.x, rv.x, s0); s1 = fma(rv.y, rv.y, s1);
		}
		if (tid == 0 && (n & 1)) {
			const long long e = n - 1;
			x[e] = fma(alpha, p[e], x[e]);
			const double rv = fma(-alpha, ap[e], r[e]);
			r[e] = rv;
			s0 = fma(rv, rv, s0);
		}
	} else {
		for (long long e = tid; e < n; e += nthreads) {
			x[e] = fma(alpha, p[e], x[e]);
			const double rv = fma(-alpha, ap[e], r[e]);
			r[e] = rv;
			s0 = fma(rv, rv, s0);
		}
	}
	Acc2 v = block_reduce<false>(Acc2{ s0 + s1, 0.0 }, smem);
	Acc2 total;
	if (reduce_finish<false>(v, partials, ticket, smem, &amLast, total))
		*dRes = total.a;
}

extern "C" void spgpuDcgUpdateDev(spgpuHandle_t handle, double* x, double* r, const double* p,
	const double* ap, int n, const double* dRr, const double* dPAp, double* dRrNew)
{
	if (n <= 0) {
		cudaMemsetAsync(dRrNew, 0, sizeof(double), handle->currentStream);
		return;
	}
	SpgpuHandlePriv* h = spgpuPriv(handle);
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int vec = (((size_t)x | (size_t)r | (size_t)p | (size_t)ap) & 15) == 0;
	long long want = ((vec ? n / 2 : n) + 255) / 256;
	long long cap = (long long)handle->multiProcessorCount * (t->vecBlocksPerSm > 0 ? t->vecBlocksPerSm : 8);
	if (cap > SPGPU_RED_MAX_BLOCKS) cap = SPGPU_RED_MAX_BLOCKS;
	if (want > cap) want = cap;
	if (want < 1) want = 1;
	dcg_update_kernel<<<(unsigned)want, 256, 0, handle->currentStream>>>(x, r, p, ap, n, dRr, dPAp, vec,
		reinterpret_cast<Acc2*>(h->dPartials), h->dTicket, dRrNew);
	spgpu_count_launch(handle);
}

/* ---- CUDA IPC + raw allocation ---------------------------------------------- */

extern "C" int spgpuIpcGetHandle(void* devPtr, void* handle64)
{
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, devPtr);
	if (e != cudaSuccess)
		return (int)e;
	memcpy(handle64, &h, sizeof(h));
	return 0;
}

extern "C" int spgpuIpcOpenHandle(const void* handle64, void** devPtr)
{
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	return (int)cudaIpcOpenMemHandle(devPtr, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int spgpuIpcCloseHandle(void* devPtr)
{
	return (int)cudaIpcCloseMemHandle(devPtr);
}

extern "C" int spgpuDeviceAlloc(void** devPtr, size_t bytes)
{
	return (int)cudaMalloc(devPtr, bytes);
}

extern "C" int spgpuDeviceFree(void* devPtr)
{
	return (int)cudaFree(devPtr);
}

/* ---- halo push over NVLink --------------------------------------------------- */

/*
 * Copies src[0..n) into a peer GPU's memory with 128-bit stores; the last CTA
 * to finish (ticket in local memory) makes the data visible system-wide and
 * release-stores flagValue into the peer's flag word.
 */
__global__ void __launch_bounds__(256)
halo_push_kernel(double* peerDst, const double* __restrict__ src, long long n, int vec,
	unsigned* peerFlag, unsigned flagValue, unsigned* ticket)
{
	__shared__ bool amLast;
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	if (vec) {
		const long long np = n >> 1;
		double2* d2 = reinterpret_cast<double2*>(peerDst);
		const double2* s2 = reinterpret_cast<const double2*>(src);
		for (long long p = tid; p < np; p += nthreads)
			d2[p] = s2[p];
		if (tid == 0 && (n & 1))
			peerDst[n - 1] = src[n - 1];
	} else {
		for (long long e = tid; e < n; e += nthreads)
			peerDst[e] = src[e];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned t = atomicAdd(ticket, 1u);
		amLast = (t == gridDim.x - 1);
	}
	__syncthreads();
	if (amLast && threadIdx.x == 0) {
		*ticket = 0u;
		if (peerFlag) {
			__threadfence_system();
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(peerFlag), "r"(flagValue) : "memory");
		}
	}
}

extern "C" void spgpuDhaloPush(spgpuHandle_t handle, double* peerDst, const double* src,
	int n, unsigned* peerFlag, unsigned flagValue)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	const int vec = (((size_t)peerDst | (size_t)src) & 15) == 0;
	long long want = ((vec ? n / 2 : n) + 255) / 256;
	if (want > 2 * handle->multiProcessorCount) want = 2 * handle->multiProcessorCount;
	if (want < 1) want = 1;
	/* the ticket word next to the reductions' one (offset 16 bytes) */
	halo_push_kernel<<<(unsigned)want, 256, 0, handle->currentStream>>>(peerDst, src, n > 0 ? n : 0,
		vec, peerFlag, flagValue, h->dTicket + 4);
	spgpu_count_launch(handle);
}

/* ---- fused halo exchange: wait acks -> push both planes -> signal -> wait arrivals ---- */

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

/* spin until *flag >= value; gives up (with a message) after timeoutNs so a bug is a wrong
 * answer the tests catch, never a hung GPU */
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned value, unsigned long long timeoutNs)
{
	unsigned long long t0, t1;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for (;;) {
		const unsigned v = ld_acquire_sys(flag);
		if ((int)(v - value) >= 0)
			return;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (t1 - t0 > timeoutNs) {
			printf("spgpu halo: timed out waiting for flag value %u (saw %u)\n", value, v);
			return;
		}
		__nanosleep(100);
	}
}

/*
 * One kernel per SpMV on each rank.  Even CTAs copy this rank's first n owned entries into
 * the LOWER neighbour's upper halo zone, odd CTAs its last n owned entries into the UPPER
 * neighbour's lower halo zone (128-bit stores through NVLink peer pointers).  Before
 * copying, a CTA waits until that neighbour has acknowledged the previous halo (so it is
 * not overwritten while still being read).  The last CTA to finish release-stores the
 * sequence number into both neighbours' "ready" flags and then waits for this rank's own
 * two "ready" flags, so the kernel completes exactly when this rank's halos have arrived.
 */
__global__ void __launch_bounds__(256)
halo_exchange_kernel(double* dstLo, const double* srcLo, double* dstHi, const double* srcHi,
	long long n, const unsigned* ackLo, const unsigned* ackHi, unsigned* peerReadyLo,
	unsigned* peerReadyHi, const unsigned* myReadyLo, const unsigned* myReadyHi,
	unsigned seq, unsigned* ticket, unsigned long long timeoutNs)
{
	__shared__ bool amLast;
	const bool toHi = (blockIdx.x & 1) != 0;
	double* dst = toHi ? dstHi : dstLo;
	const double* src = toHi ? srcHi : srcLo;
	const unsigned* ack = toHi ? ackHi : ackLo;
	if (dst) {
		if (threadIdx.x == 0 && ack && seq > 1)
			spin_until(ack, seq - 1, timeoutNs);
		__syncthreads();
		const long long half = gridDim.x >> 1;
		const long long tid = (long long)(blockIdx.x >> 1) * blockDim.x + threadIdx.x;
		const long long nthreads = half * blockDim.x;
		if ((((size_t)dst | (size_t)src) & 15) == 0) {
			double2* d2 = reinterpret_cast<double2*>(dst);
			const double2* s2 = reinterpret_cast<const double2*>(src);
			for (long long p = tid; p < (n >> 1); p += nthreads)
				d2[p] = s2[p];
			if (tid == 0 && (n & 1))
				dst[n - 1] = src[n - 1];
		} else {
			for (long long e = tid; e < n; e += nthreads)
				dst[e] = src[e];
		}
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0)
		amLast = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	__syncthreads();
	if (amLast && threadIdx.x == 0) {
		*ticket = 0u;
		__threadfence_system();
		if (peerReadyLo) st_release_sys(peerReadyLo, seq);
		if (peerReadyHi) st_release_sys(peerReadyHi, seq);
		if (myReadyLo) spin_until(myReadyLo, seq, timeoutNs);
		if (myReadyHi) spin_until(myReadyHi, seq, timeoutNs);
	}
}

extern "C" void spgpuDhaloExchange(spgpuHandle_t handle, double* peerDstLo, const double* srcLo,
	double* peerDstHi, const double* srcHi, int n, const unsigned* ackLo, const unsigned* ackHi,
	unsigned* peerReadyLo, unsigned* peerReadyHi, const unsigned* myReadyLo,
	const unsigned* myReadyHi, unsigned seq)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	long long want = 2 * (((long long)n / 2 + 255) / 256);
	const long long cap = 2LL * (handle->multiProcessorCount / 2 > 0 ? handle->multiProcessorCount / 2 : 1);
	if (want > cap) want = cap;
	if (want < 2) want = 2;
	halo_exchange_kernel<<<(unsigned)want, 256, 0, handle->currentStream>>>(peerDstLo, srcLo, peerDstHi, srcHi,
		n > 0 ? n : 0, ackLo, ackHi, peerReadyLo, peerReadyHi, myReadyLo, myReadyHi, seq,
		h->dTicket + 8, 2000000000ull);
	spgpu_count_launch(handle);
}

__global__ void halo_ack_kernel(unsigned* peerAckLo, unsigned* peerAckHi, unsigned seq)
{
	__threadfence_system();
	if (peerAckLo) st_release_sys(peerAckLo, seq);
	if (peerAckHi) st_release_sys(peerAckHi, seq);
}

extern "C" void spgpuHaloAck(spgpuHandle_t handle, unsigned* peerAckLo, unsigned* peerAckHi, unsigned seq)
{
	halo_ack_kernel<<<1, 1, 0, handle->currentStream>>>(peerAckLo, peerAckHi, seq);
	spgpu_count_launch(handle);
}

/* ---- HELL / HDIA SpMV fused with the halo exchange: ONE kernel per partitioned SpMV ---- */

struct HaloArgs {
	double* dstLo; const double* srcLo;      /* my first n owned entries -> lower neighbour's upper halo */
	double* dstHi; const double* srcHi;      /* my last n owned entries  -> upper neighbour's lower halo */
	int n;
	const unsigned* ackLo; const unsigned* ackHi;        /* local : neighbour consumed my previous halo   */
	unsigned* peerReadyLo; unsigned* peerReadyHi;        /* remote: my halo for `seq` is in place         */
	const unsigned* myReadyLo; const unsigned* myReadyHi;/* local : neighbour's halo for `seq` is in place */
	unsigned* peerAckLo; unsigned* peerAckHi;            /* remote: I have consumed their halo for `seq`   */
	unsigned seq;                            /* sequence number of this exchange, or ... */
	const unsigned* seqPtr;                  /* ... (seq == 0) device counter of COMPLETED exchanges: this one is *seqPtr + 1 */
	unsigned* pushTicket; unsigned* doneTicket;
	int pushCtas;
	int headBlocks, tailBlocks;              /* 128-row blocks that read the lower / upper halo zone */
	int fillerBlocks;                        /* interior blocks scheduled behind the boundary blocks (about three waves) */
	unsigned long long timeoutNs;
};

/*
 * grid = pushCtas + ceil(rows/128).  The first pushCtas CTAs move the two boundary planes
 * into the neighbours' halo zones over NVLink (after the neighbours acknowledged the
 * previous ones) and publish `seq`.  Every other CTA multiplies 128 rows; the CTAs are
 * numbered so that most interior row blocks come first and the blocks that read a halo zone
 * come late (followed only by a few waves of interior blocks) -- by the time the hardware
 * schedules them the neighbours' planes have long arrived, and if not they spin on the local
 * ready flag.  The last CTA to finish tells the
 * neighbours that their halo data has been consumed.  Transfer and multiply overlap inside
 * one launch; there is no separate pack / exchange / wait / ack kernel.
 */
/* z_i * x[xOffset+i] summed over the CTA, one partial per ROW BLOCK (fixed order) */
__device__ __forceinline__ void cta_dot_partial(double contrib, double* ctaPartials, unsigned slot)
{
	__shared__ double ws[4];
	contrib = warp_sum<double>(contrib);
	if ((threadIdx.x & 31) == 0)
		ws[threadIdx.x >> 5] = contrib;
	__syncthreads();
	if (threadIdx.x == 0)
		ctaPartials[slot] = (ws[0] + ws[1]) + (ws[2] + ws[3]);
}

/* what a row block of the fused kernel multiplies: the HELL or the HDIA warp body */
template <int UNROLL, int HACK>
struct HellRowBody {
	HellArgs<double> a;
	__device__ __forceinline__ int rows() const { return a.rows; }
	__device__ __forceinline__ const double* x() const { return a.x; }
	__device__ __forceinline__ double run(unsigned warpRow) const
	{
		double zval;
		hell_warp_rows_value<double, UNROLL, HACK>(a, warpRow, zval);
		return zval;
	}
};

template <int UNROLL, int HACK>
struct HdiaRowBody {
	HdiaArgs<double> a;
	__device__ __forceinline__ int rows() const { return a.rows; }
	__device__ __forceinline__ const double* x() const { return a.x; }
	__device__ __forceinline__ double run(unsigned warpRow) const
	{
		double zval;
		hdia_warp_rows_value<double, UNROLL, HACK, false>(a, warpRow, zval);
		return zval;
	}
};

template <typename Body, int MINB, bool DOT>
__global__ void __launch_bounds__(128, MINB)
spmv_halo_kernel(const Body body, const HaloArgs hx, int xOffset, double* __restrict__ ctaPartials)
{
	/* the counter is advanced by a later kernel in stream order (spgpuHaloSeqAdvance), never during this
	 * one, so every CTA reads the same value whenever it is scheduled */
	const unsigned seq = hx.seqPtr ? *reinterpret_cast<const volatile unsigned*>(hx.seqPtr) + 1u : hx.seq;
	if (blockIdx.x < (unsigned)hx.pushCtas) {
		const bool toHi = (blockIdx.x & 1) != 0;
		double* dst = toHi ? hx.dstHi : hx.dstLo;
		const double* src = toHi ? hx.srcHi : hx.srcLo;
		const unsigned* ack = toHi ? hx.ackHi : hx.ackLo;
		if (dst) {
			if (threadIdx.x == 0 && ack && seq > 1)
				spin_until(ack, seq - 1, hx.timeoutNs);
			__syncthreads();
			const long long half = hx.pushCtas >> 1;
			const long long tid = (long long)(blockIdx.x >> 1) * blockDim.x + threadIdx.x;
			const long long nthreads = half * blockDim.x;
			if ((((size_t)dst | (size_t)src) & 15) == 0) {
				double2* d2 = reinterpret_cast<double2*>(dst);
				const double2* s2 = reinterpret_cast<const double2*>(src);
				for (long long p = tid; p < (hx.n >> 1); p += nthreads)
					d2[p] = s2[p];
				if (tid == 0 && (hx.n & 1))
					dst[hx.n - 1] = src[hx.n - 1];
			} else {
				for (long long e = tid; e < hx.n; e += nthreads)
					dst[e] = src[e];
			}
		}
		__threadfence_system();
		__syncthreads();
		if (threadIdx.x == 0) {
			if (atomicAdd(hx.pushTicket, 1u) == (unsigned)hx.pushCtas - 1u) {
				*hx.pushTicket = 0u;
				__threadfence_system();
				if (hx.peerReadyLo) st_release_sys(hx.peerReadyLo, seq);
				if (hx.peerReadyHi) st_release_sys(hx.peerReadyHi, seq);
			}
		}
	} else {
		const unsigned b = blockIdx.x - hx.pushCtas;
		const unsigned rowBlocks = ((unsigned)body.rows() + 127u) >> 7;
		const unsigned head = min((unsigned)hx.headBlocks, rowBlocks);
		const unsigned tail = min((unsigned)hx.tailBlocks, rowBlocks - head);
		const unsigned interior = rowBlocks - head - tail;
		/* a few waves of interior blocks go BEHIND the boundary blocks: a boundary block ends with a barrier,
		 * a fence and a ticket, and as the very last CTAs of the grid nothing would overlap that latency */
		const unsigned filler = min(interior >> 2, (unsigned)hx.fillerBlocks);
		const unsigned early = interior - filler;
		unsigned rb;
		if (b < early) rb = head + b;                                          /* most of the interior first   */
		else if (b < early + head) rb = b - early;                             /* then the lower boundary      */
		else if (b < early + head + tail) rb = rowBlocks - tail + (b - early - head);   /* the upper boundary */
		else rb = head + early + (b - early - head - tail);                    /* the rest of the interior     */
		const bool needLo = rb < head, needHi = rb >= rowBlocks - tail;
		const unsigned myRow = rb * 128u + threadIdx.x;
		if (!needLo && !needHi) {
			/* interior: no flags, no tickets -- exactly the plain kernel */
			const double zval = body.run(rb * 128u + (threadIdx.x & ~31u));
			if (DOT)
				cta_dot_partial(myRow < (unsigned)body.rows() ? zval * __ldg(body.x() + xOffset + myRow) : 0.0, ctaPartials, rb);
			return;
		}
		if (threadIdx.x == 0) {
			if (needLo && hx.myReadyLo) spin_until(hx.myReadyLo, seq, hx.timeoutNs);
			if (needHi && hx.myReadyHi) spin_until(hx.myReadyHi, seq, hx.timeoutNs);
		}
		__syncthreads();
		const double zval = body.run(rb * 128u + (threadIdx.x & ~31u));
		if (DOT)
			cta_dot_partial(myRow < (unsigned)body.rows() ? zval * __ldg(body.x() + xOffset + myRow) : 0.0, ctaPartials, rb);
		/* the last CTA that read a halo zone tells that neighbour its data has been consumed
		 * (only the few boundary CTAs touch these counters) */
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			if (needLo && atomicAdd(hx.doneTicket, 1u) == head - 1u) {
				*hx.doneTicket = 0u;
				__threadfence_system();
				if (hx.peerAckLo) st_release_sys(hx.peerAckLo, seq);
			}
			if (needHi && atomicAdd(hx.doneTicket + 1, 1u) == tail - 1u) {
				*(hx.doneTicket + 1) = 0u;
				__threadfence_system();
				if (hx.peerAckHi) st_release_sys(hx.peerAckHi, seq);
			}
		}
	}
}

/* flag words (spgpu_ext.h): [0] ready-from-below [1] ready-from-above [2] ack-from-below [3] ack-from-above */
static HaloArgs halo_args(spgpuHandle_t handle, double* xExt, int rows, int haloN, double* peerXLoUpperHalo,
	double* peerXHiLowerHalo, unsigned* myFlags, unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	HaloArgs hx;
	hx.dstLo = peerXLoUpperHalo; hx.srcLo = xExt + haloN;
	hx.dstHi = peerXHiLowerHalo; hx.srcHi = xExt + rows;        /* last haloN owned entries */
	hx.n = haloN;
	hx.ackLo = peerFlagsLo ? myFlags + 2 : NULL;   hx.ackHi = peerFlagsHi ? myFlags + 3 : NULL;
	hx.peerReadyLo = peerFlagsLo ? peerFlagsLo + 1 : NULL;  hx.peerReadyHi = peerFlagsHi ? peerFlagsHi + 0 : NULL;
	hx.myReadyLo = peerFlagsLo ? myFlags + 0 : NULL;  hx.myReadyHi = peerFlagsHi ? myFlags + 1 : NULL;
	hx.peerAckLo = peerFlagsLo ? peerFlagsLo + 3 : NULL;  hx.peerAckHi = peerFlagsHi ? peerFlagsHi + 2 : NULL;
	hx.seq = seq;
	hx.seqPtr = (seq == 0u && h->magic == SPGPU_PRIV_MAGIC) ? h->dHaloSeq : NULL;
	hx.pushTicket = h->dTicket + 8;
	hx.doneTicket = h->dTicket + 12;
	hx.pushCtas = 8;
	hx.headBlocks = peerFlagsLo ? (haloN + 127) / 128 : 0;
	hx.tailBlocks = peerFlagsHi ? (haloN + 127) / 128 : 0;
	hx.fillerBlocks = 3 * 10 * handle->multiProcessorCount;
	hx.timeoutNs = 2000000000ull;
	return hx;
}

static void dhell_spmv_halo_launch(spgpuHandle_t handle, double* z, const double* y, double alpha,
	const double* cM, const int* rP, int hackSize, const int* hackOffsets, const int* rS,
	int avgNnzPerRow, int rows, double* xExt, double beta, int baseIndex, int haloN,
	double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq, double* ctaPartials)
{
	const SpgpuTuning* t = spgpu_tuning(handle);
	const HellArgs<double> a = { z, y, alpha, cM, rP, hackSize, hackOffsets, rS, NULL, rows, xExt, beta,
		baseIndex, spgpu_long_cut(t, avgNnzPerRow), t->hellVariant != 1, 0, NULL, NULL, 0 };
	const HaloArgs hx = halo_args(handle, xExt, rows, haloN, peerXLoUpperHalo, peerXHiLowerHalo, myFlags,
		peerFlagsLo, peerFlagsHi, seq);
	const unsigned grid = hx.pushCtas + spgpu_ceil_div(rows, 128);
	cudaStream_t s = handle->currentStream;
	const HellRowBody<8, 32> b32 = { a };
	const HellRowBody<8, 0> b0 = { a };
	if (ctaPartials) {
		if (hackSize == 32)
			spmv_halo_kernel<HellRowBody<8, 32>, 10, true><<<grid, 128, 0, s>>>(b32, hx, haloN, ctaPartials);
		else
			spmv_halo_kernel<HellRowBody<8, 0>, 8, true><<<grid, 128, 0, s>>>(b0, hx, haloN, ctaPartials);
	} else {
		if (hackSize == 32)
			spmv_halo_kernel<HellRowBody<8, 32>, 10, false><<<grid, 128, 0, s>>>(b32, hx, haloN, ctaPartials);
		else
			spmv_halo_kernel<HellRowBody<8, 0>, 8, false><<<grid, 128, 0, s>>>(b0, hx, haloN, ctaPartials);
	}
	spgpu_count_launch(handle);
}

extern "C" void spgpuDhellspmvHalo(spgpuHandle_t handle, double* z, const double* y, double alpha,
	const double* cM, const int* rP, int hackSize, const int* hackOffsets, const int* rS,
	int avgNnzPerRow, int rows, double* xExt, double beta, int baseIndex, int haloN,
	double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq)
{
	if (rows <= 0)
		return;
	dhell_spmv_halo_launch(handle, z, y, alpha, cM, rP, hackSize, hackOffsets, rS, avgNnzPerRow, rows, xExt,
		beta, baseIndex, haloN, peerXLoUpperHalo, peerXHiLowerHalo, myFlags, peerFlagsLo, peerFlagsHi, seq, NULL);
}

/* z = A*xExt with the halo exchange inside, plus dRes[0] = sum_i xExt[haloN+i]*z[i] (this rank's
 * share of p.Ap): per-row-block partials in handle scratch, folded by spgpuDsumDev. */
extern "C" void spgpuDhellspmvHaloDot(spgpuHandle_t handle, double* z, const double* cM, const int* rP,
	int hackSize, const int* hackOffsets, const int* rS, int avgNnzPerRow, int rows, double* xExt,
	int baseIndex, int haloN, double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq, double* dRes)
{
	if (rows <= 0) {
		cudaMemsetAsync(dRes, 0, sizeof(double), handle->currentStream);
		return;
	}
	const unsigned rowBlocks = spgpu_ceil_div(rows, 128);
	double* partials = (double*)spgpuScratch(handle, (size_t)rowBlocks * sizeof(double));
	if (!partials)
		return;
	dhell_spmv_halo_launch(handle, z, NULL, 1.0, cM, rP, hackSize, hackOffsets, rS, avgNnzPerRow, rows, xExt,
		0.0, baseIndex, haloN, peerXLoUpperHalo, peerXHiLowerHalo, myFlags, peerFlagsLo, peerFlagsHi, seq, partials);
	spgpuDsumDev(handle, (int)rowBlocks, partials, dRes);
}

/*
 * HDIA twin of spgpuDhellspmvHalo: z = alpha*A*xExt + beta*y for a row block in HDIA layout whose
 * offsets address xExt = [halo | owned | halo] (mg.split_hdia: global offset + haloN, cols = the
 * length of xExt), with the same halo protocol inside the launch.
 */
static void dhdia_spmv_halo_launch(spgpuHandle_t handle, double* z, const double* y, double alpha,
	const double* dM, const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols,
	double* xExt, double beta, int haloN, double* peerXLoUpperHalo, double* peerXHiLowerHalo,
	unsigned* myFlags, unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq, double* ctaPartials)
{
	const HdiaArgs<double> a = { z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, xExt, beta };
	const HaloArgs hx = halo_args(handle, xExt, rows, haloN, peerXLoUpperHalo, peerXHiLowerHalo, myFlags,
		peerFlagsLo, peerFlagsHi, seq);
	const unsigned grid = hx.pushCtas + spgpu_ceil_div(rows, 128);
	cudaStream_t s = handle->currentStream;
	const HdiaRowBody<9, 32> b32 = { a };
	const HdiaRowBody<9, 0> b0 = { a };
	if (ctaPartials) {
		if (hackSize == 32)
			spmv_halo_kernel<HdiaRowBody<9, 32>, 8, true><<<grid, 128, 0, s>>>(b32, hx, haloN, ctaPartials);
		else
			spmv_halo_kernel<HdiaRowBody<9, 0>, 8, true><<<grid, 128, 0, s>>>(b0, hx, haloN, ctaPartials);
	} else {
		if (hackSize == 32)
			spmv_halo_kernel<HdiaRowBody<9, 32>, 8, false><<<grid, 128, 0, s>>>(b32, hx, haloN, ctaPartials);
		else
			spmv_halo_kernel<HdiaRowBody<9, 0>, 8, false><<<grid, 128, 0, s>>>(b0, hx, haloN, ctaPartials);
	}
	spgpu_count_launch(handle);
}

extern "C" void spgpuDhdiaspmvHalo(spgpuHandle_t handle, double* z, const double* y, double alpha,
	const double* dM, const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols,
	double* xExt, double beta, int haloN, double* peerXLoUpperHalo, double* peerXHiLowerHalo,
	unsigned* myFlags, unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq)
{
	if (rows <= 0)
		return;
	dhdia_spmv_halo_launch(handle, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, xExt, beta, haloN,
		peerXLoUpperHalo, peerXHiLowerHalo, myFlags, peerFlagsLo, peerFlagsHi, seq, NULL);
}

/* HDIA twin of spgpuDhellspmvHaloDot: z = A*xExt (+ halo exchange) and dRes[0] = sum_i xExt[haloN+i]*z[i].
 * With no neighbours (both peer pointers NULL, haloN = 0) it is the single-GPU fused SpMV + dot. */
extern "C" void spgpuDhdiaspmvHaloDot(spgpuHandle_t handle, double* z, const double* dM, const int* offsets,
	int hackSize, const int* hackOffsets, int rows, int cols, double* xExt, int haloN,
	double* peerXLoUpperHalo, double* peerXHiLowerHalo, unsigned* myFlags,
	unsigned* peerFlagsLo, unsigned* peerFlagsHi, unsigned seq, double* dRes)
{
	if (rows <= 0) {
		cudaMemsetAsync(dRes, 0, sizeof(double), handle->currentStream);
		return;
	}
	const unsigned rowBlocks = spgpu_ceil_div(rows, 128);
	double* partials = (double*)spgpuScratch(handle, (size_t)rowBlocks * sizeof(double));
	if (!partials)
		return;
	dhdia_spmv_halo_launch(handle, z, NULL, 1.0, dM, offsets, hackSize, hackOffsets, rows, cols, xExt, 0.0, haloN,
		peerXLoUpperHalo, peerXHiLowerHalo, myFlags, peerFlagsLo, peerFlagsHi, seq, partials);
	spgpuDsumDev(handle, (int)rowBlocks, partials, dRes);
}

/* ---- one-double sum all-reduce over NVLink peer memory ------------------------------- */

#define SPGPU_MAX_RANKS 16
struct PeerTables { unsigned char* t[SPGPU_MAX_RANKS]; };
struct alignas(16) ArSlot { double value; unsigned seq; unsigned pad; };

/*
 * Every rank stores (value, seq) into slot [parity][myRank] of EVERY rank's table (remote
 * 16-byte stores over NVLink, value first, then a release store of seq), then polls its own
 * table until all `world` slots of this parity carry seq and adds the values in rank order --
 * the same order on every rank, so all ranks get the same bits.  Payload is 8 bytes, so this
 * is pure latency: one NVLink round instead of an NCCL launch + ring.  Two parities because
 * a fast rank can be at most one all-reduce ahead of the slowest.
 */
__global__ void allreduce_sum_kernel(double* dValue, int world, int myRank, PeerTables tables,
	unsigned seqValue, unsigned* seqPtr, unsigned long long timeoutNs)
{
	/* seqPtr: device counter of completed all-reduces (this one is *seqPtr + 1, stored back at the end) */
	const unsigned seq = seqPtr ? *reinterpret_cast<volatile unsigned*>(seqPtr) + 1u : seqValue;
	__syncwarp();
	const int r = threadIdx.x;
	const unsigned parity = seq & 1u;
	double v = 0.0;
	if (r < world) {
		const double mine = *dValue;
		ArSlot* dst = reinterpret_cast<ArSlot*>(tables.t[r]) + parity * world + myRank;
		dst->value = mine;
		__threadfence_system();
		st_release_sys(&dst->seq, seq);
		const ArSlot* src = reinterpret_cast<const ArSlot*>(tables.t[myRank]) + parity * world + r;
		spin_until(&src->seq, seq, timeoutNs);
		v = *reinterpret_cast<const volatile double*>(&src->value);
	}
	double total = 0.0;
	for (int k = 0; k < world; ++k)
		total += __shfl_sync(SPGPU_FULL_MASK, v, k);
	if (r == 0) {
		*dValue = total;
		if (seqPtr)
			*seqPtr = seq;
	}
}

extern "C" void spgpuAllreduceSumDev(spgpuHandle_t handle, double* dValue, int world, int myRank,
	void* const* tables, unsigned seq)
{
	if (world <= 1)
		return;
	PeerTables pt;
	for (int r = 0; r < SPGPU_MAX_RANKS; ++r)
		pt.t[r] = r < world ? (unsigned char*)tables[r] : NULL;
	SpgpuHandlePriv* h = spgpuPriv(handle);
	unsigned* seqPtr = (seq == 0u && h->magic == SPGPU_PRIV_MAGIC) ? h->dArSeq : NULL;
	allreduce_sum_kernel<<<1, 32, 0, handle->currentStream>>>(dValue, world, myRank, pt, seq, seqPtr, 2000000000ull);
	spgpu_count_launch(handle);
}

/* ---- device-resident sequence numbers (CUDA-graph replay of a partitioned iteration) ---------------- */

extern "C" int spgpuSetSeqCounters(spgpuHandle_t handle, unsigned* dHaloSeq, unsigned* dAllreduceSeq)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC)
		return -1;
	h->dHaloSeq = dHaloSeq;
	h->dArSeq = dAllreduceSeq;
	return 0;
}

__global__ void seq_advance_kernel(unsigned* counter)
{
	*counter += 1u;
}

extern "C" void spgpuHaloSeqAdvance(spgpuHandle_t handle)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC || !h->dHaloSeq)
		return;
	seq_advance_kernel<<<1, 1, 0, handle->currentStream>>>(h->dHaloSeq);
	spgpu_count_launch(handle);
}

/* Bounded spin on a flag in LOCAL device memory written by a peer GPU. */
__global__ void wait_flag_kernel(const unsigned* flag, unsigned value, unsigned long long timeoutNs)
{
	unsigned long long t0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	for (;;) {
		unsigned v;
		asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
		if ((int)(v - value) >= 0)
			return;
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		if (t1 - t0 > timeoutNs) {
			printf("spgpuWaitFlag: timed out waiting for %u (saw %u)\n", value, v);
			return;
		}
		__nanosleep(200);
	}
}

extern "C" void spgpuWaitFlag(spgpuHandle_t handle, const unsigned* flag, unsigned value)
{
	wait_flag_kernel<<<1, 1, 0, handle->currentStream>>>(flag, value, 2000000000ull);
	spgpu_count_launch(handle);
}
