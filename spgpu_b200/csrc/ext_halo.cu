/*
 * Row-partitioned multi-GPU SpMV, device side (include/spgpu_ext.h).  No reference counterpart:
 * the reference is one handle per device with no communication (reference core.h:88-93).
 *
 *   - spgpuHaloPush / spgpuWaitFlag / spgpuHaloExchange / spgpuHaloAck: the halo exchange of a
 *     1-D chain of ranks as separate kernels (NVLink peer stores + flag words);
 *   - spgpu?{hell,hdia}spmvHalo[Dot]: ONE kernel per partitioned SpMV -- the exchange travels
 *     inside the SpMV launch, for all four value types;
 *   - spgpu?allreduceSumDev: one-value sum all-reduce over NVLink peer memory.
 */
#include <cstring>
#include "launch.cuh"
#include "peer_sync.cuh"
#include "spmv_hell_body.cuh"
#include "spmv_hdia_body.cuh"

/* byte copy into (peer) memory by the CTAs [cta, ctas) of a group: 128-bit stores when both sides allow */
__device__ __forceinline__ void copy_bytes(void* dst, const void* src, size_t bytes, unsigned cta, unsigned ctas)
{
	const size_t tid = (size_t)cta * blockDim.x + threadIdx.x;
	const size_t nthreads = (size_t)ctas * blockDim.x;
	if ((((size_t)dst | (size_t)src) & 15) == 0) {
		const size_t nv = bytes >> 4;
		uint4* d4 = reinterpret_cast<uint4*>(dst);
		const uint4* s4 = reinterpret_cast<const uint4*>(src);
		size_t p = tid;
		for (; p + nthreads < nv; p += 2 * nthreads) {          /* two loads in flight per thread */
			const uint4 a = s4[p], b = s4[p + nthreads];
			d4[p] = a;
			d4[p + nthreads] = b;
		}
		if (p < nv)
			d4[p] = s4[p];
		for (size_t e = (nv << 4) + tid; e < bytes; e += nthreads)
			reinterpret_cast<unsigned char*>(dst)[e] = reinterpret_cast<const unsigned char*>(src)[e];
	} else if ((((size_t)dst | (size_t)src | bytes) & 3) == 0) {
		unsigned* d = reinterpret_cast<unsigned*>(dst);
		const unsigned* s = reinterpret_cast<const unsigned*>(src);
		for (size_t e = tid; e < (bytes >> 2); e += nthreads)
			d[e] = s[e];
	} else {
		for (size_t e = tid; e < bytes; e += nthreads)
			reinterpret_cast<unsigned char*>(dst)[e] = reinterpret_cast<const unsigned char*>(src)[e];
	}
}

static SpinCtl spin_ctl(spgpuHandle_t handle)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	SpinCtl c;
	const int ms = h->magic == SPGPU_PRIV_MAGIC ? h->tune.spinTimeoutMs : 20000;
	c.timeoutNs = ms > 0 ? (unsigned long long)ms * 1000000ull : 0ull;
	c.status = h->magic == SPGPU_PRIV_MAGIC ? h->dStatus : NULL;
	return c;
}

/* ---- halo exchange as separate kernels ------------------------------------------------- */

/*
 * Copies `bytes` bytes of this GPU into a peer GPU's memory; the last CTA to finish (ticket in
 * local memory) makes the data visible system-wide and release-stores flagValue into the peer's
 * flag word.
 */
__global__ void __launch_bounds__(256)
halo_push_kernel(void* peerDst, const void* src, size_t bytes, unsigned* peerFlag, unsigned flagValue, unsigned* ticket)
{
	__shared__ bool amLast;
	if (peerDst)
		copy_bytes(peerDst, src, bytes, blockIdx.x, gridDim.x);
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0)
		amLast = (atomicAdd(ticket, 1u) == gridDim.x - 1);
	__syncthreads();
	if (amLast && threadIdx.x == 0) {
		*ticket = 0u;
		if (peerFlag) {
			__threadfence_system();
			st_release_sys(peerFlag, flagValue);
		}
	}
}

static unsigned push_grid(spgpuHandle_t handle, size_t bytes)
{
	long long want = (long long)((bytes / 16 + 255) / 256);
	if (want > 2 * handle->multiProcessorCount) want = 2 * handle->multiProcessorCount;
	if (want < 1) want = 1;
	return (unsigned)want;
}

extern "C" void spgpuHaloPush(spgpuHandle_t handle, void* peerDst, const void* src, size_t bytes,
	unsigned* peerFlag, unsigned flagValue)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	/* the ticket word next to the reductions' one (offset 16 bytes) */
	halo_push_kernel<<<push_grid(handle, bytes), 256, 0, handle->currentStream>>>(peerDst, src, bytes,
		peerFlag, flagValue, h->dTicket + 4);
	spgpu_count_launch(handle);
}

extern "C" void spgpuDhaloPush(spgpuHandle_t handle, double* peerDst, const double* src,
	int n, unsigned* peerFlag, unsigned flagValue)
{
	spgpuHaloPush(handle, peerDst, src, n > 0 ? (size_t)n * sizeof(double) : 0, peerFlag, flagValue);
}

/*
 * One kernel: even CTAs copy srcLo into the LOWER neighbour's upper halo zone, odd CTAs srcHi into
 * the UPPER neighbour's lower halo zone.  Before copying, a CTA waits until that neighbour has
 * acknowledged the previous halo (so it is not overwritten while still being read).  The last CTA
 * to finish release-stores the sequence number into both neighbours' "ready" flags and then waits
 * for this rank's own two "ready" flags, so the kernel completes exactly when this rank's halos
 * have arrived.
 */
__global__ void __launch_bounds__(256)
halo_exchange_kernel(void* dstLo, const void* srcLo, void* dstHi, const void* srcHi,
	size_t bytes, const unsigned* ackLo, const unsigned* ackHi, unsigned* peerReadyLo,
	unsigned* peerReadyHi, const unsigned* myReadyLo, const unsigned* myReadyHi,
	unsigned seq, unsigned* ticket, SpinCtl spin)
{
	__shared__ bool amLast;
	const bool toHi = (blockIdx.x & 1) != 0;
	void* dst = toHi ? dstHi : dstLo;
	const void* src = toHi ? srcHi : srcLo;
	const unsigned* ack = toHi ? ackHi : ackLo;
	if (dst) {
		if (threadIdx.x == 0 && ack && seq > 1)
			spin_until(ack, seq - 1, spin);
		__syncthreads();
		copy_bytes(dst, src, bytes, blockIdx.x >> 1, gridDim.x >> 1);
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
		if (myReadyLo) spin_until(myReadyLo, seq, spin);
		if (myReadyHi) spin_until(myReadyHi, seq, spin);
	}
}

extern "C" void spgpuHaloExchange(spgpuHandle_t handle, void* peerDstLo, const void* srcLo,
	void* peerDstHi, const void* srcHi, size_t bytes, const unsigned* ackLo, const unsigned* ackHi,
	unsigned* peerReadyLo, unsigned* peerReadyHi, const unsigned* myReadyLo,
	const unsigned* myReadyHi, unsigned seq)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	long long want = 2 * (long long)((bytes / 32 + 255) / 256);
	const long long cap = 2LL * (handle->multiProcessorCount / 2 > 0 ? handle->multiProcessorCount / 2 : 1);
	if (want > cap) want = cap;
	if (want < 2) want = 2;
	halo_exchange_kernel<<<(unsigned)want, 256, 0, handle->currentStream>>>(peerDstLo, srcLo, peerDstHi, srcHi,
		bytes, ackLo, ackHi, peerReadyLo, peerReadyHi, myReadyLo, myReadyHi, seq, h->dTicket + 8, spin_ctl(handle));
	spgpu_count_launch(handle);
}

extern "C" void spgpuDhaloExchange(spgpuHandle_t handle, double* peerDstLo, const double* srcLo,
	double* peerDstHi, const double* srcHi, int n, const unsigned* ackLo, const unsigned* ackHi,
	unsigned* peerReadyLo, unsigned* peerReadyHi, const unsigned* myReadyLo,
	const unsigned* myReadyHi, unsigned seq)
{
	spgpuHaloExchange(handle, peerDstLo, srcLo, peerDstHi, srcHi, n > 0 ? (size_t)n * sizeof(double) : 0,
		ackLo, ackHi, peerReadyLo, peerReadyHi, myReadyLo, myReadyHi, seq);
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

/* Bounded spin on a flag in LOCAL device memory written by a peer GPU. */
__global__ void wait_flag_kernel(const unsigned* flag, unsigned value, SpinCtl spin)
{
	spin_until(flag, value, spin);
}

extern "C" void spgpuWaitFlag(spgpuHandle_t handle, const unsigned* flag, unsigned value)
{
	wait_flag_kernel<<<1, 1, 0, handle->currentStream>>>(flag, value, spin_ctl(handle));
	spgpu_count_launch(handle);
}

/* ---- HELL / HDIA SpMV fused with the halo exchange: ONE kernel per partitioned SpMV ---- */

/*
 * Protocol (per side; words of the per-rank flag block, spgpu_ext.h).  Exchange number seq = 1, 2, ...:
 *   push CTAs     (the first CTAs of the grid) tell both neighbours "my exchange seq - 1 is over" -- this kernel runs,
 *                 so every earlier kernel of this rank's stream, including the rows that read my zones, has finished:
 *                 the ACK costs the row blocks nothing -- then wait for the neighbour's ack of seq - 1 (it has finished
 *                 reading the zone this rank is about to overwrite), copy this rank's boundary entries into the
 *                 neighbour's zone over NVLink and release-store seq into the neighbour's READY word;
 *   boundary rows (the row blocks that read a zone) acquire-spin on the local ready word before their first load and
 *                 then multiply like every other block.
 * A rank's push of exchange seq needs the neighbour to have STARTED its kernel seq, and its boundary rows -- scheduled
 * late in the grid -- need the neighbour's push, i.e. the neighbour's start plus the copy: neighbours may drift most of
 * a kernel apart without anybody waiting.  Every block -- interior or boundary -- runs the SAME inlined row walk with
 * plain weak loads of x (ld.global: inside the memory model, ordered behind the CTA's acquire; the single-GPU kernels
 * use ld.global.nc).  Measured alternatives (bench/halo_dot_probe.py, the 64-plane slab of one of eight ranks, plain
 * kernel 0.275 ms): zones double-buffered by the parity of seq instead of acknowledged, the boundary rows choosing the
 * zone per gather in an out-of-line copy of the row walk: 0.280 ms, and 0.322 ms with the dot fused in (the call
 * spoils the register allocation of the common path); acks by a ticket per boundary warp: 0.298 ms (and round 1's
 * barrier + fence + ticket per boundary block: + 7 us at N = 8).  Consecutive fused calls of a rank must be ordered
 * on ONE stream.
 */
template <typename T>
struct HaloArgs {
	T* dstLo; const T* srcLo;               /* my first n owned entries -> lower neighbour's upper zone */
	T* dstHi; const T* srcHi;               /* my last n owned entries  -> upper neighbour's lower zone */
	int n;
	const unsigned* ackLo; const unsigned* ackHi;        /* local : the neighbour has consumed my previous entries     */
	unsigned* peerReadyLo; unsigned* peerReadyHi;        /* remote: my entries for `seq` are in place                  */
	const unsigned* myReadyLo; const unsigned* myReadyHi;/* local : the neighbour's entries for `seq` are in place     */
	unsigned* peerAckLo; unsigned* peerAckHi;            /* remote: I have consumed the neighbour's entries for `seq`  */
	unsigned seq;                           /* sequence number of this exchange, or ... */
	const unsigned* seqPtr;                 /* ... (seq == 0) device counter of COMPLETED exchanges: this one is *seqPtr + 1 */
	unsigned* pushTicket;                   /* handle-owned counter, zero between launches */
	int pushCtas;
	unsigned headBlocks;                    /* 128-row blocks [0, headBlocks) hold a row that reads the lower zone */
	unsigned firstHiBlock;                  /* 128-row blocks [firstHiBlock, ..) hold a row that reads the upper zone */
	/* the block order, worked out on the host (every instruction in front of a warp's first load adds to its
	 * lifetime, and this kernel lives on the edge of being latency-bound): CTAs [0, early) multiply interior
	 * blocks nLo.., then come the nLo lower-boundary blocks, the nHi upper-boundary blocks (from hiStart), and a
	 * few waves of interior blocks behind them */
	unsigned early, nLo, nHi, hiStart;
	SpinCtl spin;
	unsigned long long* trace;              /* NULL, or 8 words per exchange (haloTrace tuning key) */
};

#define SPGPU_TRACE_SLOTS 1024

/*
 * z_i * x[xOffset+i] summed over the WARP, one partial per 32 rows (fixed order).  Per warp, not per CTA: this
 * kernel sits on the edge of being latency-bound (40 resident warps just cover the HBM latency), so whatever a warp
 * does besides its loads shows up in the run time.  Measured on the 512^3 slab of one of two ranks (plain kernel
 * 1.094 ms, fused halo 1.083 ms): CTA-wide sum behind a barrier 1.276 ms; the block's last warp to arrive adds the
 * products the others left in shared memory (needs an opening barrier) 1.276 ms; a butterfly in every warp 1.233 ms.
 */
template <typename T>
__device__ __forceinline__ void warp_dot_partial(Acc2 contrib, typename DotPartial<T>::type* partials, unsigned slot)
{
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		contrib.a += __shfl_xor_sync(SPGPU_FULL_MASK, contrib.a, m);
		if (Num<T>::is_complex)
			contrib.b += __shfl_xor_sync(SPGPU_FULL_MASK, contrib.b, m);
	}
	if ((threadIdx.x & 31) == 0)
		acc2_to_partial(contrib, partials[slot]);
}

/* what a row block of the fused kernel multiplies: the HELL or the HDIA warp body */
template <typename T, int UNROLL, int HACK>
struct HellRowBody {
	HellArgs<T> a;
	__device__ __forceinline__ int rows() const { return a.rows; }
	__device__ __forceinline__ const T* x() const { return a.x; }
	template <class XG>
	__device__ __forceinline__ T run(unsigned warpRow, const XG xg) const
	{
		T zval;
		hell_warp_rows_value_x<T, UNROLL, HACK, XG>(a, warpRow, zval, xg);
		return zval;
	}
};

template <typename T, int UNROLL, int HACK>
struct HdiaRowBody {
	HdiaArgs<T> a;
	__device__ __forceinline__ int rows() const { return a.rows; }
	__device__ __forceinline__ const T* x() const { return a.x; }
	template <class XG>
	__device__ __forceinline__ T run(unsigned warpRow, const XG xg) const
	{
		T zval;
		hdia_warp_rows_value_x<T, UNROLL, HACK, false, XG>(a, warpRow, zval, xg);
		return zval;
	}
};

/*
 * grid = pushCtas + ceil(rows/128).  The first pushCtas CTAs move the two boundary runs into the
 * neighbours' zones over NVLink (after the neighbours' acks) and publish `seq`.  Every other CTA multiplies 128 rows; the CTAs
 * are numbered so that most interior row blocks come first and the blocks that read a zone come
 * late (followed only by a few waves of interior blocks, so that a late neighbour delays nothing
 * but the blocks that need it) -- by the time the hardware schedules them the neighbours' entries
 * have normally long arrived.  A block may read both zones (a rank with fewer rows than two halo
 * widths): it then waits for both words.
 */
/* which 128-row block the CTA multiplies (the order is worked out on the host: halo_args) */
__host__ __device__ __forceinline__ unsigned halo_block_of(unsigned b, unsigned early, unsigned nLo, unsigned nHi, unsigned hiStart)
{
	if (b < early) return nLo + b;                                      /* most of the interior first */
	if (b < early + nLo) return b - early;                              /* then the lower boundary    */
	if (b < early + nLo + nHi) return hiStart + (b - early - nLo);      /* the upper boundary         */
	return b - nHi;                                                     /* the rest of the interior   */
}

template <typename T>
__device__ __forceinline__ unsigned halo_row_block(const HaloArgs<T>& hx)
{
	return halo_block_of(blockIdx.x - hx.pushCtas, hx.early, hx.nLo, hx.nHi, hx.hiStart);
}

/* which row blocks wait for a zone and in which order the blocks are multiplied: host arithmetic only, shared by
 * halo_args below and by spgpuHaloBlockPlan (spgpu_ext.h), which the CPU tests sweep over (rows, haloN) */
struct HaloBlockPlan { unsigned headBlocks, firstHiBlock, early, nLo, nHi, hiStart; };

static HaloBlockPlan halo_block_plan(int rows, int haloN, bool lo, bool hi, int multiProcessorCount)
{
	HaloBlockPlan p;
	const int n = haloN < rows ? haloN : rows;
	/* rows [0, haloN) may read the lower zone, rows [rows - haloN, rows) the upper one (band |col - row| <= haloN) */
	p.headBlocks = lo ? (unsigned)((n + 127) / 128) : 0u;
	p.firstHiBlock = hi ? (unsigned)((rows - n) / 128) : 0xffffffffu;
	const unsigned rowBlocks = (unsigned)((rows + 127) / 128);
	const unsigned fillerCap = 3u * 10u * (unsigned)multiProcessorCount;    /* about three waves */
	p.nLo = p.headBlocks < rowBlocks ? p.headBlocks : rowBlocks;
	p.hiStart = p.firstHiBlock < rowBlocks ? p.firstHiBlock : rowBlocks;
	if (p.hiStart < p.nLo) p.hiStart = p.nLo;
	p.nHi = rowBlocks - p.hiStart;
	const unsigned interior = p.hiStart - p.nLo;
	const unsigned filler = (interior >> 2) < fillerCap ? (interior >> 2) : fillerCap;
	p.early = interior - filler;
	return p;
}

extern "C" int spgpuHaloBlockPlan(int rows, int haloN, int hasLower, int hasUpper, int multiProcessorCount, unsigned* plan)
{
	if (rows < 0 || haloN < 0 || multiProcessorCount < 1 || !plan)
		return -1;
	const HaloBlockPlan p = halo_block_plan(rows, haloN, hasLower != 0 && haloN > 0, hasUpper != 0 && haloN > 0, multiProcessorCount);
	plan[0] = p.headBlocks; plan[1] = p.firstHiBlock; plan[2] = p.early; plan[3] = p.nLo; plan[4] = p.nHi; plan[5] = p.hiStart;
	return 0;
}

extern "C" unsigned spgpuHaloBlockOf(const unsigned* plan, unsigned cta)
{
	return halo_block_of(cta, plan[2], plan[3], plan[4], plan[5]);
}

/*
 * HALO = false: no neighbours at all (single-GPU fused SpMV + dot) -- the exchange code is compiled out.
 */
template <typename T, class Body, int MINB, bool DOT, bool HALO>
__global__ void __launch_bounds__(128, MINB)
spmv_halo_kernel(const Body body, const HaloArgs<T> hx, int xOffset, typename DotPartial<T>::type* __restrict__ warpPartials /* one per 32 rows */)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const unsigned rows = (unsigned)body.rows();
	unsigned rb = blockIdx.x;
	bool needLo = false, needHi = false;
	unsigned seq = 0;
	unsigned long long* tr = NULL;
	T zval;
	if (HALO) {
		/* the counter is advanced by a later kernel in stream order (spgpuHaloSeqAdvance), never during this
		 * one, so every CTA reads the same value whenever it is scheduled */
		seq = hx.seqPtr ? *reinterpret_cast<const volatile unsigned*>(hx.seqPtr) + 1u : hx.seq;
		tr = hx.trace ? hx.trace + (size_t)(seq & (SPGPU_TRACE_SLOTS - 1)) * 8 : NULL;
		if (blockIdx.x < (unsigned)hx.pushCtas) {
			const bool toHi = (blockIdx.x & 1) != 0;
			T* dst = toHi ? hx.dstHi : hx.dstLo;
			const T* src = toHi ? hx.srcHi : hx.srcLo;
			const unsigned* ack = toHi ? hx.ackHi : hx.ackLo;
			if (tr && threadIdx.x == 0 && blockIdx.x == 0) {
				for (int k = 2; k < 8; ++k)
					tr[k] = 0ull;
				tr[0] = global_timer_ns();
			}
			if (threadIdx.x == 0 && seq > 1u) {
				/* This kernel runs, so every earlier kernel of this rank's stream has finished -- including the rows of
				 * exchange seq - 1 that read my zones: that IS the acknowledgement, and it costs the row blocks nothing. */
				if (blockIdx.x == 0) {
					if (hx.peerAckLo) st_release_sys(hx.peerAckLo, seq - 1u);
					if (hx.peerAckHi) st_release_sys(hx.peerAckHi, seq - 1u);
				}
				if (dst && ack)
					spin_until(ack, seq - 1u, hx.spin);      /* the neighbour has finished reading what I am about to overwrite */
			}
			__syncthreads();
			if (dst)
				copy_bytes(dst, src, (size_t)hx.n * sizeof(T), blockIdx.x >> 1, (unsigned)hx.pushCtas >> 1);
			__threadfence_system();
			__syncthreads();
			if (threadIdx.x == 0) {
				if (atomicAdd(hx.pushTicket, 1u) == (unsigned)hx.pushCtas - 1u) {
					*hx.pushTicket = 0u;
					__threadfence_system();
					if (hx.peerReadyLo) st_release_sys(hx.peerReadyLo, seq);
					if (hx.peerReadyHi) st_release_sys(hx.peerReadyHi, seq);
					if (tr)
						tr[1] = global_timer_ns();
				}
			}
			return;
		}
		rb = halo_row_block(hx);
		needLo = rb < hx.headBlocks && hx.myReadyLo != NULL;
		needHi = rb >= hx.firstHiBlock && hx.myReadyHi != NULL;
		if (needLo || needHi) {
			if (threadIdx.x == 0) {
				unsigned long long t0 = 0;
				if (tr) {
					t0 = global_timer_ns();
					atomicCAS(tr + 6, 0ull, t0);
				}
				if (needLo) spin_until(hx.myReadyLo, seq, hx.spin);
				if (tr) {
					const unsigned long long t1 = global_timer_ns();
					if (needLo && t1 - t0 > 2000ull) { atomicAdd(tr + 2, t1 - t0); atomicAdd(tr + 4, 1ull); }
					t0 = t1;
				}
				if (needHi) spin_until(hx.myReadyHi, seq, hx.spin);
				if (tr) {
					const unsigned long long t1 = global_timer_ns();
					if (needHi && t1 - t0 > 2000ull) { atomicAdd(tr + 3, t1 - t0); atomicAdd(tr + 5, 1ull); }
					atomicMax(tr + 7, t1);          /* the last boundary block to get going */
				}
			}
			__syncthreads();
		}
		/* one row walk for every block: weak loads of x */
		const XWeak<T> xg = { body.x() };
		zval = body.run(rb * 128u + (threadIdx.x & ~31u), xg);
	} else {
		const XPlain<T> xg = { body.x() };
		zval = body.run(rb * 128u + (threadIdx.x & ~31u), xg);
	}
	if (DOT) {
		/* nothing of the prologue is kept in registers across the row walk: the block is worked out again */
		if (HALO)
			rb = halo_row_block(hx);
		const unsigned myRow = rb * 128u + threadIdx.x;
		Acc2 c = { 0.0, 0.0 };
		if (myRow < rows)
			c = to_acc2<T>(Num<T>::mul(zval, __ldg(body.x() + xOffset + myRow)));
		warp_dot_partial<T>(c, warpPartials, rb * 4u + (threadIdx.x >> 5));
	}
}

/* flag words (spgpu_ext.h): [4] ready-from-below [5] ready-from-above [6] ack-from-below [7] ack-from-above */
template <typename T>
static HaloArgs<T> halo_args(spgpuHandle_t handle, T* xExt, int rows, int haloN, const spgpuHaloLinks* L, unsigned seq)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	HaloArgs<T> hx;
	memset(&hx, 0, sizeof(hx));
	const bool lo = L && L->peerFlagsLo && haloN > 0, hi = L && L->peerFlagsHi && haloN > 0;
	/* a rank with a neighbour sends its first / last haloN OWNED entries: it must own at least that many
	 * (the host layers check this: mg.py _check_block, spgpuMg*Create) */
	hx.n = haloN;
	if (lo) {
		hx.dstLo = (T*)L->peerLoUpperZone;
		hx.ackLo = L->myFlags + 6;
		hx.peerReadyLo = L->peerFlagsLo + 5;
		hx.myReadyLo = L->myFlags + 4;
		hx.peerAckLo = L->peerFlagsLo + 7;
	}
	if (hi) {
		hx.dstHi = (T*)L->peerHiLowerZone;
		hx.ackHi = L->myFlags + 7;
		hx.peerReadyHi = L->peerFlagsHi + 4;
		hx.myReadyHi = L->myFlags + 5;
		hx.peerAckHi = L->peerFlagsHi + 6;
	}
	hx.srcLo = xExt + haloN;
	hx.srcHi = xExt + rows;                                 /* last haloN owned entries */
	hx.seq = seq;
	hx.seqPtr = (seq == 0u && h->magic == SPGPU_PRIV_MAGIC) ? h->dHaloSeq : NULL;
	hx.pushTicket = h->dTicket + 8;
	hx.pushCtas = (lo || hi) ? 8 : 0;
	{
		const HaloBlockPlan p = halo_block_plan(rows, haloN, lo, hi, handle->multiProcessorCount);
		hx.headBlocks = p.headBlocks;
		hx.firstHiBlock = p.firstHiBlock;
		hx.early = p.early;
		hx.nLo = p.nLo;
		hx.nHi = p.nHi;
		hx.hiStart = p.hiStart;
	}
	hx.spin = spin_ctl(handle);
	hx.trace = h->magic == SPGPU_PRIV_MAGIC ? h->dTrace : NULL;
	return hx;
}

/* resident CTAs per SM the register allocator must leave room for (spmv_hell.cu: float 48 warps, double 40, complex 32) */
template <typename T> struct HaloMinB { static constexpr int hell = Num<T>::is_complex ? 8 : (sizeof(T) == 4 ? 12 : 10); };

#define SPGPU_DECL_PLAIN_SPMV(S, T, R)                                                                \
	extern "C" void spgpu##S##hellspmv(spgpuHandle_t, T*, const T*, T, const T*, const int*, int, const int*, \
		const int*, const int*, int, int, const T*, T, int);                                               \
	extern "C" void spgpu##S##hdiaspmv(spgpuHandle_t, T*, const T*, T, const T*, const int*, int, const int*, \
		int, int, const T*, T);
SPGPU_FOR_FLOAT_TYPES(SPGPU_DECL_PLAIN_SPMV)
static inline void plain_hellspmv(spgpuHandle_t h, float* z, const float* y, float a, const float* cM, const int* rP, int hs, const int* ho, const int* rS, int avg, int rows, const float* x, float b, int base) { spgpuShellspmv(h, z, y, a, cM, rP, hs, ho, rS, NULL, avg, rows, x, b, base); }
static inline void plain_hellspmv(spgpuHandle_t h, double* z, const double* y, double a, const double* cM, const int* rP, int hs, const int* ho, const int* rS, int avg, int rows, const double* x, double b, int base) { spgpuDhellspmv(h, z, y, a, cM, rP, hs, ho, rS, NULL, avg, rows, x, b, base); }
static inline void plain_hellspmv(spgpuHandle_t h, cuFloatComplex* z, const cuFloatComplex* y, cuFloatComplex a, const cuFloatComplex* cM, const int* rP, int hs, const int* ho, const int* rS, int avg, int rows, const cuFloatComplex* x, cuFloatComplex b, int base) { spgpuChellspmv(h, z, y, a, cM, rP, hs, ho, rS, NULL, avg, rows, x, b, base); }
static inline void plain_hellspmv(spgpuHandle_t h, cuDoubleComplex* z, const cuDoubleComplex* y, cuDoubleComplex a, const cuDoubleComplex* cM, const int* rP, int hs, const int* ho, const int* rS, int avg, int rows, const cuDoubleComplex* x, cuDoubleComplex b, int base) { spgpuZhellspmv(h, z, y, a, cM, rP, hs, ho, rS, NULL, avg, rows, x, b, base); }
static inline void plain_hdiaspmv(spgpuHandle_t h, float* z, const float* y, float a, const float* dM, const int* off, int hs, const int* ho, int rows, int cols, const float* x, float b) { spgpuShdiaspmv(h, z, y, a, dM, off, hs, ho, rows, cols, x, b); }
static inline void plain_hdiaspmv(spgpuHandle_t h, double* z, const double* y, double a, const double* dM, const int* off, int hs, const int* ho, int rows, int cols, const double* x, double b) { spgpuDhdiaspmv(h, z, y, a, dM, off, hs, ho, rows, cols, x, b); }
static inline void plain_hdiaspmv(spgpuHandle_t h, cuFloatComplex* z, const cuFloatComplex* y, cuFloatComplex a, const cuFloatComplex* dM, const int* off, int hs, const int* ho, int rows, int cols, const cuFloatComplex* x, cuFloatComplex b) { spgpuChdiaspmv(h, z, y, a, dM, off, hs, ho, rows, cols, x, b); }
static inline void plain_hdiaspmv(spgpuHandle_t h, cuDoubleComplex* z, const cuDoubleComplex* y, cuDoubleComplex a, const cuDoubleComplex* dM, const int* off, int hs, const int* ho, int rows, int cols, const cuDoubleComplex* x, cuDoubleComplex b) { spgpuZhdiaspmv(h, z, y, a, dM, off, hs, ho, rows, cols, x, b); }

template <typename T, int UNROLL>
static void hell_spmv_halo_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* cM, const int* rP, int hackSize, const int* hackOffsets, const int* rS,
	int avgNnzPerRow, int rows, T* xExt, T beta, int baseIndex, int haloN,
	const spgpuHaloLinks* links, unsigned seq, typename DotPartial<T>::type* ctaPartials)
{
	const SpgpuTuning* t = spgpu_tuning(handle);
	const HellArgs<T> a = { z, y, alpha, cM, rP, hackSize, hackOffsets, rS, NULL, rows, xExt, beta,
		baseIndex, spgpu_long_cut(t, avgNnzPerRow), t->hellVariant != 1, 0, NULL, NULL, 0, NULL, NULL,
		spgpu_hell_prefetch(handle, t) };
	const HaloArgs<T> hx = halo_args<T>(handle, xExt, rows, haloN, links, seq);
	const bool halo = hx.pushCtas > 0;
	if (!halo && !ctaPartials) {                 /* no neighbours, no dot: the plain entry point */
		plain_hellspmv(handle, z, y, alpha, cM, rP, hackSize, hackOffsets, rS, avgNnzPerRow, rows, xExt, beta, baseIndex);
		return;
	}
	const unsigned grid = hx.pushCtas + spgpu_ceil_div(rows, 128);
	constexpr int MB = HaloMinB<T>::hell;
	const HellRowBody<T, UNROLL, 32> b32 = { a };
	const HellRowBody<T, UNROLL, 0> b0 = { a };
	if (ctaPartials) {
		if (hackSize == 32) {
			/* Occupancy of the dot variants, from measurement on the 512^3 slabs of one of two / one of eight ranks
			 * (bench/halo_dot_probe.py; plain kernel 1.079 / 0.275 ms): without neighbours 48 resident warps per SM
			 * for the real types (1.082 ms against 1.094 at the plain kernel's 40); with the halo code 40 warps
			 * (1.152 / 0.300 ms against 1.231 / 0.305 at 48; 36 and 44 warps: 1.20 / 0.314 and 1.24 / 0.323).  The row walk lives on the edge of being
			 * latency-bound, and how ptxas schedules its loads changes with everything around it.  hellBlock = 192 /
			 * 256 force the plain kernel's occupancy / 48 warps for an A/B. */
			constexpr int MD = Num<T>::is_complex ? 8 : 12;
			const bool dense = t->hellBlock >= 256 || (t->hellBlock != 192 && !halo);
			if (dense) {
				if (halo) spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 32>, MD, true, true>, grid, 128, b32, hx, haloN, ctaPartials);
				else      spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 32>, MD, true, false>, grid, 128, b32, hx, haloN, ctaPartials);
			} else if (halo) spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 32>, MB, true, true>, grid, 128, b32, hx, haloN, ctaPartials);
			else      spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 32>, MB, true, false>, grid, 128, b32, hx, haloN, ctaPartials);
		} else {
			if (halo) spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 0>, 8, true, true>, grid, 128, b0, hx, haloN, ctaPartials);
			else      spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 0>, 8, true, false>, grid, 128, b0, hx, haloN, ctaPartials);
		}
	} else {
		if (hackSize == 32)
			spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 32>, MB, false, true>, grid, 128, b32, hx, haloN, ctaPartials);
		else
			spgpu_launch_dep(handle, spmv_halo_kernel<T, HellRowBody<T, UNROLL, 0>, 8, false, true>, grid, 128, b0, hx, haloN, ctaPartials);
	}
	spgpu_count_launch(handle);
}

template <typename T, int UNROLL>
static void hdia_spmv_halo_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* dM, const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols,
	T* xExt, T beta, int haloN, const spgpuHaloLinks* links, unsigned seq, typename DotPartial<T>::type* ctaPartials)
{
	const HdiaArgs<T> a = { z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, xExt, beta };
	const HaloArgs<T> hx = halo_args<T>(handle, xExt, rows, haloN, links, seq);
	const bool halo = hx.pushCtas > 0;
	if (!halo && !ctaPartials) {
		plain_hdiaspmv(handle, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows, cols, xExt, beta);
		return;
	}
	const unsigned grid = hx.pushCtas + spgpu_ceil_div(rows, 128);
	const HdiaRowBody<T, UNROLL, 32> b32 = { a };
	const HdiaRowBody<T, UNROLL, 0> b0 = { a };
	if (ctaPartials) {
		if (hackSize == 32) {
			if (halo) spgpu_launch_dep(handle, spmv_halo_kernel<T, HdiaRowBody<T, UNROLL, 32>, 8, true, true>, grid, 128, b32, hx, haloN, ctaPartials);
			else      spgpu_launch_dep(handle, spmv_halo_kernel<T, HdiaRowBody<T, UNROLL, 32>, 8, true, false>, grid, 128, b32, hx, haloN, ctaPartials);
		} else {
			if (halo) spgpu_launch_dep(handle, spmv_halo_kernel<T, HdiaRowBody<T, UNROLL, 0>, 8, true, true>, grid, 128, b0, hx, haloN, ctaPartials);
			else      spgpu_launch_dep(handle, spmv_halo_kernel<T, HdiaRowBody<T, UNROLL, 0>, 8, true, false>, grid, 128, b0, hx, haloN, ctaPartials);
		}
	} else {
		if (hackSize == 32)
			spgpu_launch_dep(handle, spmv_halo_kernel<T, HdiaRowBody<T, UNROLL, 32>, 8, false, true>, grid, 128, b32, hx, haloN, ctaPartials);
		else
			spgpu_launch_dep(handle, spmv_halo_kernel<T, HdiaRowBody<T, UNROLL, 0>, 8, false, true>, grid, 128, b0, hx, haloN, ctaPartials);
	}
	spgpu_count_launch(handle);
}

/* folds per-row-block partials (ext_krylov.cu), optionally all-reducing the total across the ranks in its last CTA */
template <typename T>
void spgpu_fold_partials(spgpuHandle_t handle, const typename DotPartial<T>::type* partials, long long n, T* dRes,
	const spgpuPeerAllreduce* ar);

ArArgs spgpu_ar_args(spgpuHandle_t handle, const spgpuPeerAllreduce* ar)
{
	ArArgs a;
	memset(&a, 0, sizeof(a));
	if (!ar || ar->world <= 1)
		return a;
	SpgpuHandlePriv* h = spgpuPriv(handle);
	a.world = ar->world < SPGPU_MAX_RANKS ? ar->world : SPGPU_MAX_RANKS;
	a.myRank = ar->myRank;
	for (int r = 0; r < a.world; ++r)
		a.tables.t[r] = (unsigned char*)ar->tables[r];
	a.seq = ar->seq;
	a.seqPtr = (ar->seq == 0u && h->magic == SPGPU_PRIV_MAGIC) ? h->dArSeq : NULL;
	a.spin = spin_ctl(handle);
	return a;
}

#define SPGPU_DEFINE_HALO(S, T, UH, UD)                                                              \
	extern "C" void spgpu##S##hellspmvHalo(spgpuHandle_t handle, T* z, const T* y, T alpha,             \
		const T* cM, const int* rP, int hackSize, const int* hackOffsets, const int* rS,                \
		int avgNnzPerRow, int rows, T* xExt, T beta, int baseIndex, int haloN,                          \
		const spgpuHaloLinks* links, unsigned seq)                                                      \
	{                                                                                                   \
		if (rows <= 0) return;                                                                          \
		hell_spmv_halo_launch<T, UH>(handle, z, y, alpha, cM, rP, hackSize, hackOffsets, rS,            \
			avgNnzPerRow, rows, xExt, beta, baseIndex, haloN, links, seq, NULL);                        \
	}                                                                                                   \
	/* z = A*xExt with the halo exchange inside, plus dRes[0] = sum_i xExt[haloN+i]*z[i] (this rank's   \
	 * share of p.Ap, all-reduced when ar is given): one partial per 32 rows in handle scratch */       \
	extern "C" void spgpu##S##hellspmvHaloDot(spgpuHandle_t handle, T* z, const T* cM, const int* rP,   \
		int hackSize, const int* hackOffsets, const int* rS, int avgNnzPerRow, int rows, T* xExt,       \
		int baseIndex, int haloN, const spgpuHaloLinks* links, unsigned seq, T* dRes,                   \
		const spgpuPeerAllreduce* ar)                                                                   \
	{                                                                                                   \
		typedef DotPartial<T>::type P;                                                                  \
		const unsigned rowBlocks = spgpu_ceil_div(rows > 0 ? rows : 0, 128);                            \
		P* partials = rowBlocks ? (P*)spgpuScratch(handle, (size_t)rowBlocks * 4 * sizeof(P)) : NULL;   \
		if (rowBlocks && !partials) return;                                                             \
		if (rowBlocks)                                                                                  \
			hell_spmv_halo_launch<T, UH>(handle, z, NULL, Num<T>::from_real(1), cM, rP, hackSize,       \
				hackOffsets, rS, avgNnzPerRow, rows, xExt, Num<T>::zero(), baseIndex, haloN, links,     \
				seq, partials);                                                                         \
		spgpu_fold_partials<T>(handle, partials, (long long)rowBlocks * 4, dRes, ar);                   \
	}                                                                                                   \
	extern "C" void spgpu##S##hellspmvDot(spgpuHandle_t handle, T* z, const T* cM, const int* rP,       \
		int hackSize, const int* hackOffsets, const int* rS, int rows, const T* x, int baseIndex,       \
		int xOffset, T* dRes)                                                                           \
	{                                                                                                   \
		spgpu##S##hellspmvHaloDot(handle, z, cM, rP, hackSize, hackOffsets, rS, 8, rows,                \
			const_cast<T*>(x), baseIndex, xOffset, NULL, 1u, dRes, NULL);                               \
	}                                                                                                   \
	extern "C" void spgpu##S##hdiaspmvHalo(spgpuHandle_t handle, T* z, const T* y, T alpha,             \
		const T* dM, const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols,      \
		T* xExt, T beta, int haloN, const spgpuHaloLinks* links, unsigned seq)                          \
	{                                                                                                   \
		if (rows <= 0) return;                                                                          \
		hdia_spmv_halo_launch<T, UD>(handle, z, y, alpha, dM, offsets, hackSize, hackOffsets, rows,     \
			cols, xExt, beta, haloN, links, seq, NULL);                                                 \
	}                                                                                                   \
	extern "C" void spgpu##S##hdiaspmvHaloDot(spgpuHandle_t handle, T* z, const T* dM,                  \
		const int* offsets, int hackSize, const int* hackOffsets, int rows, int cols, T* xExt,          \
		int haloN, const spgpuHaloLinks* links, unsigned seq, T* dRes, const spgpuPeerAllreduce* ar)    \
	{                                                                                                   \
		typedef DotPartial<T>::type P;                                                                  \
		const unsigned rowBlocks = spgpu_ceil_div(rows > 0 ? rows : 0, 128);                            \
		P* partials = rowBlocks ? (P*)spgpuScratch(handle, (size_t)rowBlocks * 4 * sizeof(P)) : NULL;   \
		if (rowBlocks && !partials) return;                                                             \
		if (rowBlocks)                                                                                  \
			hdia_spmv_halo_launch<T, UD>(handle, z, NULL, Num<T>::from_real(1), dM, offsets, hackSize,  \
				hackOffsets, rows, cols, xExt, Num<T>::zero(), haloN, links, seq, partials);            \
		spgpu_fold_partials<T>(handle, partials, (long long)rowBlocks * 4, dRes, ar);                   \
	}

SPGPU_DEFINE_HALO(S, float, 8, 9)
SPGPU_DEFINE_HALO(D, double, 8, 9)
SPGPU_DEFINE_HALO(C, cuFloatComplex, 8, 9)
SPGPU_DEFINE_HALO(Z, cuDoubleComplex, 4, 4)

/* ---- one-value sum all-reduce over NVLink peer memory as a kernel of its own --------- */

template <typename T>
__global__ void allreduce_sum_kernel(T* dValue, ArArgs ar)
{
	const Acc2 total = peer_allreduce_sum_warp(to_acc2<T>(*dValue), ar);
	if (threadIdx.x == 0)
		*dValue = from_acc2<T>(total);
}

#define SPGPU_DEFINE_ALLREDUCE(S, T, R)                                                          \
	extern "C" void spgpu##S##allreduceSumDev(spgpuHandle_t handle, T* dValue,                      \
		const spgpuPeerAllreduce* ar)                                                               \
	{                                                                                               \
		if (!ar || ar->world <= 1) return;                                                          \
		allreduce_sum_kernel<T><<<1, 32, 0, handle->currentStream>>>(dValue, spgpu_ar_args(handle, ar)); \
		spgpu_count_launch(handle);                                                                 \
	}
SPGPU_FOR_FLOAT_TYPES(SPGPU_DEFINE_ALLREDUCE)

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

/* ---- per-exchange trace of the fused kernel (haloTrace tuning key) ---------------------------------- */

extern "C" int spgpuHaloTraceRead(spgpuHandle_t handle, unsigned long long* hostOut, int firstSeq, int count)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC || !h->dTrace || count < 0 || count > SPGPU_TRACE_SLOTS)
		return -1;
	cudaStreamSynchronize(handle->currentStream);
	for (int k = 0; k < count; ++k) {
		const size_t slot = (size_t)((unsigned)(firstSeq + k) & (SPGPU_TRACE_SLOTS - 1));
		if (cudaMemcpy(hostOut + (size_t)k * 8, h->dTrace + slot * 8, 8 * sizeof(unsigned long long),
				cudaMemcpyDeviceToHost) != cudaSuccess)
			return -1;
	}
	return 0;
}

/* ---- forcing the kernels of this file into the device ahead of time ------------------------------------ */

/*
 * CUDA loads a kernel when it is first launched (lazy loading), and that load waits for the kernels already
 * running on the device.  A process that drives SEVERAL ranks must therefore never meet a first launch while
 * one of its own kernels is spinning on a peer: the load would wait for the spinning kernel, which waits for
 * a kernel the same host thread has not launched yet.  spgpuMgCreate calls this once per device.
 */
template <typename K>
static int preload_one(K kernel)
{
	cudaFuncAttributes a;
	return cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(kernel)) == cudaSuccess ? 0 : 1;
}

template <typename T, class Body, int MINB>
static int preload_body()
{
	/* also the 48-warp variants the hellBlock knob selects */
	return preload_one(spmv_halo_kernel<T, Body, MINB, false, true>) + preload_one(spmv_halo_kernel<T, Body, MINB, true, true>)
		+ preload_one(spmv_halo_kernel<T, Body, MINB, true, false>);
}

template <typename T, int UH, int UD>
static int preload_type()
{
	return preload_body<T, HellRowBody<T, UH, 32>, HaloMinB<T>::hell>() + preload_body<T, HellRowBody<T, UH, 32>, 12>()
		+ preload_body<T, HellRowBody<T, UH, 0>, 8>()
		+ preload_body<T, HdiaRowBody<T, UD, 32>, 8>() + preload_body<T, HdiaRowBody<T, UD, 0>, 8>()
		+ preload_one(allreduce_sum_kernel<T>);
}

extern "C" int spgpuPreloadHaloKernels(void)
{
	int bad = preload_type<float, 8, 9>() + preload_type<double, 8, 9>() + preload_type<cuFloatComplex, 8, 9>()
		+ preload_type<cuDoubleComplex, 4, 4>();
	bad += preload_one(halo_push_kernel) + preload_one(halo_exchange_kernel) + preload_one(halo_ack_kernel)
		+ preload_one(wait_flag_kernel) + preload_one(seq_advance_kernel);
	return bad ? -1 : 0;
}
