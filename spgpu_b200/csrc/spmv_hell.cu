/*
 * HELL SpMV for sm_100a:  z = alpha*A*x + beta*y, A in "hacked ELL".
 *
 * Replaces reference kernels/hell_spmv_base.cuh:103-157 (host entry) and
 * hell_spmv_base_template.cuh:19-357 (kernels).  Same contract: hackSize is a
 * multiple of 32, hackOffsets[h] is the ELEMENT offset of hack h (no
 * terminator), rS is mandatory and slots >= rS[i] are never read, rIdx (or
 * NULL) redirects the output row, beta == 0 means y is not read, z may alias y.
 *
 * Design: one warp per 32 consecutive rows (so a warp never straddles a hack),
 * the row walk is warp_rows_dot() of spmv_slots.cuh.  One launch per call, no
 * other host API traffic (the reference also issues cudaFuncSetCacheConfig on
 * every call).  The dead "large vector" loop of the reference is not kept
 * (SURVEY 2.2); indices are 64-bit where products can pass 2^31.
 */
#include "launch.cuh"
#include "spmv_hell_body.cuh"
#include "spmv_hell_bulk.cuh"

/*
 * HACK > 0: hackSize known at compile time (32 is what the reference recommends,
 * hell_conv.h:25) -- slot addresses become base + immediate.  MINB: CTAs of 128
 * threads the register allocator must leave room for per SM.
 */
template <typename T, int UNROLL, int HACK, int MINB>
__global__ void __launch_bounds__(128, MINB)
hell_spmv_kernel(const HellArgs<T> a)
{
	grid_dependency_wait();
	grid_launch_dependents();
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	hell_warp_rows<T, UNROLL, HACK>(a, i - (threadIdx.x & 31));
}

/*
 * Tail kernel of the split mode: persistent warps take (unit, chunk) items off the queue the
 * main kernel filled and walk slots [(c+1)T, (c+2)T) of the unit's 32 rows (row per lane,
 * coalesced), and leave one partial sum per lane in the item's slot of `partials`.  The warp
 * that finishes the LAST chunk of a unit (a counter in the unit's fold entry) adds the unit's
 * partials to z IN CHUNK ORDER (the main kernel has already stored beta*y + alpha*(first T
 * slots)), so the result does not depend on which warp took which item or when -- bit-
 * reproducible from run to run, like every other kernel here (no floating-point atomics).
 * The last warp to leave the kernel zeroes the header for the next call (no memset launch).
 */
template <typename T, int UNROLL, int HACK>
__global__ void __launch_bounds__(128, 8)
hell_tail_kernel(const HellArgs<T> a)
{
	const int hackSize = HACK > 0 ? HACK : a.hackSize;
	const unsigned lane = threadIdx.x & 31;
	const unsigned queued = min(__ldg(a.workHeader), (unsigned)a.workCap);
	for (;;) {
		unsigned idx = 0;
		if (lane == 0)
			idx = atomicAdd(a.workHeader + 1, 1u);
		idx = __shfl_sync(SPGPU_FULL_MASK, idx, 0);
		if (idx >= queued)
			break;
		const uint4 item = a.workItems[idx];
		if (item.x == SPGPU_WORK_INVALID)
			continue;
		const unsigned warpRow = item.x << 5;
		const unsigned i = warpRow + lane;
		const bool live = i < (unsigned)a.rows;
		const int len = live ? __ldg(a.rS + i) : 0;
		const int kBeg = (int)(item.y + 1u) * a.splitT;
		const int kEnd = min(len, kBeg + a.splitT);
		const unsigned hack = warpRow / (unsigned)hackSize;
		const long long at = (long long)__ldg(a.hackOffsets + hack) + (warpRow % (unsigned)hackSize) + lane;
		const T* vp = a.cM + at;
		const int* ip = a.rP + at;
		T acc = Num<T>::zero();
		const int top = __reduce_max_sync(SPGPU_FULL_MASK, kEnd);
		for (int k0 = kBeg; k0 < top; k0 += UNROLL) {
			int col[UNROLL];
			T v[UNROLL];
			T xv[UNROLL];
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const bool on = (k0 + u) < kEnd;
				col[u] = a.baseIndex;
				v[u] = Num<T>::zero();
				if (on) {
					col[u] = ld_stream64(ip + (long long)(k0 + u) * hackSize);
					v[u] = ld_stream64(vp + (long long)(k0 + u) * hackSize);
				}
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u) {
				const bool on = (k0 + u) < kEnd;
				xv[u] = Num<T>::zero();
				if (on)
					xv[u] = ld_keep(a.x + (col[u] - a.baseIndex));
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
				acc = Num<T>::fma(v[u], xv[u], acc);
		}
		a.partials[(size_t)idx * 32 + lane] = acc;               /* zero for lanes with nothing in this chunk */
		__threadfence();                                         /* partial visible before the counter moves */
		unsigned finished = 0;
		if (lane == 0)
			finished = atomicAdd(&a.foldList[item.z].w, 1u);
		finished = __shfl_sync(SPGPU_FULL_MASK, finished, 0);
		const uint4 unit = a.foldList[item.z];                   /* x, y, z written by the main kernel */
		if (finished + 1u == unit.z) {
			/* this warp finished the unit's last chunk: fold in chunk order (loads bypass L1: the
			 * other partials were written by other SMs, and the buffer is reused from call to call) */
			__threadfence();
			T sum = Num<T>::zero();
			for (unsigned c = 0; c < unit.z; ++c)
				sum = Num<T>::add(sum, __ldcg(a.partials + (size_t)(unit.y + c) * 32 + lane));
			if (live && Num<T>::nonzero(sum)) {                  /* rows without a deep part keep their bits */
				const unsigned out = a.rIdx ? (unsigned)__ldg(a.rIdx + i) : i;
				a.z[out] = Num<T>::fma(a.alpha, sum, a.z[out]);
			}
		}
	}
	/* every warp of the grid ends here exactly once; the last one leaves the header zeroed */
	if (lane == 0 && atomicAdd(a.workHeader + 3, 1u) == ((gridDim.x * blockDim.x) >> 5) - 1u) {
		a.workHeader[0] = 0u;
		a.workHeader[1] = 0u;
		a.workHeader[2] = 0u;
		a.workHeader[3] = 0u;
	}
}

/* Bulk-async variant: returns false when the call is not eligible (then the
 * direct kernel runs).  Stage capacity is sized from avgNnzPerRow. */
template <typename T, int UNROLL, int HACK>
static bool hell_spmv_try_bulk(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* cM, const int* rP, const int* hackOffsets, const int* rS,
	const int* rIdx, int avgNnzPerRow, int rows, const T* x, T beta,
	int baseIndex, int longCut)
{
	const int avg = avgNnzPerRow > 0 ? avgNnzPerRow : 1;
	const int slots = avg + (avg / 4 > 1 ? avg / 4 : 1);
	const int cap = HB_CONSUMER_WARPS * 32 * slots;
	const size_t stageBytes = (size_t)cap * (sizeof(T) + sizeof(int)) + HB_CONSUMER_WARPS * 32 * sizeof(int);
	const size_t budget = 110 * 1024;                     /* two CTAs per SM */
	int stages = (int)((budget - 256) / stageBytes);
	if (stages > 4) stages = 4;
	if (stages < 2)
		return false;
	if ((((size_t)cM | (size_t)rP | (size_t)rS) & 15) != 0)
		return false;                                     /* bulk copies need 16-byte aligned sources */
	const size_t smem = stages * stageBytes + 2 * stages * sizeof(uint64_t);
	/* function attributes are per device: set on every call (a process may drive several GPUs) */
	if (cudaFuncSetAttribute(hell_spmv_bulk_kernel<T, HACK, UNROLL>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(112 * 1024)) != cudaSuccess)
		return false;
	const int tiles = (rows + HB_CONSUMER_WARPS * 32 - 1) / (HB_CONSUMER_WARPS * 32);
	int grid = 2 * handle->multiProcessorCount;
	if (grid > tiles) grid = tiles;
	hell_spmv_bulk_kernel<T, HACK, UNROLL><<<grid, HB_THREADS, smem, handle->currentStream>>>(
		z, y, alpha, cM, rP, hackOffsets, rS, rIdx, rows, x, beta, baseIndex, longCut, cap, stages);
	spgpu_count_launch(handle);
	return true;
}

template <typename T, int UNROLL>
static void hell_spmv_launch(spgpuHandle_t handle, T* z, const T* y, T alpha,
	const T* cM, const int* rP, int hackSize, const int* hackOffsets,
	const int* rS, const int* rIdx, int avgNnzPerRow, int rows, const T* x,
	T beta, int baseIndex)
{
	if (rows <= 0)
		return;
	const SpgpuTuning* t = spgpu_tuning(handle);
	const int block = 128;
	const unsigned grid = spgpu_ceil_div(rows, block);
	const int longCut = spgpu_long_cut(t, avgNnzPerRow);
	/* hellVariant: 0 auto, 1 direct loads predicated on rS, 2 direct loads with
	 * unpredicated slab reads, 3 bulk-async pipeline */
	int variant = t->hellVariant;
	if (variant == 0)
		variant = 2;
	if (variant == 3 && rows >= 8 * HB_CONSUMER_WARPS * 32) {
		bool done = false;
		if (hackSize == 32)
			done = hell_spmv_try_bulk<T, UNROLL, 32>(handle, z, y, alpha, cM, rP, hackOffsets, rS, rIdx,
				avgNnzPerRow, rows, x, beta, baseIndex, longCut);
		else if (hackSize == 64)
			done = hell_spmv_try_bulk<T, UNROLL, 64>(handle, z, y, alpha, cM, rP, hackOffsets, rS, rIdx,
				avgNnzPerRow, rows, x, beta, baseIndex, longCut);
		if (done)
			return;
	}
	const int speculate = variant != 1;
	cudaStream_t s = handle->currentStream;
	HellArgs<T> args = { z, y, alpha, cM, rP, hackSize, hackOffsets, rS, rIdx, rows, x, beta, baseIndex, longCut, speculate,
		0, NULL, NULL, 0, NULL, NULL, spgpu_hell_prefetch(handle, t) };
	/* Split mode: on for length-sorted matrices (rIdx given: their long rows sit together in a
	 * few hacks whose warps would be the critical path), or forced with hellSplit = 1; off with
	 * hellSplit = -1.  Costs one small extra launch per call (no memset: the header resets itself). */
	const bool split = t->hellSplit > 0 || (t->hellSplit == 0 && rIdx != NULL);
	if (split) {
		const int cap = t->hellSplit > 1 ? t->hellSplit : (1 << 16);     /* hellSplit > 1: queue capacity (tests) */
		/* scratch: items | fold list | partials (32 lanes per item); the 4-word header lives in the
		 * handle's counter block (spgpu_internal.h; zero at creation and again after every call) */
		const size_t foldAt = (size_t)cap * sizeof(uint4);
		const size_t partAt = foldAt + (size_t)cap * sizeof(uint4);
		unsigned char* scratch = (unsigned char*)spgpuScratch(handle, partAt + (size_t)cap * 32 * sizeof(T));
		SpgpuHandlePriv* hp = spgpuPriv(handle);
		if (scratch && hp->magic == SPGPU_PRIV_MAGIC) {
			args.splitT = longCut > 64 ? longCut : 64;
			args.workHeader = hp->dTicket + SPGPU_TICKET_SPLIT;
			args.workItems = reinterpret_cast<uint4*>(scratch);
			args.foldList = reinterpret_cast<uint4*>(scratch + foldAt);
			args.partials = reinterpret_cast<T*>(scratch + partAt);
			args.workCap = cap;
		}
	}
#define HELL_ARGS args
	/* Resident warps per SM are set by the register budget (__launch_bounds__ minimum
	 * CTAs): more warps = more loads in flight, until the allocator starts spilling.
	 * Measured on B200 (512^3 Laplacian, double): 32 warps 2.33 ms, 40 warps (48 regs, no
	 * spill) 2.15 ms, 48 warps (40 regs, 20 B spill) 2.11-2.21 ms.  Defaults: float 48
	 * warps (39 regs, no spill), double 40, complex 32 (they spill below 64 registers and
	 * are already at the HBM peak).  hellBlock forces a level: <=64 -> 32 warps, 192 -> 40,
	 * >=256 -> 48. */
	int level = Num<T>::is_complex ? 8 : (sizeof(T) == 4 ? 12 : 10);
	if (t->hellBlock >= 256) level = 12;
	else if (t->hellBlock == 192) level = 10;
	else if (t->hellBlock > 0 && t->hellBlock <= 64) level = 8;
	if (hackSize == 32) {
		if (level == 10)      spgpu_launch_dep(handle, hell_spmv_kernel<T, UNROLL, 32, 10>, grid, block, HELL_ARGS);
		else if (level == 12) spgpu_launch_dep(handle, hell_spmv_kernel<T, UNROLL, 32, 12>, grid, block, HELL_ARGS);
		else                  spgpu_launch_dep(handle, hell_spmv_kernel<T, UNROLL, 32, 8>, grid, block, HELL_ARGS);
	} else if (hackSize == 64) {
		if (level >= 10) spgpu_launch_dep(handle, hell_spmv_kernel<T, UNROLL, 64, 10>, grid, block, HELL_ARGS);
		else             spgpu_launch_dep(handle, hell_spmv_kernel<T, UNROLL, 64, 8>, grid, block, HELL_ARGS);
	} else {
		spgpu_launch_dep(handle, hell_spmv_kernel<T, UNROLL, 0, 8>, grid, block, HELL_ARGS);
	}
#undef HELL_ARGS
	spgpu_count_launch(handle);
	if (args.splitT > 0) {
		const unsigned tg = (unsigned)handle->multiProcessorCount * 4u;
		if (hackSize == 32)      hell_tail_kernel<T, UNROLL, 32><<<tg, 128, 0, s>>>(args);
		else if (hackSize == 64) hell_tail_kernel<T, UNROLL, 64><<<tg, 128, 0, s>>>(args);
		else                     hell_tail_kernel<T, UNROLL, 0><<<tg, 128, 0, s>>>(args);
		spgpu_count_launch(handle);
	}
}

#define SPGPU_DEFINE_HELLSPMV(S, T, U)                                        \
	extern "C" void spgpu##S##hellspmv(spgpuHandle_t handle, T* z, const T* y, \
		T alpha, const T* cM, const int* rP, int hackSize,                     \
		const int* hackOffsets, const int* rS, const int* rIdx,                \
		int avgNnzPerRow, int rows, const T* x, T beta, int baseIndex)         \
	{                                                                          \
		hell_spmv_launch<T, U>(handle, z, y, alpha, cM, rP, hackSize,          \
			hackOffsets, rS, rIdx, avgNnzPerRow, rows, x, beta, baseIndex);    \
	}

SPGPU_DEFINE_HELLSPMV(S, float, 8)
SPGPU_DEFINE_HELLSPMV(D, double, 8)
SPGPU_DEFINE_HELLSPMV(C, cuFloatComplex, 8)
SPGPU_DEFINE_HELLSPMV(Z, cuDoubleComplex, 4)
