/*
 * Handle / stream core of the B200-native spGPU drop-in.
 *
 * Behaviour follows reference src/core/core.c:9-97: spgpuCreate caches a few
 * device properties and owns one stream; the handle is returned even when the
 * property query fails (status SPGPU_UNSPECIFIED); spgpuSetStream(h, 0) falls
 * back to the handle's own stream.  New here: the handle also owns the scratch
 * its reductions need (device partials, a ticket counter and one pinned result
 * slot), so reductions on different handles never share state.
 */
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>

#include "spgpu_internal.h"
#include "spgpu_ext.h"

static void default_tuning(SpgpuTuning* t)
{
	t->hellVariant = 0;
	t->hellBlock = 0;          /* 0 = per-type default occupancy, <=64 force 32 warps, >=256 force 48 */
	t->hellLongFactor = 2;     /* a row deeper than factor x avgNnzPerRow (>= 32) slots counts as a spike */
	t->hellSplit = 0;
	t->hdiaVariant = 0;
	t->hdiaBlock = 0;
	t->ellRows = 0;
	t->redBlocksPerSm = 8;
	t->vecBlocksPerSm = 8;
	t->spinTimeoutMs = 20000;
	t->haloTrace = 0;
	t->l2Fetch = 0;
	t->redInflight = 0;
	t->ellShortMinB = 0;
	t->hellPrefetch = 0;
	t->hdiaPrefetch = 0;
	t->pdl = 1;               /* programmatic dependent launch on (launch.cuh); 0 = plain stream order */
}

spgpuStatus_t spgpuCreate(spgpuHandle_t* pHandle, int device)
{
	struct cudaDeviceProp prop;
	cudaError_t propErr = cudaGetDeviceProperties(&prop, device);
	SpgpuHandlePriv* h = (SpgpuHandlePriv*)calloc(1, sizeof(SpgpuHandlePriv));
	int previous = 0;
	cudaError_t err = cudaSuccess;

	if (!h) {
		*pHandle = NULL;
		return SPGPU_OUTOFMEMORY;
	}
	h->magic = SPGPU_PRIV_MAGIC;
	default_tuning(&h->tune);
	{
		const char* dbg = getenv("SPGPU_DEBUG");
		const char* pdl = getenv("SPGPU_PDL");          /* overrides the default of the `pdl` tuning key (A/B runs) */
		h->debug = dbg && dbg[0] && dbg[0] != '0';
		if (pdl && pdl[0])
			h->tune.pdl = pdl[0] != '0';
	}

	cudaGetDevice(&previous);
	cudaSetDevice(device);
	cudaStreamCreate(&h->pub.defaultStream);
	h->pub.currentStream = h->pub.defaultStream;

	/* reduction scratch: partial slots + ticket on the device, result in
	 * mapped pinned memory so the finishing block stores it straight to the host */
	if (propErr == cudaSuccess) {
		err = cudaMalloc(&h->dPartials, (size_t)SPGPU_RED_MAX_BLOCKS * SPGPU_RED_SLOT_BYTES);
		if (err == cudaSuccess)
			err = cudaMalloc((void**)&h->dTicket, SPGPU_TICKET_WORDS * sizeof(unsigned));
		if (err == cudaSuccess)
			err = cudaMemset(h->dTicket, 0, SPGPU_TICKET_WORDS * sizeof(unsigned));
		if (err == cudaSuccess)
			err = cudaHostAlloc(&h->hResult, 64, cudaHostAllocMapped);
		if (err == cudaSuccess) {
			memset(h->hResult, 0, 64);
			err = cudaHostGetDevicePointer(&h->dResult, h->hResult, 0);
		}
		if (err == cudaSuccess) {
			/* sticky device status word: byte 32 of the same mapped pinned block */
			h->hStatus = (unsigned*)((char*)h->hResult + 32);
			h->dStatus = (unsigned*)((char*)h->dResult + 32);
			err = cudaEventCreateWithFlags(&h->switchEvent, cudaEventDisableTiming);
		}
	}
	cudaSetDevice(previous);

	h->pub.device = device;
	if (propErr == cudaSuccess) {
		h->pub.warpSize = prop.warpSize;
		h->pub.maxThreadsPerBlock = prop.maxThreadsPerBlock;
		h->pub.multiProcessorCount = prop.multiProcessorCount;
		h->pub.maxGridSizeX = prop.maxGridSize[0];
		h->pub.maxGridSizeY = prop.maxGridSize[1];
		h->pub.maxGridSizeZ = prop.maxGridSize[2];
		h->pub.capabilityMajor = prop.major;
		h->pub.capabilityMinor = prop.minor;
		h->l2Bytes = prop.l2CacheSize;
		h->smemPerBlockOptin = (int)prop.sharedMemPerBlockOptin;
	}

	*pHandle = &h->pub;

	if (propErr != cudaSuccess)
		return SPGPU_UNSPECIFIED;
	if (err == cudaErrorMemoryAllocation)
		return SPGPU_OUTOFMEMORY;
	if (err != cudaSuccess)
		return SPGPU_UNSPECIFIED;
	/* this build carries sm_100a code only */
	if (prop.major < 10)
		return SPGPU_UNSUPPORTED;
	return SPGPU_SUCCESS;
}

void spgpuDestroy(spgpuHandle_t pHandle)
{
	SpgpuHandlePriv* h = spgpuPriv(pHandle);
	if (!h)
		return;
	cudaStreamDestroy(h->pub.defaultStream);
	if (h->magic == SPGPU_PRIV_MAGIC) {
		if (h->dPartials) cudaFree(h->dPartials);
		if (h->dTicket) cudaFree(h->dTicket);
		if (h->hResult) cudaFreeHost(h->hResult);
		if (h->dBig) cudaFree(h->dBig);
		if (h->dTrace) cudaFree(h->dTrace);
		if (h->dTicketsMany) cudaFree(h->dTicketsMany);
		if (h->switchEvent) cudaEventDestroy(h->switchEvent);
		h->magic = 0;
	}
	free(h);
}

void spgpuStreamCreate(spgpuHandle_t pHandle, cudaStream_t* stream)
{
	int previous = 0;
	cudaGetDevice(&previous);
	cudaSetDevice(pHandle->device);
	cudaStreamCreate(stream);
	cudaSetDevice(previous);
}

void spgpuStreamDestroy(cudaStream_t stream)
{
	cudaStreamDestroy(stream);
}

/*
 * Reference core.c:62-72 only swaps the pointer.  Here the handle owns scratch that its kernels
 * reuse in stream order (reduction partials, tickets, the split-mode queue), so work queued on the
 * new stream must not overtake what is still in flight on the old one: an event recorded on the
 * old stream is waited for by the new one (a device-side dependency, no host synchronisation).
 * Skipped while either stream is being captured into a CUDA graph (the caller orders those).
 */
void spgpuSetStream(spgpuHandle_t pHandle, cudaStream_t stream)
{
	SpgpuHandlePriv* h = spgpuPriv(pHandle);
	cudaStream_t next = stream ? stream : h->pub.defaultStream;
	cudaStream_t prev = h->pub.currentStream;
	if (next != prev && h->magic == SPGPU_PRIV_MAGIC && h->switchEvent) {
		enum cudaStreamCaptureStatus a = cudaStreamCaptureStatusNone, b = cudaStreamCaptureStatusNone;
		int previous = 0;
		cudaGetDevice(&previous);
		if (previous != h->pub.device)
			cudaSetDevice(h->pub.device);
		if (cudaStreamIsCapturing(prev, &a) == cudaSuccess && cudaStreamIsCapturing(next, &b) == cudaSuccess &&
				a == cudaStreamCaptureStatusNone && b == cudaStreamCaptureStatusNone) {
			if (cudaEventRecord(h->switchEvent, prev) == cudaSuccess)
				cudaStreamWaitEvent(next, h->switchEvent, 0);
		}
		(void)cudaGetLastError();        /* a stream the caller has already destroyed is not this call's error */
		if (previous != h->pub.device)
			cudaSetDevice(previous);
	}
	h->pub.currentStream = next;
}

cudaStream_t spgpuGetStream(spgpuHandle_t pHandle)
{
	return pHandle->currentStream;
}

size_t spgpuSizeOf(spgpuType_t typeCode)
{
	static const size_t bytes[5] = {
		sizeof(int), sizeof(float), sizeof(double),
		sizeof(cuFloatComplex), sizeof(cuDoubleComplex)
	};
	if (typeCode < 0 || typeCode > SPGPU_TYPE_COMPLEX_DOUBLE)
		return 0;
	return bytes[typeCode];
}

void* spgpuScratch(spgpuHandle_t handle, size_t bytes)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC)
		return NULL;
	if (h->bigBytes >= bytes)
		return h->dBig;
	/* grow: earlier work on the stream may still use the old block.  This synchronises and allocates, which a
	 * CUDA-graph capture does not allow: callers that capture size the scratch first (spgpuReserveScratch, or one
	 * eager call of the same shape).  The block lives on the HANDLE's device whatever device is current. */
	{
		int previous = 0;
		void* fresh = NULL;
		cudaGetDevice(&previous);
		if (previous != h->pub.device)
			cudaSetDevice(h->pub.device);
		cudaStreamSynchronize(h->pub.currentStream);
		if (h->dBig)
			cudaFree(h->dBig);
		h->dBig = NULL;
		h->bigBytes = 0;
		if (cudaMalloc(&fresh, bytes) == cudaSuccess) {
			h->dBig = fresh;
			h->bigBytes = bytes;
		}
		if (previous != h->pub.device)
			cudaSetDevice(previous);
	}
	return h->dBig;
}

unsigned* spgpuTickets(spgpuHandle_t handle, size_t count)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC)
		return NULL;
	if (h->ticketsMany < count) {
		int previous = 0;
		void* fresh = NULL;
		size_t want = count < 1024 ? 1024 : count;
		cudaGetDevice(&previous);
		if (previous != h->pub.device)
			cudaSetDevice(h->pub.device);
		cudaStreamSynchronize(h->pub.currentStream);
		if (h->dTicketsMany)
			cudaFree(h->dTicketsMany);
		h->dTicketsMany = NULL;
		h->ticketsMany = 0;
		if (cudaMalloc(&fresh, want * sizeof(unsigned)) == cudaSuccess &&
				cudaMemset(fresh, 0, want * sizeof(unsigned)) == cudaSuccess) {
			h->dTicketsMany = (unsigned*)fresh;
			h->ticketsMany = want;
		} else if (fresh) {
			cudaFree(fresh);
		}
		if (previous != h->pub.device)
			cudaSetDevice(previous);
	}
	return h->dTicketsMany;
}

int spgpuReserveScratch(spgpuHandle_t handle, size_t bytes)
{
	return spgpuScratch(handle, bytes) ? 0 : -1;
}

int spgpuGetDeviceStatus(spgpuHandle_t handle, int clear)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	unsigned v;
	if (!h || h->magic != SPGPU_PRIV_MAGIC || !h->hStatus)
		return -1;
	v = *(volatile unsigned*)h->hStatus;
	if (clear)
		*(volatile unsigned*)h->hStatus = 0u;
	return (int)v;
}

/* ---- additive API (include/spgpu_ext.h) ---------------------------------- */

#define TUNE_KEYS(X)          \
	X(hellVariant)            \
	X(hellBlock)              \
	X(hellLongFactor)         \
	X(hellSplit)              \
	X(hdiaVariant)            \
	X(hdiaBlock)              \
	X(ellRows)                \
	X(redBlocksPerSm)         \
	X(vecBlocksPerSm)         \
	X(spinTimeoutMs)          \
	X(haloTrace)              \
	X(l2Fetch)                \
	X(redInflight)            \
	X(ellShortMinB)           \
	X(hellPrefetch)           \
	X(hdiaPrefetch)           \
	X(pdl)

int spgpuSetTuning(spgpuHandle_t handle, const char* key, int value)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC || !key)
		return -1;
	/* keys with a side effect on the device */
	if (strcmp(key, "haloTrace") == 0 || strcmp(key, "l2Fetch") == 0) {
		int previous = 0, rc = 0;
		cudaGetDevice(&previous);
		if (previous != h->pub.device)
			cudaSetDevice(h->pub.device);
		if (key[0] == 'h') {
			if (value && !h->dTrace) {
				void* p = NULL;
				const size_t bytes = (size_t)1024 * 8 * sizeof(unsigned long long);
				if (cudaMalloc(&p, bytes) == cudaSuccess && cudaMemset(p, 0, bytes) == cudaSuccess)
					h->dTrace = (unsigned long long*)p;
				else
					rc = -1;
			} else if (!value && h->dTrace) {
				cudaStreamSynchronize(h->pub.currentStream);
				cudaFree(h->dTrace);
				h->dTrace = NULL;
			}
			if (rc == 0)
				h->tune.haloTrace = value;
		} else {
			if (value > 0 && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value) != cudaSuccess) {
				(void)cudaGetLastError();
				rc = -1;
			} else {
				h->tune.l2Fetch = value;
			}
		}
		if (previous != h->pub.device)
			cudaSetDevice(previous);
		return rc;
	}
#define SET_KEY(name) if (strcmp(key, #name) == 0) { h->tune.name = value; return 0; }
	TUNE_KEYS(SET_KEY)
#undef SET_KEY
	return -1;
}

int spgpuGetTuning(spgpuHandle_t handle, const char* key)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC || !key)
		return -1;
#define GET_KEY(name) if (strcmp(key, #name) == 0) return h->tune.name;
	TUNE_KEYS(GET_KEY)
#undef GET_KEY
	return -1;
}

unsigned long long spgpuGetLaunchCount(spgpuHandle_t handle)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	return (h && h->magic == SPGPU_PRIV_MAGIC) ? h->launches : 0ull;
}

const char* spgpuB200Version(void)
{
	return "spgpu-b200 0.2 (sm_100a)";
}
