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
	t->hellLongFactor = 4;     /* a row deeper than factor x avgNnzPerRow (>= 32) slots counts as a spike */
	t->hellSplit = 0;
	t->hdiaVariant = 0;
	t->hdiaBlock = 0;
	t->ellRows = 0;
	t->redBlocksPerSm = 4;
	t->vecBlocksPerSm = 8;
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
		h->debug = dbg && dbg[0] && dbg[0] != '0';
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

void spgpuSetStream(spgpuHandle_t pHandle, cudaStream_t stream)
{
	SpgpuHandleStruct* h = (SpgpuHandleStruct*)pHandle;
	h->currentStream = stream ? stream : h->defaultStream;
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
	/* grow: earlier work on the stream may still use the old block */
	cudaStreamSynchronize(h->pub.currentStream);
	if (h->dBig)
		cudaFree(h->dBig);
	h->dBig = NULL;
	h->bigBytes = 0;
	if (cudaMalloc(&h->dBig, bytes) != cudaSuccess)
		return NULL;
	h->bigBytes = bytes;
	return h->dBig;
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
	X(vecBlocksPerSm)

int spgpuSetTuning(spgpuHandle_t handle, const char* key, int value)
{
	SpgpuHandlePriv* h = spgpuPriv(handle);
	if (!h || h->magic != SPGPU_PRIV_MAGIC || !key)
		return -1;
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
	return "spgpu-b200 0.1 (sm_100a)";
}
