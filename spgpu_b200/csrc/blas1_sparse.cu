/*
 * Sparse-vector companions for sm_100a: gather, scatter (I, S, D, C, Z) and the
 * ELL coefficient update.
 *
 * Replaces reference kernels/gath_base.cuh:32-85, scat_base.cuh:32-88 and
 * ell_csput_base.cuh:33-75.  gath: xValues[i] = y[idx[i]-base]; scat:
 * y[p] = beta != 0 ? beta*y[p] + xValues[i] : xValues[i]; entries whose
 * position p = idx[i]-base is negative are skipped by both.  Duplicate
 * positions in a scatter race exactly as they do in the reference.
 * Index-bound kernels: the index and value streams are read coalesced with
 * evict-first loads, the indirect side goes through L2.
 */
#include "launch.cuh"
#include "numeric.cuh"

template <typename T>
__global__ void __launch_bounds__(256)
gath_kernel(T* __restrict__ values, long long count, const int* __restrict__ indices,
	int firstIndex, const T* __restrict__ vector)
{
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += nthreads) {
		const long long p = (long long)ld_stream(indices + i) - firstIndex;
		if (p >= 0)
			values[i] = ld_keep(vector + p);
	}
}

template <typename T>
__global__ void __launch_bounds__(256)
scat_kernel(T* vector, long long count, const int* __restrict__ indices,
	const T* __restrict__ values, int firstIndex, T beta)
{
	const bool useBeta = Num<T>::nonzero(beta);
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += nthreads) {
		const long long p = (long long)ld_stream(indices + i) - firstIndex;
		if (p < 0)
			continue;
		const T v = ld_stream(values + i);
		vector[p] = useBeta ? Num<T>::fma(beta, vector[p], v) : v;
	}
}

static unsigned sparse_grid(spgpuHandle_t handle, long long n)
{
	long long want = (n + 255) / 256;
	const long long cap = (long long)handle->multiProcessorCount * 16;
	if (cap > 0 && want > cap) want = cap;
	return (unsigned)(want < 1 ? 1 : want);
}

#define SPGPU_DEFINE_SPVEC(S, T)                                               \
	extern "C" void spgpu##S##gath(spgpuHandle_t h, T* xValues, int xNnz,       \
		const int* xIndices, int xBaseIndex, const T* y)                        \
	{                                                                           \
		if (xNnz <= 0) return;                                                  \
		gath_kernel<T><<<sparse_grid(h, xNnz), 256, 0, h->currentStream>>>(     \
			xValues, xNnz, xIndices, xBaseIndex, y);                            \
		spgpu_count_launch(h);                                                  \
	}                                                                           \
	extern "C" void spgpu##S##scat(spgpuHandle_t h, T* y, int xNnz,             \
		const T* xValues, const int* xIndices, int xBaseIndex, T beta)          \
	{                                                                           \
		if (xNnz <= 0) return;                                                  \
		scat_kernel<T><<<sparse_grid(h, xNnz), 256, 0, h->currentStream>>>(     \
			y, xNnz, xIndices, xValues, xBaseIndex, beta);                      \
		spgpu_count_launch(h);                                                  \
	}

SPGPU_DEFINE_SPVEC(I, int)
SPGPU_DEFINE_SPVEC(S, float)
SPGPU_DEFINE_SPVEC(D, double)
SPGPU_DEFINE_SPVEC(C, cuFloatComplex)
SPGPU_DEFINE_SPVEC(Z, cuDoubleComplex)

/*
 * ELL coefficient update, reference ell_csput_base.cuh:33-75: for each triple
 * (aI, aJ, aVal) binary-search column aJ among the (sorted) rS[row] columns of
 * row aI - baseIndex and overwrite that slot's value.  As in the reference,
 * alpha is accepted but unused and aJ is compared to the stored index as is.
 */
template <typename T>
__global__ void __launch_bounds__(256)
ellcsput_kernel(T* cM, const int* __restrict__ rP, int cMPitch, int rPPitch,
	const int* __restrict__ rS, int nnz, const int* __restrict__ aI,
	const int* __restrict__ aJ, const T* __restrict__ aVal, int baseIndex)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nnz)
		return;
	const long long row = (long long)aI[i] - baseIndex;
	if (row < 0)
		return;
	const int column = aJ[i];
	int lo = 0, hi = rS[row] - 1;
	while (lo <= hi) {
		const int mid = (lo + hi) >> 1;
		const int c = rP[row + (long long)mid * rPPitch];
		if (c == column) {
			cM[row + (long long)mid * cMPitch] = aVal[i];
			return;
		}
		if (c < column) lo = mid + 1; else hi = mid - 1;
	}
}

#define SPGPU_DEFINE_ELLCSPUT(S, T)                                            \
	extern "C" void spgpu##S##ellcsput(spgpuHandle_t h, T alpha, T* cM,         \
		const int* rP, int cMPitch, int rPPitch, const int* rS, int nnz,        \
		int* aI, int* aJ, T* aVal, int baseIndex)                               \
	{                                                                           \
		(void)alpha;                                                            \
		if (nnz <= 0) return;                                                   \
		ellcsput_kernel<T><<<(nnz + 255) / 256, 256, 0, h->currentStream>>>(    \
			cM, rP, cMPitch, rPPitch, rS, nnz, aI, aJ, aVal, baseIndex);        \
		spgpu_count_launch(h);                                                  \
	}

SPGPU_DEFINE_ELLCSPUT(S, float)
SPGPU_DEFINE_ELLCSPUT(D, double)
SPGPU_DEFINE_ELLCSPUT(C, cuFloatComplex)
SPGPU_DEFINE_ELLCSPUT(Z, cuDoubleComplex)
