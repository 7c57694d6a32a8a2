/*
 * How fast can an SM gather random 4-byte (or 8-byte) elements of an L2-resident vector, per load flavour?
 * cfg3 (BASELINE configs[2]: uniformly random columns over 4 M floats) makes every non-zero such a gather; ncu shows
 * l1tex__throughput 94 % on the sorted matrix's kernel.  This probe isolates the gather: idx[] (64 M random ints, read
 * coalesced with .cs) selects elements of x (16 MB float / 32 MB double), 8 gathers in flight per thread.
 * Output: gathers per ns, and per SM per clock at the measured SM clock.
 *
 * build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
 */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int F, typename T> __device__ __forceinline__ T ld(const T* p);
#define DEF(F, Q)                                                                                                   \
	template <> __device__ __forceinline__ float ld<F, float>(const float* p)                                       \
	{ float v; asm volatile("ld.global" Q ".f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }                         \
	template <> __device__ __forceinline__ double ld<F, double>(const double* p)                                    \
	{ double v; asm volatile("ld.global" Q ".f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
DEF(0, ".ca")
DEF(1, ".cg")
DEF(2, ".nc")
DEF(3, ".L1::no_allocate")
DEF(4, ".nc.L1::no_allocate")
DEF(5, ".cv")
DEF(6, "")
static const char* NAMES[7] = { "ld.global.ca", "ld.global.cg", "ld.global.nc", "ld.global.L1::no_allocate",
	"ld.global.nc.L1::no_allocate", "ld.global.cv", "ld.global" };

template <int F, typename T>
__global__ void __launch_bounds__(256) gather(const int* __restrict__ idx, const T* x, long long n, T* out)
{
	const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long nthreads = (long long)gridDim.x * blockDim.x;
	T acc = 0;
	for (long long i = tid; i < n; i += 8 * nthreads) {
		int c[8];
		T v[8];
#pragma unroll
		for (int u = 0; u < 8; ++u)
			c[u] = (i + u * nthreads < n) ? __ldcs(idx + i + u * nthreads) : 0;
#pragma unroll
		for (int u = 0; u < 8; ++u)
			v[u] = ld<F, T>(x + c[u]);
#pragma unroll
		for (int u = 0; u < 8; ++u)
			acc += v[u];
	}
	if (acc == (T)12345.678)
		out[0] = acc;
}

template <int F, typename T>
static void run(const int* idx, const T* x, long long n, T* out, int sms, int ctasPerSm, double mhz, const char* pattern)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	float best = 1e30f;
	for (int it = 0; it < 5; ++it) {
		cudaEventRecord(a);
		gather<F, T><<<sms * ctasPerSm, 256>>>(idx, x, n, out);
		cudaEventRecord(b);
		cudaEventSynchronize(b);
		float ms;
		cudaEventElapsedTime(&ms, a, b);
		if (it > 0 && ms < best) best = ms;
	}
	printf("{\"pattern\": \"%s\", \"load\": \"%s\", \"bytes\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"gathers_per_ns\": %.2f, "
		"\"per_sm_per_clock\": %.3f}\n", pattern, NAMES[F], (int)sizeof(T), ctasPerSm, best, n / (best * 1e6),
		n / (best * 1e-3) / sms / (mhz * 1e6));
	fflush(stdout);
}

int main()
{
	const long long n = 64ll << 20;
	const int xn = 4 << 20;
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1e3;
	std::vector<int> h(n);
	unsigned long long s = 88172645463325252ull;
	int *idx;
	float* xf;
	double *xd, *out;
	cudaMalloc(&idx, n * sizeof(int));
	cudaMalloc(&xf, xn * sizeof(float));
	cudaMalloc(&xd, xn * sizeof(double));
	cudaMalloc(&out, 64);
	cudaMemset(xf, 0, xn * sizeof(float));
	cudaMemset(xd, 0, xn * sizeof(double));
	for (int pattern = 0; pattern < 3; ++pattern) {
		/* 0: uniformly random (cfg3); 1: random 128-byte line, the 32 lanes of a warp in the SAME line (one tag lookup per
		 * warp); 2: random line per group of 8 consecutive threads (4 lookups per warp) */
		for (long long i = 0; i < n; ++i) {
			const long long key = pattern == 0 ? i : pattern == 1 ? (i / 32) : (i / 8);
			unsigned long long z = (unsigned long long)key * 0x9E3779B97F4A7C15ull + s;
			z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
			const int line = (int)(z % (xn / 32));
			h[i] = pattern == 0 ? (int)(z % xn) : line * 32 + (int)(i % (pattern == 1 ? 32 : 8));
		}
		cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
		const char* pn = pattern == 0 ? "every lane its own line" : pattern == 1 ? "one line per warp" : "one line per 8 lanes";
		for (int c = 8; c >= 4; c -= 4) {
			run<0, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
			run<1, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
			run<2, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
			run<3, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
			run<4, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
			run<5, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
			run<6, float>(idx, xf, n, (float*)out, p.multiProcessorCount, c, mhz, pn);
		}
		run<2, double>(idx, xd, n, out, p.multiProcessorCount, 8, mhz, pn);
		run<1, double>(idx, xd, n, out, p.multiProcessorCount, 8, mhz, pn);
	}
	return 0;
}
