/*
 * How much DRAM traffic does a PARTIALLY used 128-byte line cost on B200, per load flavour?
 * (cfg3's power-law HELL reads 2.84 GB from DRAM where 32-byte sectors would need 1.37 GB.)
 *
 * Each warp walks 128-byte lines of a 1 GiB float buffer; in every line only the lanes selected by
 * `mask8` (one bit per 8-lane = 32-byte sector) load their element.  Run under
 *   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
 * and compare bytes read with lines x used sectors x 32.
 *
 * build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_probe sector_probe.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int F> __device__ __forceinline__ float ld(const float* p);
template <> __device__ __forceinline__ float ld<0>(const float* p) { float v; asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<1>(const float* p) { float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<2>(const float* p) { float v; asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<3>(const float* p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<4>(const float* p) { float v; asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<5>(const float* p) { float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<6>(const float* p) { float v; asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<7>(const float* p) { float v; asm volatile("ld.global.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<8>(const float* p) { float v; asm volatile("ld.global.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
template <> __device__ __forceinline__ float ld<9>(const float* p) { float v; asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

template <> __device__ __forceinline__ float ld<10>(const float* p) { float v; asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

static const char* NAMES[11] = { "ld.global.ca", "ld.global.cg", "ld.global.cs", "ld.global.nc", "ld.global.L1::no_allocate",
	"ld.global.nc.L1::no_allocate", "ld.global.L2::64B", "ld.global.L2::128B", "ld.global.L1::evict_first", "ld.volatile.global", "ld.global" };
static const char* UNUSED_NAMES[10] = { "ld.global.ca", "ld.global.cg", "ld.global.cs", "ld.global.nc", "ld.global.L1::no_allocate",
	"ld.global.nc.L1::no_allocate", "ld.global.L2::64B", "ld.global.L2::128B", "ld.global.L1::evict_first", "ld.volatile.global" };

/* kernel name carries flavour and mask so that the ncu CSV is self-describing */
template <int F, int MASK>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ buf, long long lines, float* out)
{
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	const bool on = (MASK >> (lane >> 3)) & 1;
	float acc = 0.f;
	for (long long l = warp; l < lines; l += 4 * nwarps) {
		float v[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const long long m = l + u * nwarps;
			if (on && m < lines)
				v[u] = ld<F>(buf + m * 32 + lane);
		}
		acc += (v[0] + v[1]) + (v[2] + v[3]);
	}
	if (acc == 12345.678f)
		out[0] = acc;
}

template <int F, int MASK>
static void run(const float* buf, long long lines, float* out, int sms)
{
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	probe<F, MASK><<<sms * 8, 256>>>(buf, lines, out);
	cudaEventRecord(a);
	probe<F, MASK><<<sms * 8, 256>>>(buf, lines, out);
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms = 0.f;
	cudaEventElapsedTime(&ms, a, b);
	const int used = __builtin_popcount(MASK);
	printf("{\"flavour\": \"%s\", \"sector_mask\": %d, \"sectors_used_per_line\": %d, \"ms\": %.4f, \"useful_gbs\": %.1f, \"line_gbs\": %.1f}\n",
		NAMES[F], MASK, used, ms, lines * used * 32.0 / ms / 1e6, lines * 128.0 / ms / 1e6);
}

template <int F> static void run_masks(const float* buf, long long lines, float* out, int sms)
{
	run<F, 0x1>(buf, lines, out, sms);     /* 1 sector of 4            */
	run<F, 0x5>(buf, lines, out, sms);     /* 2 sectors, not adjacent  */
	run<F, 0x3>(buf, lines, out, sms);     /* 2 adjacent sectors       */
	run<F, 0xF>(buf, lines, out, sms);     /* the whole line           */
}

int main()
{
	const long long bytes = 1ll << 30, lines = bytes / 128;
	float *buf, *out;
	cudaMalloc(&buf, bytes);
	cudaMalloc(&out, 4);
	cudaMemset(buf, 0, bytes);
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	run_masks<0>(buf, lines, out, sms);
	run_masks<1>(buf, lines, out, sms);
	run_masks<2>(buf, lines, out, sms);
	run_masks<3>(buf, lines, out, sms);
	run_masks<4>(buf, lines, out, sms);
	run_masks<5>(buf, lines, out, sms);
	run_masks<6>(buf, lines, out, sms);
	run_masks<7>(buf, lines, out, sms);
	run_masks<8>(buf, lines, out, sms);
	run_masks<9>(buf, lines, out, sms);
	run_masks<10>(buf, lines, out, sms);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
