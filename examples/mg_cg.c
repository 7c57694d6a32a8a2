/*
 * A C-only consumer of the multi-GPU API (include/spgpu_mg.h): BASELINE configs[4] -- the 3-D 7-point
 * Laplacian n^3 in double HELL, hackSize 32, row-sharded by z-slabs over the GPUs of one box, one
 * x halo plane per neighbour -- with no Python anywhere: ONE process, one rank per device.
 *
 *   gcc -O2 -fopenmp examples/mg_cg.c -Iinclude -I/usr/local/cuda/include -Lspgpu_b200/lib -lspgpu \
 *       -Wl,-rpath,$PWD/spgpu_b200/lib -L/usr/local/cuda/lib64 -lcudart -lm -o examples/mg_cg
 *   examples/mg_cg [n = 128] [ranks = all devices] [spmv repetitions = 20] [cg iterations = 50]
 *
 * Every rank's slab is assembled on the host directly in the partitioned form (local column indices
 * into x_ext = [plane | owned planes | plane]) and handed over with spgpuMgHellCreateFromBlocks --
 * the 512^3 matrix (11 GB) never exists as one global array.  The program multiplies, checks every
 * row against the stencil applied on the host, times the product, then runs CG on A x = b with a
 * known solution.  Exit status 0 = rows within 1e-12 (relative to sum |a_ik||x_k|) and CG converged.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda_runtime_api.h>

#include "spgpu_mg.h"

#define OK(c) do { spgpuStatus_t s_ = (c); if (s_ != SPGPU_SUCCESS) { \
	fprintf(stderr, "%s:%d: spgpu status %d\n", __FILE__, __LINE__, (int)s_); exit(2); } } while (0)

static double now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* rows of planes [z0, z1) of the n^3 7-point Laplacian (6 on the diagonal, -1 to the in-grid neighbours, columns
 * ascending) as a HELL block whose column indices address x_ext = [one plane | (z1-z0) planes | one plane] */
typedef struct Block { int rows; long long elements; double* cM; int* rP; int* hackOffsets; int* rS; } Block;

static Block build_slab(int n, int z0, int z1)
{
	const long long plane = (long long)n * n;
	const int rows = (int)((z1 - z0) * plane), hacks = (rows + 31) / 32;
	const long long lo = z0 * plane;
	Block b;
	int h;
	b.rows = rows;
	b.rS = (int*)malloc((size_t)(rows ? rows : 1) * sizeof(int));
	b.hackOffsets = (int*)malloc((size_t)(hacks ? hacks : 1) * sizeof(int));
#pragma omp parallel for schedule(static)
	for (int i = 0; i < rows; ++i) {
		const long long g = lo + i;
		const int x = (int)(g % n), y = (int)((g / n) % n), z = (int)(g / plane);
		b.rS[i] = 1 + (x > 0) + (x < n - 1) + (y > 0) + (y < n - 1) + (z > 0) + (z < n - 1);
	}
	b.elements = 0;
	for (h = 0; h < hacks; ++h) {
		int deepest = 0, i;
		for (i = h * 32; i < rows && i < (h + 1) * 32; ++i)
			if (b.rS[i] > deepest)
				deepest = b.rS[i];
		b.hackOffsets[h] = (int)b.elements;
		b.elements += 32ll * deepest;
	}
	b.cM = (double*)calloc((size_t)(b.elements ? b.elements : 1), sizeof(double));
	b.rP = (int*)calloc((size_t)(b.elements ? b.elements : 1), sizeof(int));
#pragma omp parallel for schedule(static)
	for (int i = 0; i < rows; ++i) {
		const long long g = lo + i;
		const int x = (int)(g % n), y = (int)((g / n) % n), z = (int)(g / plane);
		const long long at = (long long)b.hackOffsets[i / 32] + i % 32;
		const long long shift = lo - plane;                 /* x_ext position of global column c: c - (lo - plane) */
		int k = 0;
#define PUT(col, val) do { b.cM[at + 32ll * k] = (val); b.rP[at + 32ll * k] = (int)((col) - shift); ++k; } while (0)
		if (z > 0) PUT(g - plane, -1.0);
		if (y > 0) PUT(g - n, -1.0);
		if (x > 0) PUT(g - 1, -1.0);
		PUT(g, 6.0);
		if (x < n - 1) PUT(g + 1, -1.0);
		if (y < n - 1) PUT(g + n, -1.0);
		if (z < n - 1) PUT(g + plane, -1.0);
#undef PUT
	}
	return b;
}

static void stencil_host(int n, const double* x, double* z, double* scale)
{
	const long long plane = (long long)n * n, N = plane * n;
#pragma omp parallel for schedule(static)
	for (long long g = 0; g < N; ++g) {
		const int xx = (int)(g % n), y = (int)((g / n) % n), zz = (int)(g / plane);
		double acc = 0.0, sc = 0.0;
#define TERM(c, v) do { acc = fma((v), x[c], acc); sc += fabs(v) * fabs(x[c]); } while (0)
		if (zz > 0) TERM(g - plane, -1.0);
		if (y > 0) TERM(g - n, -1.0);
		if (xx > 0) TERM(g - 1, -1.0);
		TERM(g, 6.0);
		if (xx < n - 1) TERM(g + 1, -1.0);
		if (y < n - 1) TERM(g + n, -1.0);
		if (zz < n - 1) TERM(g + plane, -1.0);
#undef TERM
		z[g] = acc;
		scale[g] = sc;
	}
}

int main(int argc, char** argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 128;
	int ndev = 0, ranks, reps, iters, r, i, devices[16];
	spgpuMgHandle_t mg;
	spgpuMgMatrix_t A;
	spgpuMgVector_t vx, vz, vb;
	spgpuMgCg_t cg;
	Block blocks[16];
	int blockRows[16];
	long long elements[16], nnz = 0;
	const void* cM[16];
	const int *rP[16], *ho[16], *rS[16];
	cudaGetDeviceCount(&ndev);
	ranks = argc > 2 ? atoi(argv[2]) : ndev;
	reps = argc > 3 ? atoi(argv[3]) : 20;
	iters = argc > 4 ? atoi(argv[4]) : 50;
	if (ndev < 1 || ranks < 1 || ranks > 16 || n % 8 != 0 || n % ranks != 0 || (long long)n * n * n / ranks >= 2147483647ll) {
		fprintf(stderr, "usage: mg_cg [n (multiple of 8 and of the ranks)] [ranks <= 16] [reps] [cg iterations]\n");
		return 2;
	}
	for (r = 0; r < ranks; ++r)
		devices[r] = r % ndev;                              /* more ranks than devices: several ranks per device (EVENTS mode) */
	OK(spgpuMgCreate(&mg, devices, ranks));
	printf("%d^3 7-point Laplacian, double HELL hackSize 32, %d rank(s) on %d device(s), exchange: %s\n", n, ranks, ndev,
		spgpuMgExchange(mg) == SPGPU_MG_FUSED ? "fused into the SpMV kernel (NVLink peer stores)" : "push kernels + CUDA events");

	{
		const double t0 = now();
		for (r = 0; r < ranks; ++r) {
			blocks[r] = build_slab(n, r * (n / ranks), (r + 1) * (n / ranks));
			blockRows[r] = blocks[r].rows;
			elements[r] = blocks[r].elements;
			cM[r] = blocks[r].cM; rP[r] = blocks[r].rP; ho[r] = blocks[r].hackOffsets; rS[r] = blocks[r].rS;
			for (i = 0; i < blocks[r].rows; ++i)
				nnz += blocks[r].rS[i];
		}
		OK(spgpuMgHellCreateFromBlocks(mg, &A, SPGPU_TYPE_DOUBLE, 32, n * n, 0, 7, blockRows, cM, rP, ho, rS, elements, 0));
		for (r = 0; r < ranks; ++r) {
			free(blocks[r].cM); free(blocks[r].rP); free(blocks[r].hackOffsets); free(blocks[r].rS);
		}
		printf("assembled and uploaded %lld rows, %lld non-zeros in %.2f s\n", (long long)spgpuMgMatrixRows(A), nnz, now() - t0);
	}

	{
		const long long N = (long long)n * n * n;
		double* x = (double*)malloc((size_t)N * sizeof(double));
		double* z = (double*)malloc((size_t)N * sizeof(double));
		double* want = (double*)malloc((size_t)N * sizeof(double));
		double* scale = (double*)malloc((size_t)N * sizeof(double));
		double worst = 0.0, t0, dt, rr0 = 0.0, rr = 0.0;
		unsigned long long state = 12345;
		long long g, bad = 0;
		int it;
		for (g = 0; g < N; ++g) {                            /* splitmix-style LCG, U(0,1) */
			state = state * 6364136223846793005ull + 1442695040888963407ull;
			x[g] = (double)(state >> 11) / 9007199254740992.0;
		}
		OK(spgpuMgVectorCreate(A, &vx));
		OK(spgpuMgVectorCreate(A, &vz));
		OK(spgpuMgVectorCreate(A, &vb));
		OK(spgpuMgVectorSet(vx, x));
		for (it = 0; it < 3; ++it)
			OK(spgpuMgDhellspmv(mg, vz, NULL, 1.0, A, vx, 0.0));
		OK(spgpuMgSynchronize(mg));
		OK(spgpuMgVectorGet(vz, z));
		stencil_host(n, x, want, scale);
		for (g = 0; g < N; ++g) {
			const double e = fabs(z[g] - want[g]) / (scale[g] > 0 ? scale[g] : 1.0);
			if (e > worst) worst = e;
			if (!(e <= 1e-12)) ++bad;
		}
		printf("SpMV: worst row error %.3g relative to sum|a_ik||x_k| (%lld rows over 1e-12)\n", worst, bad);

		t0 = now();
		for (it = 0; it < reps; ++it)
			OK(spgpuMgDhellspmv(mg, vz, NULL, 1.0, A, vx, 0.0));
		OK(spgpuMgSynchronize(mg));
		dt = (now() - t0) / reps;
		printf("SpMV: %.4f ms per product, %.1f GFLOP/s, %.1f GB/s algorithmic (host clock over %d products)\n", dt * 1e3,
			2.0 * nnz / dt / 1e9, (12.0 * nnz + 4.0 * N + 4.0 * (N / 32) + 16.0 * N) / dt / 1e9, reps);

		/* CG on A u = b with b = A x: the iteration must drive r.r down */
		OK(spgpuMgDhellspmv(mg, vb, NULL, 1.0, A, vx, 0.0));
		OK(spgpuMgDcgCreate(A, &cg));
		OK(spgpuMgDcgStart(cg, vb, &rr0));
		OK(spgpuMgDcgStep(cg, 2, &rr));                      /* untimed: the first launches of a process are not the iteration's cost */
		t0 = now();
		OK(spgpuMgDcgStep(cg, iters, &rr));
		dt = (now() - t0) / (iters > 0 ? iters : 1);
		printf("CG: r.r %.6e -> %.6e after 2 + %d iterations, %.4f ms per iteration\n", rr0, rr, iters, dt * 1e3);
		if (!(rr < rr0))
			++bad;
		spgpuMgDcgDestroy(cg);
		spgpuMgVectorDestroy(vx); spgpuMgVectorDestroy(vz); spgpuMgVectorDestroy(vb);
		free(x); free(z); free(want); free(scale);
		spgpuMgMatrixDestroy(A);
		spgpuMgDestroy(mg);
		printf(bad ? "FAILED\n" : "OK\n");
		return bad ? 1 : 0;
	}
}
