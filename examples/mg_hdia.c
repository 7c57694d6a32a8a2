/*
 * A C-only consumer of the HDIA side of the multi-GPU API (include/spgpu_mg.h): BASELINE configs[1] -- the 3-D
 * 27-point stencil n^3 in double HDIA, hackSize 32 -- built with the library's own host conversion (cooToHdia,
 * reference hdia_conv.h:45-70), handed over whole with spgpuMgDhdiaCreate, multiplied on every device of the box with
 * the halo exchange inside the SpMV kernel, checked against the stencil applied on the host, then solved with CG.
 *
 *   gcc -O2 -fopenmp examples/mg_hdia.c -Iinclude -I/usr/local/cuda/include -Lspgpu_b200/lib -lspgpu \
 *       -Wl,-rpath,$PWD/spgpu_b200/lib -L/usr/local/cuda/lib64 -lcudart -lm -o examples/mg_hdia
 *   examples/mg_hdia [n = 48] [ranks = all devices] [spmv repetitions = 20] [cg iterations = 30]
 *
 * Exit status 0 = every row within 1e-12 of the host stencil (relative to sum |a_ik||x_k|) and CG reduced r.r.
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

/* 27-point stencil: 26 on the diagonal, -1 to every in-grid neighbour of the 3x3x3 cube (symmetric, diagonally dominant) */
static double weight(int dz, int dy, int dx)
{
	return (dz == 0 && dy == 0 && dx == 0) ? 26.0 : -1.0;
}

int main(int argc, char** argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 48;
	int ndev = 0, ranks, reps, iters, r, devices[16];
	cudaGetDeviceCount(&ndev);
	ranks = argc > 2 ? atoi(argv[2]) : ndev;
	reps = argc > 3 ? atoi(argv[3]) : 20;
	iters = argc > 4 ? atoi(argv[4]) : 30;
	if (ndev < 1 || n < 2 || n > 400 || ranks < 1 || ranks > 16) {
		fprintf(stderr, "usage: mg_hdia [n <= 400] [ranks <= 16] [reps] [cg iterations]\n");
		return 2;
	}
	{
		const long long plane = (long long)n * n, N = plane * n;
		const int rows = (int)N, hackSize = 32;
		long long nnz = 0, at = 0, g;
		int *ci, *cj, *hackOffsets, *offsets, height = 0, hacks;
		double *cv, *dM, *x, *z, *want, *scale, worst = 0.0, t0, dt, rr0 = 0.0, rr = 0.0;
		long long bad = 0;
		unsigned long long state = 4321;
		int it;
		spgpuMgHandle_t mg;
		spgpuMgMatrix_t A;
		spgpuMgVector_t vx, vz, vb;
		spgpuMgCg_t cg;

		/* COO of the stencil, row-major, columns ascending */
		for (g = 0; g < N; ++g) {
			const int xx = (int)(g % n), yy = (int)((g / n) % n), zz = (int)(g / plane);
			int dz, dy, dx;
			for (dz = -1; dz <= 1; ++dz) for (dy = -1; dy <= 1; ++dy) for (dx = -1; dx <= 1; ++dx)
				if (zz + dz >= 0 && zz + dz < n && yy + dy >= 0 && yy + dy < n && xx + dx >= 0 && xx + dx < n)
					++nnz;
		}
		ci = (int*)malloc((size_t)nnz * sizeof(int));
		cj = (int*)malloc((size_t)nnz * sizeof(int));
		cv = (double*)malloc((size_t)nnz * sizeof(double));
		for (g = 0; g < N; ++g) {
			const int xx = (int)(g % n), yy = (int)((g / n) % n), zz = (int)(g / plane);
			int dz, dy, dx;
			for (dz = -1; dz <= 1; ++dz) for (dy = -1; dy <= 1; ++dy) for (dx = -1; dx <= 1; ++dx)
				if (zz + dz >= 0 && zz + dz < n && yy + dy >= 0 && yy + dy < n && xx + dx >= 0 && xx + dx < n) {
					ci[at] = (int)g;
					cj[at] = (int)(g + dz * plane + dy * n + dx);
					cv[at] = weight(dz, dy, dx);
					++at;
				}
		}

		/* the library's own host conversion: COO -> HDIA */
		t0 = now();
		hacks = getHdiaHacksCount(hackSize, rows);
		hackOffsets = (int*)malloc((size_t)(hacks + 1) * sizeof(int));
		computeHdiaHackOffsetsFromCoo(&height, hackOffsets, hackSize, rows, rows, (int)nnz, ci, cj, 0);
		dM = (double*)calloc((size_t)height * hackSize, sizeof(double));
		offsets = (int*)malloc((size_t)height * sizeof(int));
		cooToHdia(dM, offsets, hackOffsets, hackSize, rows, rows, (int)nnz, ci, cj, cv, 0, SPGPU_TYPE_DOUBLE);
		printf("%d^3 27-point stencil: %d rows, %lld non-zeros, HDIA hackSize 32: %d hacks, %d hack-diagonals (cooToHdia on the host: %.2f s)\n",
			n, rows, nnz, hacks, height, now() - t0);
		free(ci); free(cj); free(cv);

		for (r = 0; r < ranks; ++r)
			devices[r] = r % ndev;
		OK(spgpuMgCreate(&mg, devices, ranks));
		{
			int plan[17], halo = 0, fits = 0;
			OK(spgpuMgHdiaPlan(ranks, SPGPU_TYPE_DOUBLE, dM, offsets, hackSize, hackOffsets, rows, rows, plan, &halo, &fits));
			printf("%d rank(s) on %d device(s), exchange: %s; plan: halo %d entries, %s\n", ranks, ndev,
				spgpuMgExchange(mg) == SPGPU_MG_FUSED ? "fused into the SpMV kernel (NVLink peer stores)" : "push kernels + CUDA events",
				halo, fits ? "fits" : "does NOT fit a neighbouring block");
			if (!fits) {
				spgpuMgDestroy(mg);
				return 2;
			}
		}
		OK(spgpuMgDhdiaCreate(mg, &A, dM, offsets, hackSize, hackOffsets, rows, rows));
		free(dM); free(offsets); free(hackOffsets);

		x = (double*)malloc((size_t)N * sizeof(double));
		z = (double*)malloc((size_t)N * sizeof(double));
		want = (double*)malloc((size_t)N * sizeof(double));
		scale = (double*)malloc((size_t)N * sizeof(double));
		for (g = 0; g < N; ++g) {
			state = state * 6364136223846793005ull + 1442695040888963407ull;
			x[g] = (double)(state >> 11) / 9007199254740992.0;
		}
#pragma omp parallel for schedule(static)
		for (long long q = 0; q < N; ++q) {
			const int xx = (int)(q % n), yy = (int)((q / n) % n), zz = (int)(q / plane);
			double acc = 0.0, sc = 0.0;
			int dz, dy, dx;
			for (dz = -1; dz <= 1; ++dz) for (dy = -1; dy <= 1; ++dy) for (dx = -1; dx <= 1; ++dx)
				if (zz + dz >= 0 && zz + dz < n && yy + dy >= 0 && yy + dy < n && xx + dx >= 0 && xx + dx < n) {
					const double a = weight(dz, dy, dx), v = x[q + dz * plane + dy * n + dx];
					acc = fma(a, v, acc);
					sc += fabs(a) * fabs(v);
				}
			want[q] = acc;
			scale[q] = sc;
		}
		OK(spgpuMgVectorCreate(A, &vx));
		OK(spgpuMgVectorCreate(A, &vz));
		OK(spgpuMgVectorCreate(A, &vb));
		OK(spgpuMgVectorSet(vx, x));
		for (it = 0; it < 3; ++it)
			OK(spgpuMgDhdiaspmv(mg, vz, NULL, 1.0, A, vx, 0.0));
		OK(spgpuMgSynchronize(mg));
		OK(spgpuMgVectorGet(vz, z));
		for (g = 0; g < N; ++g) {
			const double e = fabs(z[g] - want[g]) / (scale[g] > 0 ? scale[g] : 1.0);
			if (e > worst) worst = e;
			if (!(e <= 1e-12)) ++bad;
		}
		printf("SpMV: worst row error %.3g relative to sum|a_ik||x_k| (%lld rows over 1e-12)\n", worst, bad);
		t0 = now();
		for (it = 0; it < reps; ++it)
			OK(spgpuMgDspmv(mg, vz, NULL, 1.0, A, vx, 0.0));       /* the format-agnostic entry point */
		OK(spgpuMgSynchronize(mg));
		dt = (now() - t0) / (reps > 0 ? reps : 1);
		printf("SpMV: %.4f ms per product, %.1f GFLOP/s (host clock over %d products)\n", dt * 1e3, 2.0 * nnz / dt / 1e9, reps);

		OK(spgpuMgDhdiaspmv(mg, vb, NULL, 1.0, A, vx, 0.0));
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
		spgpuMgMatrixDestroy(A);
		spgpuMgDestroy(mg);
		free(x); free(z); free(want); free(scale);
		printf(bad ? "FAILED\n" : "OK\n");
		return bad ? 1 : 0;
	}
}
