/*
 * A C-only consumer of the library, in the shape of the reference's perf drivers
 * (reference src/tests/hellPerf.cpp / diaPerf.cpp): read a MatrixMarket file, unfold it if
 * symmetric, convert COO -> ELL -> HELL and COO -> HDIA with the reference's own conversion
 * calls, multiply on the GPU through spgpuDellspmv / spgpuDhellspmv / spgpuDhdiaspmv, check every
 * result against a serial host loop over the COO entries and print the timings.
 *
 *   gcc -O2 examples/spmv_mtx.c -Iinclude -I/usr/local/cuda/include \
 *       -Lspgpu_b200/lib -lspgpu -Wl,-rpath,$PWD/spgpu_b200/lib -L/usr/local/cuda/lib64 -lcudart -lm -o spmv_mtx
 *   ./spmv_mtx matrix.mtx [repetitions]
 *
 * Exit status 0 = all three formats agree with the host loop within 1e-12 per row (scaled by
 * sum |a_ik||x_k|), the tolerance BASELINE.json states for double.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "spgpu.h"
#include "spgpu_mm.h"

#define CHECK(c) do { cudaError_t e_ = (c); if (e_ != cudaSuccess) { \
	fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

static void* to_device(const void* host, size_t bytes)
{
	void* d = NULL;
	CHECK(cudaMalloc(&d, bytes ? bytes : 1));
	if (bytes)
		CHECK(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
	return d;
}

static double check(const char* name, const double* z, const double* want, const double* scale, int rows)
{
	double worst = 0.0;
	for (int i = 0; i < rows; ++i) {
		const double err = fabs(z[i] - want[i]) / (scale[i] > 0.0 ? scale[i] : 1.0);
		if (err > worst) worst = err;
	}
	printf("  %-5s max scaled row error %.3e %s\n", name, worst, worst <= 1e-12 ? "ok" : "FAILED");
	return worst;
}

int main(int argc, char** argv)
{
	if (argc < 2) {
		fprintf(stderr, "usage: %s matrix.mtx [repetitions]\n", argv[0]);
		return 2;
	}
	const int reps = argc > 2 ? atoi(argv[2]) : 20;
	spgpuMmProperties pr;
	if (!spgpuMmLoadProperties(argv[1], &pr) || !pr.isStoredSparse) {
		fprintf(stderr, "%s: not a MatrixMarket coordinate matrix\n", argv[1]);
		return 2;
	}
	int nnz = pr.nonZerosCount;
	int* rows = malloc(sizeof(int) * (nnz ? nnz : 1));
	int* cols = malloc(sizeof(int) * (nnz ? nnz : 1));
	double* vals = malloc(sizeof(double) * (nnz ? nnz : 1));
	if (spgpuMmLoadMatrixToCoo(argv[1], vals, rows, cols, SPGPU_TYPE_DOUBLE) != MATRIX_READ_SUCCESS) {
		fprintf(stderr, "%s: cannot read the entries as double\n", argv[1]);
		return 2;
	}
	if (pr.matrixType == MATRIX_TYPE_SYMMETRIC) {
		const int n = spgpuMmUnfoldedSymmetricSize(vals, rows, cols, nnz, SPGPU_TYPE_DOUBLE);
		int* ur = malloc(sizeof(int) * (n ? n : 1));
		int* uc = malloc(sizeof(int) * (n ? n : 1));
		double* uv = malloc(sizeof(double) * (n ? n : 1));
		spgpuMmUnfoldSymmetric(ur, uc, uv, rows, cols, vals, nnz, SPGPU_TYPE_DOUBLE);
		free(rows); free(cols); free(vals);
		rows = ur; cols = uc; vals = uv; nnz = n;
	}
	const int R = pr.rowsCount, C = pr.columnsCount;
	printf("%s: %d x %d, %d non-zeros%s\n", argv[1], R, C, nnz, pr.matrixType == MATRIX_TYPE_SYMMETRIC ? " (symmetric, unfolded)" : "");

	/* x, and the host answer straight from the COO entries */
	double* x = malloc(sizeof(double) * (C ? C : 1));
	unsigned long long state = 12345;
	for (int i = 0; i < C; ++i) {
		state = state * 6364136223846793005ULL + 1442695040888963407ULL;
		x[i] = (double)(state >> 11) / 9007199254740992.0;
	}
	double* want = calloc(R ? R : 1, sizeof(double));
	double* scale = calloc(R ? R : 1, sizeof(double));
	for (int e = 0; e < nnz; ++e) {
		want[rows[e]] += vals[e] * x[cols[e]];
		scale[rows[e]] += fabs(vals[e]) * fabs(x[cols[e]]);
	}

	spgpuHandle_t h;
	if (spgpuCreate(&h, 0) != SPGPU_SUCCESS) {
		fprintf(stderr, "spgpuCreate failed (no GPU?)\n");
		return 2;
	}
	double* dx = to_device(x, sizeof(double) * C);
	double* dz = NULL;
	CHECK(cudaMalloc((void**)&dz, sizeof(double) * (R ? R : 1)));
	double* z = malloc(sizeof(double) * (R ? R : 1));
	cudaEvent_t t0, t1;
	CHECK(cudaEventCreate(&t0));
	CHECK(cudaEventCreate(&t1));
	cudaStream_t stream = spgpuGetStream(h);
	float ms;
	double worst = 0.0;

	/* ---- ELL and HELL (reference ell_conv.h / hell_conv.h) -------------------------------- */
	int* rs = calloc(R ? R : 1, sizeof(int));
	int maxLen = 0;
	computeEllRowLenghts(rs, &maxLen, R, nnz, rows, 0);
	const int pitch = computeEllAllocPitch(R);
	double* ellV = calloc((size_t)pitch * (maxLen ? maxLen : 1), sizeof(double));
	int* ellI = calloc((size_t)pitch * (maxLen ? maxLen : 1), sizeof(int));
	cooToEll(ellV, ellI, pitch, pitch, maxLen, 0, R, nnz, rows, cols, vals, 0, SPGPU_TYPE_DOUBLE);
	const int avg = R ? (nnz + R - 1) / R : 1;
	{
		double* dV = to_device(ellV, sizeof(double) * pitch * maxLen);
		int* dI = to_device(ellI, sizeof(int) * pitch * maxLen);
		int* dRs = to_device(rs, sizeof(int) * R);
		spgpuDellspmv(h, dz, NULL, 1.0, dV, dI, pitch, pitch, dRs, NULL, avg, maxLen, R, dx, 0.0, 0);
		CHECK(cudaEventRecord(t0, stream));
		for (int k = 0; k < reps; ++k)
			spgpuDellspmv(h, dz, NULL, 1.0, dV, dI, pitch, pitch, dRs, NULL, avg, maxLen, R, dx, 0.0, 0);
		CHECK(cudaEventRecord(t1, stream));
		CHECK(cudaEventSynchronize(t1));
		CHECK(cudaEventElapsedTime(&ms, t0, t1));
		CHECK(cudaMemcpy(z, dz, sizeof(double) * R, cudaMemcpyDeviceToHost));
		printf("ELL   %9.3f us / SpMV  %8.1f GFLOP/s\n", 1e3 * ms / reps, 2.0 * nnz / (1e6 * ms / reps));
		worst = fmax(worst, check("ELL", z, want, scale, R));

		const int hack = 32;
		int hellHeight = 0;
		computeHellAllocSize(&hellHeight, hack, R, rs);
		const int hacks = (R + hack - 1) / hack;
		double* hellV = calloc((size_t)hellHeight * hack + 1, sizeof(double));
		int* hellI = calloc((size_t)hellHeight * hack + 1, sizeof(int));
		int* hoff = calloc(hacks + 1, sizeof(int));
		ellToHell(hellV, hellI, hoff, hack, ellV, ellI, pitch, pitch, rs, R, SPGPU_TYPE_DOUBLE);
		double* dHV = to_device(hellV, sizeof(double) * hellHeight * hack);
		int* dHI = to_device(hellI, sizeof(int) * hellHeight * hack);
		int* dHo = to_device(hoff, sizeof(int) * hacks);
		CHECK(cudaMemset(dz, 0xff, sizeof(double) * R));
		spgpuDhellspmv(h, dz, NULL, 1.0, dHV, dHI, hack, dHo, dRs, NULL, avg, R, dx, 0.0, 0);
		CHECK(cudaEventRecord(t0, stream));
		for (int k = 0; k < reps; ++k)
			spgpuDhellspmv(h, dz, NULL, 1.0, dHV, dHI, hack, dHo, dRs, NULL, avg, R, dx, 0.0, 0);
		CHECK(cudaEventRecord(t1, stream));
		CHECK(cudaEventSynchronize(t1));
		CHECK(cudaEventElapsedTime(&ms, t0, t1));
		CHECK(cudaMemcpy(z, dz, sizeof(double) * R, cudaMemcpyDeviceToHost));
		printf("HELL  %9.3f us / SpMV  %8.1f GFLOP/s   (%.1f MB stored, ELL %.1f MB)\n", 1e3 * ms / reps,
			2.0 * nnz / (1e6 * ms / reps), 12.0 * hellHeight * hack / 1e6, 12.0 * pitch * maxLen / 1e6);
		worst = fmax(worst, check("HELL", z, want, scale, R));
		cudaFree(dV); cudaFree(dI); cudaFree(dRs); cudaFree(dHV); cudaFree(dHI); cudaFree(dHo);
		free(hellV); free(hellI); free(hoff);
	}

	/* ---- HDIA (reference hdia_conv.h) ------------------------------------------------------ */
	{
		const int hack = 32;
		const int hacks = getHdiaHacksCount(hack, R);
		int* hoff = calloc(hacks + 1, sizeof(int));
		int height = 0;
		computeHdiaHackOffsetsFromCoo(&height, hoff, hack, R, C, nnz, rows, cols, 0);
		double* hv = calloc((size_t)height * hack + 1, sizeof(double));
		int* off = calloc(height + 1, sizeof(int));
		cooToHdia(hv, off, hoff, hack, R, C, nnz, rows, cols, vals, 0, SPGPU_TYPE_DOUBLE);
		double* dHv = to_device(hv, sizeof(double) * height * hack);
		int* dOff = to_device(off, sizeof(int) * height);
		int* dHo = to_device(hoff, sizeof(int) * (hacks + 1));
		CHECK(cudaMemset(dz, 0xff, sizeof(double) * R));
		spgpuDhdiaspmv(h, dz, NULL, 1.0, dHv, dOff, hack, dHo, R, C, dx, 0.0);
		CHECK(cudaEventRecord(t0, stream));
		for (int k = 0; k < reps; ++k)
			spgpuDhdiaspmv(h, dz, NULL, 1.0, dHv, dOff, hack, dHo, R, C, dx, 0.0);
		CHECK(cudaEventRecord(t1, stream));
		CHECK(cudaEventSynchronize(t1));
		CHECK(cudaEventElapsedTime(&ms, t0, t1));
		CHECK(cudaMemcpy(z, dz, sizeof(double) * R, cudaMemcpyDeviceToHost));
		printf("HDIA  %9.3f us / SpMV  %8.1f GFLOP/s   (%d hack-diagonals, %.1f MB stored)\n", 1e3 * ms / reps,
			2.0 * nnz / (1e6 * ms / reps), height, 8.0 * height * hack / 1e6);
		worst = fmax(worst, check("HDIA", z, want, scale, R));
		cudaFree(dHv); cudaFree(dOff); cudaFree(dHo);
		free(hv); free(off); free(hoff);
	}

	/* the dot of the result with itself, as the reference drivers print it */
	CHECK(cudaMemcpy(dz, want, sizeof(double) * R, cudaMemcpyHostToDevice));
	printf("dot(z, z) on the device: %.15e\n", spgpuDdot(h, R, dz, dz));

	spgpuDestroy(h);
	return worst <= 1e-12 ? 0 : 1;
}
