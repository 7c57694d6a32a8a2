/*
 * COO -> ELL and ELL -> ordered ELL, host side.
 *
 * Output contract (bit-exact with reference src/core/ell.c):
 *   computeEllRowLenghts  ell.c:5-31    per-row counts and their maximum
 *   computeEllAllocPitch  ell.c:33-37   rows rounded up to 32 elements
 *   cooToEll              ell.c:39-80   slot order inside a row = order of
 *                                       appearance in the COO arrays; stored
 *                                       index = col - cooBase + ellBase
 *   ellToOell             ell.c:161-202 rows by descending length; the
 *                                       reference's merge (ell.c:85-157) takes
 *                                       the right run on ties, which makes the
 *                                       result "length descending, then row
 *                                       index descending" -- produced here by a
 *                                       counting sort over the lengths.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "spgpu.h"

/* Copy one element of `bytes` bytes; the common sizes avoid the memcpy call. */
static inline void copy_element(void* dst, const void* src, size_t bytes)
{
	switch (bytes) {
	case 4:  memcpy(dst, src, 4); break;
	case 8:  memcpy(dst, src, 8); break;
	case 16: memcpy(dst, src, 16); break;
	default: memcpy(dst, src, bytes); break;
	}
}

void computeEllRowLenghts(int* ellRowLengths, int* ellMaxRowSize, int rowsCount,
	int nonZerosCount, const int* cooRowIndices, int cooBaseIndex)
{
	int longest = 0;
	memset(ellRowLengths, 0, (size_t)rowsCount * sizeof(int));
	for (int e = 0; e < nonZerosCount; ++e) {
		int n = ++ellRowLengths[cooRowIndices[e] - cooBaseIndex];
		if (n > longest)
			longest = n;
	}
	*ellMaxRowSize = longest;
}

int computeEllAllocPitch(int rowsCount)
{
	return (rowsCount + 31) & ~31;
}

void cooToEll(void* ellValues, int* ellIndices, int ellValuesPitch,
	int ellIndicesPitch, int ellMaxRowSize, int ellBaseIndex, int rowsCount,
	int nonZerosCount, const int* cooRowIndices, const int* cooColsIndices,
	const void* cooValues, int cooBaseIndex, spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	const int shift = ellBaseIndex - cooBaseIndex;
	int* fill = (int*)calloc((size_t)(rowsCount > 0 ? rowsCount : 1), sizeof(int));
	char* vals = (char*)ellValues;
	const char* src = (const char*)cooValues;
	(void)ellMaxRowSize;

	for (int e = 0; e < nonZerosCount; ++e) {
		const int row = cooRowIndices[e] - cooBaseIndex;
		const size_t slot = (size_t)fill[row]++;
		ellIndices[(size_t)row + slot * (size_t)ellIndicesPitch] = cooColsIndices[e] + shift;
		copy_element(vals + ((size_t)row + slot * (size_t)ellValuesPitch) * bytes,
			src + (size_t)e * bytes, bytes);
	}
	free(fill);
}

void ellToOell(int* rIdx, void* dstEllValues, int* dstEllIndices, int* dstRs,
	const void* srcEllValues, const int* srcEllIndices, const int* srcRs,
	int ellValuesPitch, int ellIndicesPitch, int rowsCount,
	spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	int longest = 0;
	int* start;

	if (rowsCount <= 0)
		return;
	for (int r = 0; r < rowsCount; ++r)
		if (srcRs[r] > longest)
			longest = srcRs[r];

	/* counting sort: bucket L starts after all longer rows; inside a bucket the
	 * rows are laid down from the highest source index to the lowest */
	start = (int*)calloc((size_t)longest + 2, sizeof(int));
	for (int r = 0; r < rowsCount; ++r)
		++start[srcRs[r]];
	{
		int run = 0;
		for (int len = longest; len >= 0; --len) {
			int n = start[len];
			start[len] = run;
			run += n;
		}
	}
	for (int r = rowsCount - 1; r >= 0; --r) {
		int pos = start[srcRs[r]]++;
		rIdx[pos] = r;
		dstRs[pos] = srcRs[r];
	}
	free(start);

	for (int i = 0; i < rowsCount; ++i) {
		const int from = rIdx[i];
		const int len = srcRs[from];
		for (int k = 0; k < len; ++k) {
			const size_t vOff = (size_t)k * (size_t)ellValuesPitch;
			const size_t iOff = (size_t)k * (size_t)ellIndicesPitch;
			copy_element((char*)dstEllValues + ((size_t)i + vOff) * bytes,
				(const char*)srcEllValues + ((size_t)from + vOff) * bytes, bytes);
			dstEllIndices[(size_t)i + iOff] = srcEllIndices[(size_t)from + iOff];
		}
	}
}
