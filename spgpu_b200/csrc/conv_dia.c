/*
 * COO -> DIA, host side.
 *
 * Output contract (bit-exact with reference src/core/dia.c):
 *   computeDiaAllocPitch      dia.c:5-9     rows rounded up to 32 elements
 *   computeDiaDiagonalsCount  dia.c:11-38   number of distinct (col - row);
 *                                           raw indices are used, the base
 *                                           cancels in the difference
 *   coo2dia                   dia.c:40-104  offsets[] = the distinct (col - row)
 *                                           in ascending order; A(r, c) is
 *                                           stored at values[(r - cooBase) +
 *                                           pos*pitch]; cells without an entry
 *                                           are left as the caller filled them.
 */
#include <stdlib.h>
#include <string.h>

#include "spgpu.h"

int computeDiaAllocPitch(int rowsCount)
{
	return (rowsCount + 31) & ~31;
}

/* Marks which of the rows+cols-1 possible diagonals are populated; returns how many. */
static int mark_diagonals(unsigned char* seen, int rowsCount, int span,
	int nonZerosCount, const int* rowIdx, const int* colIdx)
{
	int distinct = 0;
	memset(seen, 0, (size_t)span);
	for (int e = 0; e < nonZerosCount; ++e) {
		int slot = rowsCount - 1 + colIdx[e] - rowIdx[e];
		if (!seen[slot]) {
			seen[slot] = 1;
			++distinct;
		}
	}
	return distinct;
}

int computeDiaDiagonalsCount(int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices)
{
	const int span = rowsCount + columnsCount - 1;
	unsigned char* seen = (unsigned char*)malloc((size_t)(span > 0 ? span : 1));
	int n = mark_diagonals(seen, rowsCount, span, nonZerosCount, cooRowIndices, cooColsIndices);
	free(seen);
	return n;
}

void coo2dia(void* values, int* offsets, int valuesPitch, int diagonals,
	int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, const void* cooValues,
	int cooBaseIndex, spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	const int span = rowsCount + columnsCount - 1;
	unsigned char* seen = (unsigned char*)malloc((size_t)(span > 0 ? span : 1));
	int* position = (int*)malloc((size_t)(span > 0 ? span : 1) * sizeof(int));
	int next = 0;
	(void)diagonals;

	mark_diagonals(seen, rowsCount, span, nonZerosCount, cooRowIndices, cooColsIndices);
	for (int slot = 0; slot < span; ++slot) {
		if (seen[slot]) {
			position[slot] = next;
			offsets[next++] = slot - (rowsCount - 1);
		}
	}
	for (int e = 0; e < nonZerosCount; ++e) {
		const int row = cooRowIndices[e];
		const int pos = position[rowsCount - 1 + cooColsIndices[e] - row];
		memcpy((char*)values + ((size_t)(row - cooBaseIndex) + (size_t)pos * (size_t)valuesPitch) * bytes,
			(const char*)cooValues + (size_t)e * bytes, bytes);
	}
	free(position);
	free(seen);
}
