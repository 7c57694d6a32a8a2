/*
 * ELL -> HELL ("hacked ELL"), host side.
 *
 * Output contract (bit-exact with reference src/core/hell.c):
 *   computeHellAllocSize  hell.c:4-44    sum over hacks of the longest row of
 *                                        the hack (the tail hack included)
 *   ellToHell             hell.c:46-104  hackOffsets[h] = hackSize * (sum of the
 *                                        longest rows of hacks < h), one entry
 *                                        per hack and NO terminator; slot k of
 *                                        row r of hack h goes to
 *                                        hackOffsets[h] + k*hackSize + r%hackSize.
 *                                        Slots beyond a row's length are left
 *                                        untouched (callers must not read them).
 */
#include <string.h>

#include "spgpu.h"

static int longest_row(const int* lengths, int first, int end)
{
	int longest = 0;
	for (int r = first; r < end; ++r)
		if (lengths[r] > longest)
			longest = lengths[r];
	return longest;
}

void computeHellAllocSize(int* allocationHeight, int hackSize, int rowsCount,
	const int* ellRowLengths)
{
	int height = 0;
	for (int first = 0; first < rowsCount; first += hackSize) {
		int end = first + hackSize < rowsCount ? first + hackSize : rowsCount;
		height += longest_row(ellRowLengths, first, end);
	}
	*allocationHeight = height;
}

void ellToHell(void* hellValues, int* hellIndices, int* hackOffsets,
	int hackSize, const void* ellValues, const int* ellIndices,
	int ellValuesPitch, int ellIndicesPitch, int* ellRowLengths, int rowsCount,
	spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	const char* src = (const char*)ellValues;
	char* dst = (char*)hellValues;
	int base = 0;   /* element offset of the current hack */
	int hack = 0;

	for (int first = 0; first < rowsCount; first += hackSize, ++hack) {
		const int end = first + hackSize < rowsCount ? first + hackSize : rowsCount;
		hackOffsets[hack] = base;
		for (int row = first; row < end; ++row) {
			const int lane = row - first;
			const int len = ellRowLengths[row];
			for (int k = 0; k < len; ++k) {
				const size_t to = (size_t)base + (size_t)k * (size_t)hackSize + (size_t)lane;
				hellIndices[to] = ellIndices[(size_t)k * (size_t)ellIndicesPitch + (size_t)row];
				memcpy(dst + to * bytes,
					src + ((size_t)k * (size_t)ellValuesPitch + (size_t)row) * bytes, bytes);
			}
		}
		base += hackSize * longest_row(ellRowLengths, first, end);
	}
}
