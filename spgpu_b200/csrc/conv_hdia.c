/*
 * DIA -> HDIA and COO -> HDIA ("hacked DIA"), host side, plain C.
 *
 * Output contract (bit-exact with reference src/core/hdia.cpp):
 *   getHdiaHacksCount               hdia.cpp:8-11
 *   computeHdiaHackOffsets          hdia.cpp:13-61   a DIA diagonal belongs to a
 *       hack iff any BYTE of its cells inside the hack is non-zero (so -0.0
 *       counts); hackOffsets is the prefix sum of kept diagonals, hacks+1 entries
 *   diaToHdia                       hdia.cpp:68-153  kept diagonals stay in DIA
 *       order; cells are copied as they are, zeros included
 *   computeHdiaHackOffsetsFromCoo   hdia.cpp:161-228 per hack, count the distinct
 *       values of (col - base) - ((row - base) % hackSize)
 *   cooToHdia                       hdia.cpp:230-349 per hack, diagonals ordered
 *       by that key ascending; stored offset is the global col - row; cell
 *       (h, d, r) lives at ((row-base) % hackSize) + hackSize*(hackOffsets[h]+d)
 *
 * The reference uses std::vector / std::map per hack; here the non-zeros are
 * bucketed by hack with one counting pass and each hack's keys are sorted in a
 * reusable scratch array.
 */
#include <stdlib.h>
#include <string.h>

#include "spgpu.h"

int getHdiaHacksCount(int hackSize, int rowsCount)
{
	return (rowsCount + hackSize - 1) / hackSize;
}

/* 1 when any byte of the `n` elements of `bytes` bytes starting at p is non-zero */
static int any_byte_set(const unsigned char* p, size_t n)
{
	for (size_t i = 0; i < n; ++i)
		if (p[i])
			return 1;
	return 0;
}

void computeHdiaHackOffsets(int* allocationHeight, int* hackOffsets,
	int hackSize, const void* diaValues, int diaValuesPitch, int diagonals,
	int rowsCount, spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	const int hacks = getHdiaHacksCount(hackSize, rowsCount);
	const unsigned char* dia = (const unsigned char*)diaValues;
	int kept = 0;

	hackOffsets[0] = 0;
	for (int h = 0; h < hacks; ++h) {
		const int first = h * hackSize;
		const int n = first + hackSize <= rowsCount ? hackSize : rowsCount - first;
		for (int d = 0; d < diagonals; ++d)
			kept += any_byte_set(dia + ((size_t)first + (size_t)d * (size_t)diaValuesPitch) * bytes,
				(size_t)n * bytes);
		hackOffsets[h + 1] = kept;
	}
	*allocationHeight = hackOffsets[hacks];
}

void diaToHdia(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, const void* diaValues, const int* diaOffsets,
	int diaValuesPitch, int diagonals, int rowsCount, spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	const int hacks = getHdiaHacksCount(hackSize, rowsCount);
	const unsigned char* dia = (const unsigned char*)diaValues;
	unsigned char* out = (unsigned char*)hdiaValues;

	for (int h = 0; h < hacks; ++h) {
		const int first = h * hackSize;
		const int n = first + hackSize <= rowsCount ? hackSize : rowsCount - first;
		int slot = hackOffsets[h];
		for (int d = 0; d < diagonals; ++d) {
			const unsigned char* cells = dia + ((size_t)first + (size_t)d * (size_t)diaValuesPitch) * bytes;
			if (!any_byte_set(cells, (size_t)n * bytes))
				continue;
			hdiaOffsets[slot] = diaOffsets[d];
			memcpy(out + (size_t)slot * (size_t)hackSize * bytes, cells, (size_t)n * bytes);
			++slot;
		}
	}
}

/* ---- COO path ------------------------------------------------------------ */

typedef struct HackBuckets {
	int* begin;    /* hacks+1 entries: nnz range of each hack inside `entry` */
	int* entry;    /* COO positions grouped by hack, COO order kept inside   */
} HackBuckets;

static void bucket_by_hack(HackBuckets* b, int hacks, int hackSize, int nnz,
	const int* rowIdx, int base)
{
	b->begin = (int*)calloc((size_t)hacks + 2, sizeof(int));
	b->entry = (int*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int));
	for (int e = 0; e < nnz; ++e)
		++b->begin[(rowIdx[e] - base) / hackSize + 1];
	for (int h = 0; h < hacks; ++h)
		b->begin[h + 1] += b->begin[h];
	{
		int* cursor = (int*)malloc((size_t)(hacks > 0 ? hacks : 1) * sizeof(int));
		memcpy(cursor, b->begin, (size_t)hacks * sizeof(int));
		for (int e = 0; e < nnz; ++e)
			b->entry[cursor[(rowIdx[e] - base) / hackSize]++] = e;
		free(cursor);
	}
}

static void free_buckets(HackBuckets* b)
{
	free(b->begin);
	free(b->entry);
}

static int cmp_int(const void* a, const void* b)
{
	int x = *(const int*)a, y = *(const int*)b;
	return (x > y) - (x < y);
}

/* hack-local diagonal key of a COO entry (column minus row-inside-hack) */
static inline int local_key(int row, int col, int base, int hackSize)
{
	return (col - base) - ((row - base) % hackSize);
}

/* Fills keys[] with the sorted distinct keys of hack h; returns how many. */
static int distinct_keys(int* keys, const HackBuckets* b, int h, int hackSize,
	const int* rowIdx, const int* colIdx, int base)
{
	const int lo = b->begin[h], hi = b->begin[h + 1];
	int n = 0, out = 0;
	for (int p = lo; p < hi; ++p) {
		int e = b->entry[p];
		keys[n++] = local_key(rowIdx[e], colIdx[e], base, hackSize);
	}
	qsort(keys, (size_t)n, sizeof(int), cmp_int);
	for (int i = 0; i < n; ++i)
		if (out == 0 || keys[i] != keys[out - 1])
			keys[out++] = keys[i];
	return out;
}

static int largest_bucket(const HackBuckets* b, int hacks)
{
	int most = 1;
	for (int h = 0; h < hacks; ++h)
		if (b->begin[h + 1] - b->begin[h] > most)
			most = b->begin[h + 1] - b->begin[h];
	return most;
}

void computeHdiaHackOffsetsFromCoo(int* allocationHeight, int* hackOffsets,
	int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, int cooBaseIndex)
{
	const int hacks = getHdiaHacksCount(hackSize, rowsCount);
	HackBuckets b;
	int* keys;
	(void)columnsCount;

	bucket_by_hack(&b, hacks, hackSize, nonZerosCount, cooRowIndices, cooBaseIndex);
	keys = (int*)malloc((size_t)largest_bucket(&b, hacks) * sizeof(int));
	hackOffsets[0] = 0;
	for (int h = 0; h < hacks; ++h)
		hackOffsets[h + 1] = hackOffsets[h] +
			distinct_keys(keys, &b, h, hackSize, cooRowIndices, cooColsIndices, cooBaseIndex);
	*allocationHeight = hackOffsets[hacks];
	free(keys);
	free_buckets(&b);
}

/* shared by cooToHdia and bcooToBhdia (reference hdia.cpp:230-325, cooToHdia_size): `bytes` is the
 * size of one stored cell -- one element, or one block of elements */
static void coo_to_hdia_cells(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, int rowsCount, int nonZerosCount, const int* cooRowIndices,
	const int* cooColsIndices, const void* cooValues, int cooBaseIndex, size_t bytes)
{
	const int hacks = getHdiaHacksCount(hackSize, rowsCount);
	HackBuckets b;
	int* keys;
	int written = 0;   /* diagonals emitted so far (the reference advances its
	                      offsets cursor by the count it finds, hdia.cpp:303) */

	bucket_by_hack(&b, hacks, hackSize, nonZerosCount, cooRowIndices, cooBaseIndex);
	keys = (int*)malloc((size_t)largest_bucket(&b, hacks) * sizeof(int));

	for (int h = 0; h < hacks; ++h) {
		const int nkeys = distinct_keys(keys, &b, h, hackSize, cooRowIndices, cooColsIndices, cooBaseIndex);
		/* every entry of a hack-local diagonal shares one global col - row:
		 * key - h*hackSize */
		for (int d = 0; d < nkeys; ++d)
			hdiaOffsets[written + d] = keys[d] - h * hackSize;
		written += nkeys;

		for (int p = b.begin[h]; p < b.begin[h + 1]; ++p) {
			const int e = b.entry[p];
			const int row = cooRowIndices[e];
			const int key = local_key(row, cooColsIndices[e], cooBaseIndex, hackSize);
			const int* hit = (const int*)bsearch(&key, keys, (size_t)nkeys, sizeof(int), cmp_int);
			const size_t d = (size_t)(hit - keys);
			const size_t cell = (size_t)((row - cooBaseIndex) % hackSize) +
				(size_t)hackSize * ((size_t)hackOffsets[h] + d);
			memcpy((char*)hdiaValues + cell * bytes, (const char*)cooValues + (size_t)e * bytes, bytes);
		}
	}
	free(keys);
	free_buckets(&b);
}

void cooToHdia(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, const void* cooValues,
	int cooBaseIndex, spgpuType_t valuesType)
{
	(void)columnsCount;
	coo_to_hdia_cells(hdiaValues, hdiaOffsets, hackOffsets, hackSize, rowsCount, nonZerosCount,
		cooRowIndices, cooColsIndices, cooValues, cooBaseIndex, spgpuSizeOf(valuesType));
}

/* reference hdia.cpp:351-373: cooToHdia over BLOCKS -- the COO entries are block coordinates
 * (cooToBcoo's bRows / bCols) and every value is a block of blockSize elements. */
void bcooToBhdia(void* hdiaValues, int* hdiaOffsets, const int* hackOffsets,
	int hackSize, int rowsCount, int columnsCount, int nonZerosCount,
	const int* cooRowIndices, const int* cooColsIndices, const void* cooValues,
	int cooBaseIndex, spgpuType_t valuesType, int blockSize)
{
	(void)columnsCount;
	coo_to_hdia_cells(hdiaValues, hdiaOffsets, hackOffsets, hackSize, rowsCount, nonZerosCount,
		cooRowIndices, cooColsIndices, cooValues, cooBaseIndex, (size_t)blockSize * spgpuSizeOf(valuesType));
}
