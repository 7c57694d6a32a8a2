/*
 * Block-COO construction (host), bit-exact port of reference src/core/coo.cpp:8-98
 * (computeBcooSize, cooToBcoo; declared in coo_conv.h:20-24).  No kernel of the
 * reference consumes block formats; the functions are exported by its library, so they are
 * kept for link compatibility.
 *
 * The reference numbers the non-zero blocks in the order in which their first entry appears
 * (a std::map from block id to position); here the same first-seen order comes from an
 * open-addressing hash table keyed by the same 64-bit id (blockRow | blockCol << 32).
 */
#include <stdlib.h>
#include <string.h>

#include "spgpu.h"

typedef struct BlockTable {
	unsigned long long* key;
	int* pos;
	size_t mask;
} BlockTable;

static int table_init(BlockTable* t, int nonZeros)
{
	size_t cap = 16;
	while (cap < 2 * (size_t)(nonZeros > 0 ? nonZeros : 1))
		cap <<= 1;
	t->key = (unsigned long long*)malloc(cap * sizeof(unsigned long long));
	t->pos = (int*)malloc(cap * sizeof(int));
	t->mask = cap - 1;
	if (!t->key || !t->pos)
		return -1;
	memset(t->pos, 0xff, cap * sizeof(int));       /* -1 = empty */
	return 0;
}

static void table_free(BlockTable* t)
{
	free(t->key);
	free(t->pos);
}

/* position of block `id`, inserting it with position *count (then incremented) when new */
static int table_find_or_add(BlockTable* t, unsigned long long id, int* count, int* isNew)
{
	size_t h = (size_t)((id * 0x9E3779B97F4A7C15ull) >> 17) & t->mask;
	while (t->pos[h] >= 0) {
		if (t->key[h] == id) {
			*isNew = 0;
			return t->pos[h];
		}
		h = (h + 1) & t->mask;
	}
	t->key[h] = id;
	t->pos[h] = *count;
	*isNew = 1;
	return (*count)++;
}

static unsigned long long block_id(int row, int col, int blockRows, int blockCols)
{
	const unsigned long long br = (unsigned long long)(row / blockRows);
	const unsigned long long bc = (unsigned long long)(col / blockCols);
	return br | (bc << 32);
}

int computeBcooSize(int blockRows, int blockCols, const int* rows, const int* cols, int nonZeros)
{
	BlockTable t;
	int count = 0, isNew;
	if (table_init(&t, nonZeros) != 0)
		return 0;
	for (int i = 0; i < nonZeros; ++i)
		table_find_or_add(&t, block_id(rows[i], cols[i], blockRows, blockCols), &count, &isNew);
	table_free(&t);
	return count;
}

/* blocks are stored column-major (coo.cpp:84-88), zero-filled when first seen */
void cooToBcoo(int* bRows, int* bCols, void* blockValues, int blockRows, int blockCols,
	const int* rows, const int* cols, const void* values, int nonZeros, spgpuType_t valuesType)
{
	const size_t bytes = spgpuSizeOf(valuesType);
	const size_t blockBytes = bytes * (size_t)blockRows * (size_t)blockCols;
	BlockTable t;
	int count = 0, isNew;
	if (table_init(&t, nonZeros) != 0)
		return;
	for (int i = 0; i < nonZeros; ++i) {
		const int pos = table_find_or_add(&t, block_id(rows[i], cols[i], blockRows, blockCols), &count, &isNew);
		if (isNew) {
			bRows[pos] = rows[i] / blockRows;
			bCols[pos] = cols[i] / blockCols;
			memset((char*)blockValues + (size_t)pos * blockBytes, 0, blockBytes);
		}
		const size_t inBlock = (size_t)(rows[i] % blockRows) + (size_t)(cols[i] % blockCols) * (size_t)blockRows;
		memcpy((char*)blockValues + (size_t)pos * blockBytes + inBlock * bytes, (const char*)values + (size_t)i * bytes, bytes);
	}
	table_free(&t);
}
