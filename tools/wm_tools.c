/* Data-generation helpers for bench.py / tools (NOT on the product path): canonical k-mer counting for the
 * meryl stand-in of tools/gen_data.py (ext/meryl does not build offline).  k <= 16: direct-index
 * table of 4^k uint32 counters.  Built by __graft_entry__.build() into tools/libwm_tools.so. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* adds the canonical k-mers of seq[0..n) (ASCII, anything but ACGTacgt breaks the k-mer) to table[4^k] */
void wm_tools_count(const uint8_t *seq, int64_t n, int k, uint32_t *table)
{
	static int8_t lut[256]; static int init = 0;
	if (!init) { memset(lut, 4, 256); lut['A'] = lut['a'] = 0; lut['C'] = lut['c'] = 1; lut['G'] = lut['g'] = 2; lut['T'] = lut['t'] = 3; init = 1; }
	const uint64_t mask = (1ULL << (2 * k)) - 1, shift = 2 * (k - 1);
	#pragma omp parallel
	{
		#pragma omp for schedule(static)
		for (int64_t blk = 0; blk < (n + (1 << 20) - 1) >> 20; ++blk) {
			const int64_t b = blk << 20;
			int64_t e = b + (1 << 20) + k - 1;
			if (e > n) e = n;
			uint64_t fw = 0, rv = 0; int l = 0;
			for (int64_t i = b; i < e; ++i) {
				const int c = lut[seq[i]];
				if (c < 4) {
					fw = (fw << 2 | (uint64_t)c) & mask;
					rv = rv >> 2 | (uint64_t)(3 - c) << shift;
					if (++l >= k) __atomic_fetch_add(&table[fw < rv ? fw : rv], 1u, __ATOMIC_RELAXED);
				} else l = 0;
			}
		}
	}
}

/* histogram of the non-zero counters: hist[min(c, hist_n - 1)] += 1; returns the number of distinct k-mers */
int64_t wm_tools_hist(const uint32_t *table, int k, int64_t *hist, int hist_n)
{
	const int64_t n = 1LL << (2 * k);
	int64_t distinct = 0;
	#pragma omp parallel
	{
		int64_t *h = (int64_t*)calloc(hist_n, sizeof(int64_t)), d = 0;
		#pragma omp for schedule(static) nowait
		for (int64_t i = 0; i < n; ++i) {
			const uint32_t c = table[i];
			if (c) { ++d; ++h[c < (uint32_t)hist_n ? c : (uint32_t)hist_n - 1]; }
		}
		#pragma omp critical
		{ for (int j = 0; j < hist_n; ++j) hist[j] += h[j]; distinct += d; }
		free(h);
	}
	return distinct;
}

/* k-mers with count > thr, ascending: writes up to cap (kmer, count) pairs, returns how many there are */
int64_t wm_tools_above(const uint32_t *table, int k, uint32_t thr, uint64_t *kmers, uint32_t *counts, int64_t cap)
{
	const int64_t n = 1LL << (2 * k);
	int64_t m = 0;
	for (int64_t i = 0; i < n; ++i)
		if (table[i] > thr) { if (m < cap) kmers[m] = (uint64_t)i, counts[m] = table[i]; ++m; }
	return m;
}
