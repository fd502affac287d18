// Host side of index construction (mm_idx_gen / worker_post, reference src/index.c:196-449): the (minimizer, position)
// pairs the sketch kernel produced are ordered by minimizer hash, then position -- the order in which mm_idx_get hands
// out occurrence lists (src/index.c:239) -- and cut into the CSR the device index is built from.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>
#include <omp.h>

namespace wmh {

// mm_seq4_set for a whole sequence (src/mmpriv.h:29, src/index.c:321-332): bases [o0, o0 + L) of the concatenated reference, 4 bits
// each, eight per 32-bit word; A C G T (either case) -> 0..3, everything else 4.  S must be zero where it is written.  Whole
// words in parallel, the (at most two) words shared with a neighbouring sequence serially.
static inline void pack_seq4(uint32_t *S, uint64_t o0, const char *seq, uint64_t L)
{
	static const uint8_t code4[256] = {
#define WM_W4 4, 4, 4, 4
#define WM_W16 WM_W4, WM_W4, WM_W4, WM_W4
		WM_W16, WM_W16, WM_W16, WM_W16,
		4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
		4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
		WM_W16, WM_W16, WM_W16, WM_W16, WM_W16, WM_W16, WM_W16, WM_W16
#undef WM_W16
#undef WM_W4
	};
	const uint64_t o1 = o0 + L, w0 = (o0 + 7) / 8, w1 = o1 / 8; // words [w0, w1) lie inside this sequence
	const unsigned char *sq = (const unsigned char*)seq;
	for (uint64_t o = o0; o < o1 && o < w0 * 8; ++o) S[o >> 3] |= (uint32_t)code4[sq[o - o0]] << ((o & 7) << 2);
	#pragma omp parallel for schedule(static) if (w1 > w0 + 4096)
	for (int64_t wi = (int64_t)w0; wi < (int64_t)w1; ++wi) {
		const unsigned char *q = sq + ((uint64_t)wi * 8 - o0);
		S[wi] = (uint32_t)code4[q[0]] | (uint32_t)code4[q[1]] << 4 | (uint32_t)code4[q[2]] << 8 | (uint32_t)code4[q[3]] << 12 |
		        (uint32_t)code4[q[4]] << 16 | (uint32_t)code4[q[5]] << 20 | (uint32_t)code4[q[6]] << 24 | (uint32_t)code4[q[7]] << 28;
	}
	for (uint64_t o = w1 >= w0 ? w1 * 8 : o1; o < o1; ++o) S[o >> 3] |= (uint32_t)code4[sq[o - o0]] << ((o & 7) << 2);
}

// Sort by (x >> 8, y).  Keys are partitioned by their leading bits, the partitions are sorted concurrently; the result
// is the unique sorted order (no two entries share hash and position), so it does not depend on the thread count.
template <typename T>
void sort_index_pairs(std::vector<T> &a, int n_threads)
{
	auto less = [](const T &p, const T &q) { return (p.x >> 8) != (q.x >> 8) ? (p.x >> 8) < (q.x >> 8) : p.y < q.y; };
	const size_t n = a.size();
	if (n_threads < 2 || n < (size_t)1 << 16) { std::sort(a.begin(), a.end(), less); return; }
	uint64_t max_key = 0;
	#pragma omp parallel for reduction(max : max_key) num_threads(n_threads)
	for (int64_t i = 0; i < (int64_t)n; ++i) max_key = std::max<uint64_t>(max_key, a[i].x >> 8);
	int bits = 0;
	while (bits < 64 && (max_key >> bits) != 0) ++bits;
	const int pbits = 10, shift = bits > pbits ? bits - pbits : 0, P = 1 << pbits;
	std::vector<size_t> cnt((size_t)n_threads * P, 0), start((size_t)n_threads * P + 1, 0);
	#pragma omp parallel num_threads(n_threads)
	{
		const int t = omp_get_thread_num(), nt = omp_get_num_threads();
		const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
		size_t *c = cnt.data() + (size_t)t * P;
		for (size_t i = lo; i < hi; ++i) ++c[(a[i].x >> 8) >> shift];
	}
	// offsets: partition-major, then thread (the input order inside a partition does not matter, it is sorted next)
	std::vector<size_t> part_off(P + 1, 0);
	{
		size_t acc = 0;
		for (int p = 0; p < P; ++p) {
			part_off[p] = acc;
			for (int t = 0; t < n_threads; ++t) { start[(size_t)t * P + p] = acc; acc += cnt[(size_t)t * P + p]; }
		}
		part_off[P] = acc;
	}
	std::vector<T> b(n);
	#pragma omp parallel num_threads(n_threads)
	{
		const int t = omp_get_thread_num(), nt = omp_get_num_threads();
		const size_t lo = n * t / nt, hi = n * (t + 1) / nt;
		size_t *s = start.data() + (size_t)t * P;
		for (size_t i = lo; i < hi; ++i) b[s[(a[i].x >> 8) >> shift]++] = a[i];
	}
	#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
	for (int p = 0; p < P; ++p) std::sort(b.begin() + part_off[p], b.begin() + part_off[p + 1], less);
	a.swap(b);
}

} // namespace wmh
