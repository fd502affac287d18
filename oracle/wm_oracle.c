/*
 * oracle/wm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference algorithms on the seed-chain-align hot path
 * of marbl/Winnowmap v2.03.  Each function cites the reference
 * file:line it follows (paths relative to /root/reference).  This file is the checker
 * for the CUDA kernels: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.  The product (winnowmap_b200/) never links, imports or executes it.
 *
 * Parity pinning: tests/test_oracle_vs_ref.py checks every function here against the
 * REAL reference functions (oracle/_ref/libref_harness.so, built from /root/reference by
 * oracle/build_ref.sh) on seeded random inputs, and tests/golden/ holds vectors generated
 * from the reference by tools/make_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------
 * Bloom filter of down-weighted k-mers.
 * ext/bloom/bloom_filter.hpp:108-147 (table sizing), :183-195 (ctor, seed), :461-465
 * (index), :513-529 (salts), :551-565 (hash_ap on an 8-byte key); parameters from
 * src/index.c:411-414 (n = max(count,1000), p = 0.001, at most 2 hashes).
 * ---------------------------------------------------------------------------------- */
typedef struct {
	uint64_t table_bits;   /* table_size_ (multiple of 8) */
	uint32_t salt[2];
	int n_salt;
	uint8_t *table;
} wmo_bloom_t;

static uint32_t wmo_hash_ap8(uint64_t key, uint32_t hash)
{ /* bloom_filter.hpp:556-565, one 8-byte round */
	uint32_t i1 = (uint32_t)key, i2 = (uint32_t)(key >> 32);
	hash ^= (hash << 7) ^ (i1 * (hash >> 3)) ^ (~((hash << 11) + (i2 ^ (hash >> 5))));
	return hash;
}

wmo_bloom_t *wmo_bloom_init(uint64_t n_kmers)
{
	wmo_bloom_t *b = (wmo_bloom_t*)calloc(1, sizeof(wmo_bloom_t));
	double n = (double)(n_kmers > 1000 ? n_kmers : 1000), p = 0.001;
	double min_m = INFINITY, min_k = 0.0, k = 1.0;
	while (k < 1000.0) { /* bloom_filter.hpp:121-136 */
		double num = -k * n, den = log(1.0 - pow(p, 1.0 / k)), m = num / den;
		if (m < min_m) min_m = m, min_k = k;
		k += 1.0;
	}
	unsigned nh = (unsigned)min_k;
	uint64_t ts = (uint64_t)min_m;
	ts += (ts % 8) != 0 ? 8 - ts % 8 : 0;
	if (nh < 1) nh = 1; else if (nh > 2) nh = 2; /* maximum_number_of_hashes = 2 (index.c:414) */
	b->table_bits = ts; b->n_salt = (int)nh;
	{ /* bloom_filter.hpp:186 and :513-529 */
		uint64_t seed = 0xA5A5A5A55A5A5A5AULL * 0xA5A5A5A5ULL + 1;
		uint32_t predef[2] = { 0xAAAAAAAAu, 0x55555555u };
		int i;
		for (i = 0; i < b->n_salt; ++i) b->salt[i] = predef[i];
		for (i = 0; i < b->n_salt; ++i) b->salt[i] = b->salt[i] * b->salt[(i + 3) % b->n_salt] + (uint32_t)seed;
	}
	b->table = (uint8_t*)calloc(ts / 8 + 1, 1);
	return b;
}

void wmo_bloom_insert(wmo_bloom_t *b, uint64_t key)
{ /* bloom_filter.hpp:260-273 */
	int i;
	for (i = 0; i < b->n_salt; ++i) {
		uint64_t bit = wmo_hash_ap8(key, b->salt[i]) % b->table_bits;
		b->table[bit >> 3] |= (uint8_t)(1u << (bit & 7));
	}
}

int wmo_bloom_contains(const wmo_bloom_t *b, uint64_t key)
{ /* bloom_filter.hpp:303-319 */
	int i;
	for (i = 0; i < b->n_salt; ++i) {
		uint64_t bit = wmo_hash_ap8(key, b->salt[i]) % b->table_bits;
		if (!(b->table[bit >> 3] >> (bit & 7) & 1)) return 0;
	}
	return 1;
}

uint64_t wmo_bloom_bits(const wmo_bloom_t *b) { return b->table_bits; }
const uint8_t *wmo_bloom_table(const wmo_bloom_t *b) { return b->table; }
void wmo_bloom_salts(const wmo_bloom_t *b, uint32_t *s) { s[0] = b->salt[0]; s[1] = b->n_salt > 1 ? b->salt[1] : 0; }
int wmo_bloom_nsalt(const wmo_bloom_t *b) { return b->n_salt; }
void wmo_bloom_free(wmo_bloom_t *b) { if (b) { free(b->table); free(b); } }

/* ------------------------------------------------------------------------------------
 * Weighted robust-winnowing minimizers: src/sketch.c:43-219.
 * ---------------------------------------------------------------------------------- */
static const unsigned char wmo_nt4[256] = { /* src/sketch.c:19-36 */
	0, 1, 2, 3,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  3, 3, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  3, 3, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
	4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4
};

uint64_t wmo_murmur64(uint64_t key)
{ /* sketch.c:43-51 (mask = UINT64_MAX) */
	key ^= key >> 33; key *= 0xff51afd7ed558ccdULL;
	key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ULL;
	key ^= key >> 33;
	return key;
}

uint64_t wmo_hash64(uint64_t key, uint64_t mask)
{ /* sketch.c:53-63 */
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

double wmo_weight(uint64_t kmer, const wmo_bloom_t *bf)
{ /* sketch.c:70-90 */
	uint64_t h = wmo_murmur64(kmer);
	double x = (double)h * 1.0 / 18446744073709551616.0; /* (double)UINT64_MAX == 2^64 */
	if (bf && wmo_bloom_contains(bf, kmer)) {
		double p2 = x * x, p4 = p2 * p2;
		return -1.0 * (p4 * p4);
	}
	return -1.0 * x;
}

/* returns the number of minimizers; writes up to max_out (x,y) pairs */
long wmo_sketch(const char *str, int len, int w, int k, uint32_t rid, const wmo_bloom_t *bf, uint64_t *out_xy, long max_out)
{ /* sketch.c:128-219, is_hpc == 0 */
	uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1, kmer[2] = {0, 0};
	int i, j, l, buf_pos, min_pos, kmer_span = 0;
	uint64_t bufx[256], bufy[256], minx = UINT64_MAX, miny = UINT64_MAX;
	double buf_order[256], min_order = 2.0;
	long n = 0;
#define WMO_EMIT() do { if (n < max_out) out_xy[2*n] = minx, out_xy[2*n+1] = miny; ++n; } while (0)
	for (i = 0; i < w; ++i) bufx[i] = bufy[i] = UINT64_MAX, buf_order[i] = 2.0;
	for (i = l = buf_pos = min_pos = 0; i < len; ++i) {
		int c = wmo_nt4[(uint8_t)str[i]];
		uint64_t ix = UINT64_MAX, iy = UINT64_MAX;
		double io = 2.0;
		if (c < 4) {
			int z;
			kmer_span = l + 1 < k ? l + 1 : k;
			kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
			kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
			if (kmer[0] == kmer[1]) continue; /* sketch.c:166 */
			z = kmer[0] < kmer[1] ? 0 : 1;
			++l;
			if (l >= k && kmer_span < 256) {
				ix = wmo_hash64(kmer[z], mask) << 8 | (uint64_t)kmer_span;
				iy = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint64_t)z;
				io = wmo_weight(kmer[z], bf);
			}
		} else l = 0, kmer_span = 0;
		bufx[buf_pos] = ix, bufy[buf_pos] = iy, buf_order[buf_pos] = io;
		if (io < min_order) { /* sketch.c:180-189 */
			if (l >= w + k && minx != UINT64_MAX) WMO_EMIT();
			minx = ix, miny = iy, min_pos = buf_pos, min_order = io;
		} else if (buf_pos == min_pos) { /* sketch.c:191-205 */
			if (l >= w + k - 1 && minx != UINT64_MAX) WMO_EMIT();
			for (j = buf_pos + 1, minx = UINT64_MAX, miny = UINT64_MAX, min_order = 2.0; j < w; ++j)
				if (min_order >= buf_order[j]) minx = bufx[j], miny = bufy[j], min_pos = j, min_order = buf_order[j];
			for (j = 0; j <= buf_pos; ++j)
				if (min_order >= buf_order[j]) minx = bufx[j], miny = bufy[j], min_pos = j, min_order = buf_order[j];
		}
		if (++buf_pos == w) buf_pos = 0;
	}
	if (minx != UINT64_MAX) WMO_EMIT(); /* sketch.c:208 */
#undef WMO_EMIT
	return n;
}

/* canonical k-mer of a string as in src/index.c:362-376 (encodeKmer) */
uint64_t wmo_encode_kmer(const char *s, int k)
{
	uint64_t kmer[2] = {0, 0}, shift1 = 2 * (k - 1);
	int i;
	for (i = 0; i < k; ++i) {
		int c = wmo_nt4[(uint8_t)s[i]];
		kmer[0] = kmer[0] << 2 | (uint64_t)c;
		kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
	}
	return kmer[0] < kmer[1] ? kmer[0] : kmer[1];
}

/* ------------------------------------------------------------------------------------
 * Unstable in-place MSD radix sort + insertion sort: src/ksort.h:98-151 instantiated at
 * src/misc.c:156-159 (radix_sort_128x keyed on .x, radix_sort_64).  The tie order of the
 * cycle-leader permutation is part of the expected output.
 * ---------------------------------------------------------------------------------- */
typedef struct { uint64_t x, y; } wmo128_t;

#define WMO_RS(name, type_t, KEY) \
static void wmo_ins_##name(type_t *beg, type_t *end) \
{ /* ksort.h:105-115 */ \
	type_t *i; \
	for (i = beg + 1; i < end; ++i) \
		if (KEY(*i) < KEY(*(i - 1))) { \
			type_t *j, tmp = *i; \
			for (j = i; j > beg && KEY(tmp) < KEY(*(j - 1)); --j) *j = *(j - 1); \
			*j = tmp; \
		} \
} \
static void wmo_rs_##name(type_t *beg, type_t *end, int s) \
{ /* ksort.h:116-145 with n_bits = 8 */ \
	type_t *i; \
	struct { type_t *b, *e; } b[256], *k, *be = b + 256; \
	for (k = b; k != be; ++k) k->b = k->e = beg; \
	for (i = beg; i != end; ++i) ++b[KEY(*i) >> s & 255].e; \
	for (k = b + 1; k != be; ++k) k->e += (k - 1)->e - beg, k->b = (k - 1)->e; \
	for (k = b; k != be;) { \
		if (k->b != k->e) { \
			__typeof__(k) l; \
			if ((l = b + (KEY(*k->b) >> s & 255)) != k) { \
				type_t tmp = *k->b, swap; \
				do { \
					swap = tmp; tmp = *l->b; *l->b++ = swap; \
					l = b + (KEY(tmp) >> s & 255); \
				} while (l != k); \
				*k->b++ = tmp; \
			} else ++k->b; \
		} else ++k; \
	} \
	for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k - 1)->e; \
	if (s) { \
		s = s > 8 ? s - 8 : 0; \
		for (k = b; k != be; ++k) \
			if (k->e - k->b > 64) wmo_rs_##name(k->b, k->e, s); \
			else if (k->e - k->b > 1) wmo_ins_##name(k->b, k->e); \
	} \
} \
void wmo_radix_sort_##name(type_t *beg, long n) \
{ /* ksort.h:146-150 */ \
	if (n <= 64) wmo_ins_##name(beg, beg + n); \
	else wmo_rs_##name(beg, beg + n, 56); \
}
#define WMO_KEY128(a) ((a).x)
#define WMO_KEY64(a) (a)
WMO_RS(128x, wmo128_t, WMO_KEY128)
WMO_RS(64, uint64_t, WMO_KEY64)

/* ------------------------------------------------------------------------------------
 * Reference-side minimizer index, flattened: the lookup contract of mm_idx_get
 * (src/index.c:88-105) is "hash -> occurrence list sorted ascending by (rid,pos,strand)"
 * (src/index.c:239 sorts each list with radix_sort_64).  Any structure with that contract
 * is equivalent; here: all (hash,y) pairs sorted by (hash,y) + binary search.
 * ---------------------------------------------------------------------------------- */
typedef struct {
	long n;
	uint64_t *h, *y;
} wmo_idx_t;

static int wmo_cmp_hy(const void *a, const void *b)
{
	const wmo128_t *p = (const wmo128_t*)a, *q = (const wmo128_t*)b;
	if (p->x != q->x) return p->x < q->x ? -1 : 1;
	return p->y < q->y ? -1 : p->y > q->y;
}

/* minimizers: n (x,y) pairs as produced by wmo_sketch over all reference sequences */
wmo_idx_t *wmo_idx_build(const uint64_t *mz_xy, long n)
{
	wmo_idx_t *ix = (wmo_idx_t*)calloc(1, sizeof(wmo_idx_t));
	wmo128_t *t = (wmo128_t*)malloc((n > 0 ? n : 1) * sizeof(wmo128_t));
	long i;
	for (i = 0; i < n; ++i) t[i].x = mz_xy[2*i] >> 8, t[i].y = mz_xy[2*i+1];
	qsort(t, n, sizeof(wmo128_t), wmo_cmp_hy);
	ix->n = n; ix->h = (uint64_t*)malloc((n > 0 ? n : 1) * 8); ix->y = (uint64_t*)malloc((n > 0 ? n : 1) * 8);
	for (i = 0; i < n; ++i) ix->h[i] = t[i].x, ix->y[i] = t[i].y;
	free(t);
	return ix;
}
void wmo_idx_free(wmo_idx_t *ix) { if (ix) { free(ix->h); free(ix->y); free(ix); } }

const uint64_t *wmo_idx_get(const wmo_idx_t *ix, uint64_t minier, int *n)
{
	long lo = 0, hi = ix->n, e;
	while (lo < hi) { long m = (lo + hi) >> 1; if (ix->h[m] < minier) lo = m + 1; else hi = m; }
	for (e = lo; e < ix->n && ix->h[e] == minier; ++e) {}
	*n = (int)(e - lo);
	return *n ? ix->y + lo : 0;
}

/* ------------------------------------------------------------------------------------
 * Seed collection: collect_matches + collect_seed_hits, src/map.c:97-130 and :222-254
 * (flags NO_DIAG/NO_DUAL/FOR_ONLY/REV_ONLY all off => skip_seed() is always 0).
 * Outputs: anchors (sorted with the unstable radix sort), rep_len, mini_pos.
 * ---------------------------------------------------------------------------------- */
long wmo_collect_seed_hits(const wmo_idx_t *ix, int max_occ, const uint64_t *mv_xy, long n_mv, int qlen,
                           uint64_t *a_xy, long max_a, int *rep_len_, uint64_t *mini_pos, int *n_mini_pos_)
{
	int rep_st = 0, rep_en = 0, rep_len = 0, n_mini_pos = 0;
	long i, n_a = 0;
	for (i = 0; i < n_mv; ++i) {
		uint64_t px = mv_xy[2*i], py = mv_xy[2*i+1];
		uint32_t q_pos = (uint32_t)py, q_span = px & 0xff, k;
		int t, is_tandem = 0;
		const uint64_t *cr = wmo_idx_get(ix, px >> 8, &t);
		if (t >= max_occ) { /* map.c:111-116 */
			int en = (q_pos >> 1) + 1, st = en - (int)q_span;
			if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st, rep_en = en; }
			else rep_en = en;
			continue;
		}
		if (i > 0 && px >> 8 == mv_xy[2*(i-1)] >> 8) is_tandem = 1; /* map.c:121-122 */
		if (i < n_mv - 1 && px >> 8 == mv_xy[2*(i+1)] >> 8) is_tandem = 1;
		if (mini_pos) mini_pos[n_mini_pos] = (uint64_t)q_span << 32 | q_pos >> 1;
		++n_mini_pos;
		for (k = 0; k < (uint32_t)t; ++k) { /* map.c:233-249 */
			uint64_t r = cr[k], x, y;
			int32_t rpos = (uint32_t)r >> 1;
			if ((r & 1) == (q_pos & 1)) {
				x = (r & 0xffffffff00000000ULL) | (uint32_t)rpos;
				y = (uint64_t)q_span << 32 | q_pos >> 1;
			} else {
				x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | (uint32_t)rpos;
				y = (uint64_t)q_span << 32 | (uint32_t)(qlen - ((q_pos >> 1) + 1 - q_span) - 1);
			}
			y |= (uint64_t)(py >> 32) << 48;
			if (is_tandem) y |= 1ULL << 42;
			if (n_a < max_a) a_xy[2*n_a] = x, a_xy[2*n_a+1] = y;
			++n_a;
		}
	}
	rep_len += rep_en - rep_st;
	*rep_len_ = rep_len; *n_mini_pos_ = n_mini_pos;
	if (n_a <= max_a) wmo_radix_sort_128x((wmo128_t*)a_xy, n_a); /* map.c:252 */
	return n_a;
}

/* ------------------------------------------------------------------------------------
 * Chaining: src/chain.c:15-167 with n_segs == 1, is_cdna == 0.
 * a_xy is modified (re-ordered chains are written to b_xy).  Returns n_u.
 * ---------------------------------------------------------------------------------- */
static inline int wmo_ilog2_32(uint32_t v)
{ /* chain.c:8-20: floor(log2(v)) for v > 0 */
	int r = 0;
	while (v >>= 1) ++r;
	return r;
}

int wmo_chain_dp(int max_dist_x, int min_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                 float gap_scale, long n, uint64_t *a_xy, uint64_t *u_out, uint64_t *b_xy, long *n_b)
{
	wmo128_t *a = (wmo128_t*)a_xy, *b = (wmo128_t*)b_xy, *w;
	int32_t k, *f, *p, *t, *v, n_u, n_v;
	int64_t i, j, st = 0;
	uint64_t *u, *u2, sum_qspan = 0;
	float avg_qspan;
	*n_b = 0;
	if (n == 0) return 0;
	f = (int32_t*)malloc(n * 4); p = (int32_t*)malloc(n * 4);
	t = (int32_t*)calloc(n, 4); v = (int32_t*)malloc(n * 4);
	for (i = 0; i < n; ++i) sum_qspan += a[i].y >> 32 & 0xff;
	avg_qspan = (float)sum_qspan / n; /* chain.c:41-42 */
	for (i = 0; i < n; ++i) { /* chain.c:45-90 */
		uint64_t ri = a[i].x;
		int64_t max_j = -1;
		int32_t qi = (int32_t)a[i].y, q_span = a[i].y >> 32 & 0xff;
		int32_t max_f = q_span, n_skip = 0, min_d;
		while (st < i && ri > a[st].x + max_dist_x) ++st;
		if (i - st > max_iter)
			while (i - st > max_iter && ri > a[st].x + min_dist_x) ++st;
		for (j = i - 1; j >= st; --j) {
			int64_t dr = ri - a[j].x;
			int32_t dq = qi - (int32_t)a[j].y, dd, sc, log_dd, gap_cost;
			if (dr == 0 || dq <= 0) continue;
			if (dq > max_dist_y || dq > max_dist_x) continue;
			dd = dr > dq ? dr - dq : dq - dr;
			if (dd > bw) continue;
			min_d = dq < dr ? dq : dr;
			sc = min_d > q_span ? q_span : dq < dr ? dq : dr;
			log_dd = dd ? wmo_ilog2_32(dd) : 0;
			gap_cost = (int)(dd * .01 * avg_qspan) + (log_dd >> 1);
			sc -= (int)((double)gap_cost * gap_scale + .499);
			sc += f[j];
			if (sc > max_f) {
				max_f = sc, max_j = j;
				if (n_skip > 0) --n_skip;
			} else if (t[j] == i) {
				if (++n_skip > max_skip) break;
			}
			if (p[j] >= 0) t[p[j]] = i;
		}
		f[i] = max_f, p[i] = max_j;
		v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
	}
	/* chain.c:92-116 */
	memset(t, 0, n * 4);
	for (i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
	for (i = n_u = 0; i < n; ++i) if (t[i] == 0 && v[i] >= min_sc) ++n_u;
	if (n_u == 0) { free(f); free(p); free(t); free(v); return 0; }
	u = (uint64_t*)malloc(n_u * 8);
	for (i = n_u = 0; i < n; ++i)
		if (t[i] == 0 && v[i] >= min_sc) {
			j = i;
			while (j >= 0 && f[j] < v[j]) j = p[j];
			if (j < 0) j = i;
			u[n_u++] = (uint64_t)f[j] << 32 | j;
		}
	wmo_radix_sort_64(u, n_u);
	for (i = 0; i < n_u >> 1; ++i) { uint64_t tt = u[i]; u[i] = u[n_u - i - 1], u[n_u - i - 1] = tt; }
	/* chain.c:118-135 */
	memset(t, 0, n * 4);
	for (i = n_v = k = 0; i < n_u; ++i) {
		int32_t n_v0 = n_v, k0 = k;
		j = (int32_t)u[i];
		do { v[n_v++] = j; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
		if (j < 0) {
			if (n_v - n_v0 >= min_cnt) u[k++] = u[i] >> 32 << 32 | (n_v - n_v0);
		} else if ((int32_t)(u[i] >> 32) - f[j] >= min_sc) {
			if (n_v - n_v0 >= min_cnt) u[k++] = ((u[i] >> 32) - f[j]) << 32 | (n_v - n_v0);
		}
		if (k0 == k) n_v = n_v0;
	}
	n_u = k;
	/* chain.c:141-147 */
	for (i = 0, k = 0; i < n_u; ++i) {
		int32_t k0 = k, ni = (int32_t)u[i];
		for (j = 0; j < ni; ++j) b[k] = a[v[k0 + (ni - j - 1)]], ++k;
	}
	/* chain.c:149-165 */
	w = (wmo128_t*)malloc((n_u > 0 ? n_u : 1) * sizeof(wmo128_t));
	for (i = k = 0; i < n_u; ++i) { w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | i; k += (int32_t)u[i]; }
	wmo_radix_sort_128x(w, n_u);
	u2 = (uint64_t*)malloc((n_u > 0 ? n_u : 1) * 8);
	for (i = k = 0; i < n_u; ++i) {
		int32_t jj = (int32_t)w[i].y, nn = (int32_t)u[jj];
		u2[i] = u[jj];
		memcpy(&a[k], &b[w[i].y >> 32], nn * sizeof(wmo128_t));
		k += nn;
	}
	for (i = 0; i < n_u; ++i) u_out[i] = u2[i];
	if (k) memcpy(b, a, k * sizeof(wmo128_t));
	*n_b = k;
	free(f); free(p); free(t); free(v); free(u); free(u2); free(w);
	return n_u;
}

/* ------------------------------------------------------------------------------------
 * ksw_extd2: src/ksw2_extd2_sse.c:26-393, src/ksw2.h:103-176.  Scalar restatement that
 * keeps the SSE implementation's observable artefacts: 16-cell block rounding of the band
 * (:139), the contiguous [u|v|x|y|x2|y2|s|sf|qr] int8 layout (:99-102) incl. the unaligned
 * 16-byte score refresh that runs past en0 (:158-172), wrap-around int8 arithmetic, both
 * H-tracking modes (:315-375) and the backtrack over block bounds (ksw2.h:119-151).
 * ---------------------------------------------------------------------------------- */
#define WMO_NEG_INF (-0x40000000)
#define WMO_EZ_RIGHT      0x02
#define WMO_EZ_GENERIC_SC 0x04
#define WMO_EZ_SPLICE_FOR 0x100
#define WMO_EZ_SPLICE_REV 0x200
#define WMO_EZ_SPLICE_FLANK 0x400
#define WMO_EZ_APPROX_MAX 0x08
#define WMO_EZ_APPROX_DROP 0x10
#define WMO_EZ_EXTZ_ONLY  0x40
#define WMO_EZ_REV_CIGAR  0x80

typedef struct {
	int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end, n_cigar;
} wmo_ez_t;

static int wmo_apply_zdrop(wmo_ez_t *ez, int32_t H, int r, int t, int zdrop, int e)
{ /* ksw2.h:160-176, is_rot = 1 */
	if (H > ez->max) {
		ez->max = H, ez->max_t = t, ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && ez->max - H > zdrop + l * e) { ez->zdropped = 1; return 1; }
	}
	return 0;
}

typedef struct { uint32_t *a; int n, m; } wmo_cig_t;
static void wmo_push_cigar(wmo_cig_t *c, uint32_t op, int len)
{ /* ksw2.h:103-113 */
	if (c->n == 0 || op != (c->a[c->n - 1] & 0xf)) {
		if (c->n == c->m) { c->m = c->m ? c->m << 1 : 4; c->a = (uint32_t*)realloc(c->a, (size_t)c->m << 2); }
		c->a[c->n++] = (uint32_t)len << 4 | op;
	} else c->a[c->n - 1] += (uint32_t)len << 4;
}

static void wmo_backtrack(int is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0, wmo_cig_t *c)
{ /* ksw2.h:119-151, is_rot = 1, min_intron_len = 0 */
	int i = i0, j = j0, r, state = 0;
	uint32_t tmp;
	c->n = 0;
	while (i >= 0 && j >= 0) {
		int force_state = -1;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (i > off_end[r]) force_state = 1;
		tmp = force_state < 0 ? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) wmo_push_cigar(c, 0, 1), --i, --j;
		else if (state == 1 || state == 3) wmo_push_cigar(c, 2, 1), --i;
		else wmo_push_cigar(c, 1, 1), --j;
	}
	if (i >= 0) wmo_push_cigar(c, 2, i + 1);
	if (j >= 0) wmo_push_cigar(c, 1, j + 1);
	if (!is_rev)
		for (i = 0; i < c->n >> 1; ++i) tmp = c->a[i], c->a[i] = c->a[c->n - 1 - i], c->a[c->n - 1 - i] = tmp;
}

static void wmo_reset_ez(wmo_ez_t *ez)
{ /* ksw2.h:153-158 */
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = WMO_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
}

/* optional work counters (for the algorithmic-bytes accounting of SURVEY.md 8d) */
static uint64_t wmo_cnt_block_cells, wmo_cnt_band_cells, wmo_cnt_calls;
void wmo_counters_reset(void) { wmo_cnt_block_cells = wmo_cnt_band_cells = wmo_cnt_calls = 0; }
void wmo_counters_get(uint64_t *c) { c[0] = wmo_cnt_calls; c[1] = wmo_cnt_band_cells; c[2] = wmo_cnt_block_cells; }

int wmo_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                  int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag,
                  wmo_ez_t *ez, uint32_t *cigar_out, int max_cigar)
{
	const int m = 5;
	int r, t, qe = q + e, qe2, n_col_, qe_h = q + e /* :61: the scalar qe is taken BEFORE the swap at :70 */, *off = 0, *off_end = 0, tlen_, qlen_, last_st, last_en, max_sc, min_sc, long_thres, long_diff;
	int approx_max = !!(flag & WMO_EZ_APPROX_MAX), right = !!(flag & WMO_EZ_RIGHT);
	int32_t *H = 0, H0 = 0, last_H0_t = 0;
	int8_t *mem, *u, *v, *x, *y, *x2, *y2, *s, sc_N;
	uint8_t *sf, *qr, *p;
	wmo_cig_t cig = {0, 0, 0};

	wmo_reset_ez(ez);
	if (qlen <= 0 || tlen <= 0) return 0;
	if (q2 + e2 < q + e) t = q, q = q2, q2 = t, t = e, e = e2, e2 = t; /* :70 */
	qe = q + e, qe2 = q2 + e2;
	sc_N = mat[m*m-1] == 0 ? -e2 : mat[m*m-1]; /* :79 */
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return 0; /* :92 */
	long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0; /* :94-97 */
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	mem = (int8_t*)calloc((size_t)tlen_ * 8 + qlen_ + 1, 16); /* :99-102 */
	u = mem; v = u + tlen_ * 16; x = v + tlen_ * 16; y = x + tlen_ * 16; x2 = y + tlen_ * 16; y2 = x2 + tlen_ * 16;
	s = y2 + tlen_ * 16; sf = (uint8_t*)(s + tlen_ * 16); qr = sf + tlen_ * 16;
	memset(u, -q - e, tlen_ * 16); memset(v, -q - e, tlen_ * 16);
	memset(x, -q - e, tlen_ * 16); memset(y, -q - e, tlen_ * 16);
	memset(x2, -q2 - e2, tlen_ * 16); memset(y2, -q2 - e2, tlen_ * 16);
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)tlen_ * 16 * 4);
		for (t = 0; t < tlen_ * 16; ++t) H[t] = WMO_NEG_INF;
	}
	p = (uint8_t*)malloc(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
	off = (int*)malloc((size_t)(qlen + tlen - 1) * sizeof(int) * 2);
	off_end = off + qlen + tlen - 1;
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);
	++wmo_cnt_calls;

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, x21, v1;
		uint8_t *qrr = qr + (qlen - 1 - r);
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { ez->zdropped = 1; break; }
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1; /* :139 */
		wmo_cnt_band_cells += en0 - st0 + 1; wmo_cnt_block_cells += en - st + 1;
		if (st > 0) { /* :141-151 */
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = -q - e, x21 = -q2 - e2, v1 = -q - e;
		} else {
			x1 = -q - e, x21 = -q2 - e2;
			v1 = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		}
		if (en >= r) { /* :152-155 */
			y[r] = -q - e, y2[r] = -q2 - e2;
			u[r] = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		}
		for (t = st0; t <= en0; t += 16) { /* :158-172; runs up to 15 cells past en0 and may spill into sf[] */
			int kk;
			for (kk = 0; kk < 16; ++kk) {
				uint8_t sq = sf[t + kk], sq2 = qrr[t + kk];
				int8_t val = (sq == m - 1 || sq2 == m - 1) ? sc_N : (sq == sq2 ? mat[0] : mat[1]);
				s[t + kk] = val;
			}
		}
		off[r] = st, off_end[r] = en;
		{
			int8_t cx = x1, cv = v1, cx2 = x21; /* values of cell t-1 on the previous diagonal */
			uint8_t *pr = p + (size_t)r * n_col_ * 16;
			for (t = st; t <= en; ++t) {
				int8_t z = s[t], xo = x[t], vo = v[t], x2o = x2[t], uo = u[t], a, b, a2, b2, tmp, d;
				a = (int8_t)(cx + cv); b = (int8_t)(y[t] + uo);
				a2 = (int8_t)(cx2 + cv); b2 = (int8_t)(y2[t] + uo);
				if (!right) { /* :227-234 */
					d = 0;
					if (a > z) d = 1, z = a;
					if (b > z) d = 2, z = b;
					if (a2 > z) d = 3, z = a2;
					if (b2 > z) d = 4, z = b2;
				} else { /* :274-281 */
					d = z > a ? 0 : 1; z = z > a ? z : a;
					d = z > b ? d : 2; z = z > b ? z : b;
					d = z > a2 ? d : 3; z = z > a2 ? z : a2;
					d = z > b2 ? d : 4; z = z > b2 ? z : b2;
				}
				z = z < mat[0] ? z : mat[0];
				u[t] = (int8_t)(z - cv); v[t] = (int8_t)(z - uo); /* :51-52 */
				tmp = (int8_t)(z - q); a = (int8_t)(a - tmp); b = (int8_t)(b - tmp);
				tmp = (int8_t)(z - q2); a2 = (int8_t)(a2 - tmp); b2 = (int8_t)(b2 - tmp);
				if (!right) { /* :253-264 */
					x[t] = (int8_t)((a > 0 ? a : 0) - qe);   d |= a > 0 ? 0x08 : 0;
					y[t] = (int8_t)((b > 0 ? b : 0) - qe);   d |= b > 0 ? 0x10 : 0;
					x2[t] = (int8_t)((a2 > 0 ? a2 : 0) - qe2); d |= a2 > 0 ? 0x20 : 0;
					y2[t] = (int8_t)((b2 > 0 ? b2 : 0) - qe2); d |= b2 > 0 ? 0x40 : 0;
				} else { /* :300-311 */
					x[t] = (int8_t)((a < 0 ? 0 : a) - qe);   d |= a < 0 ? 0 : 0x08;
					y[t] = (int8_t)((b < 0 ? 0 : b) - qe);   d |= b < 0 ? 0 : 0x10;
					x2[t] = (int8_t)((a2 < 0 ? 0 : a2) - qe2); d |= a2 < 0 ? 0 : 0x20;
					y2[t] = (int8_t)((b2 < 0 ? 0 : b2) - qe2); d |= b2 < 0 ? 0 : 0x40;
				}
				pr[t - st] = (uint8_t)d;
				cx = xo, cv = vo, cx2 = x2o;
			}
		}
		if (!approx_max) { /* :315-358 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += (int32_t)v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe_h, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (wmo_apply_zdrop(ez, max_H, r, max_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* :359-375 */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t];
				} else {
					++last_H0_t, H0 += u[last_H0_t];
				}
			} else H0 = v[0] - qe_h, last_H0_t = 0;
			if ((flag & WMO_EZ_APPROX_DROP) && wmo_apply_zdrop(ez, H0, r, last_H0_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(mem); free(H);
	{ /* :379-391 */
		int rev_cigar = !!(flag & WMO_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & WMO_EZ_EXTZ_ONLY)) {
			wmo_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, tlen - 1, qlen - 1, &cig);
		} else if (!ez->zdropped && (flag & WMO_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) {
			ez->reach_end = 1;
			wmo_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, ez->mqe_t, qlen - 1, &cig);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			wmo_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, ez->max_t, ez->max_q, &cig);
		}
	}
	free(p); free(off);
	ez->n_cigar = cig.n;
	if (cigar_out) memcpy(cigar_out, cig.a, (size_t)(cig.n < max_cigar ? cig.n : max_cigar) * 4);
	free(cig.a);
	return ez->n_cigar;
}

/* ------------------------------------------------------------------------------------
 * ksw_extz2_sse: src/ksw2_extz2_sse.c:23-304 (single affine gap), the SSE4.1
 * code path (what ksw2_dispatch.c selects on the hosts used here).  Same skeleton as extd2, but the
 * state rows hold UNSIGNED offsets (u, v, x, y are the differences plus q + e resp. shifted so that
 * they are >= 0, rows start at 0 thanks to kcalloc, :106), the score is z = s + 2(q+e) (:27), the
 * three-way maximum mixes a signed compare for `a` with an unsigned maximum for `b` (:40, :163-166)
 * and the clamp with max_sc_ is unsigned (:41).  Every operation below is the 8-bit operation of the
 * reference (wrap-around add/sub, signed or unsigned compare as written there), so out-of-range
 * parameter sets misbehave identically.  One more artefact is kept: x1 / v1 are int8_t and go through
 * _mm_cvtsi32_si128 WITHOUT a cast (:153-154), so a negative value ORs 0xff into the next three
 * cells of the first block (:30,:34).
 * ---------------------------------------------------------------------------------- */
int wmo_ksw_extz2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                  int q, int e, int w, int zdrop, int end_bonus, int flag,
                  wmo_ez_t *ez, uint32_t *cigar_out, int max_cigar)
{
	const int m = 5;
	int r, t, qe = q + e, n_col_, *off = 0, *off_end = 0, tlen_, qlen_, last_st, last_en, max_sc, min_sc;
	int approx_max = !!(flag & WMO_EZ_APPROX_MAX), right = !!(flag & WMO_EZ_RIGHT);
	int32_t *H = 0, H0 = 0, last_H0_t = 0;
	uint8_t *mem, *u, *v, *x, *y, *sf, *qr, *p, max_sc_u8, qe2_u8 = (uint8_t)((q + e) * 2);
	int8_t *s, sc_N;
	wmo_cig_t cig = {0, 0, 0};

	wmo_reset_ez(ez);
	if (qlen <= 0 || tlen <= 0) return 0;
	sc_N = mat[m*m-1] == 0 ? -e : mat[m*m-1]; /* :79 */
	max_sc_u8 = (uint8_t)(mat[0] + (q + e) * 2); /* :81 */
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return 0; /* :94 */

	mem = (uint8_t*)calloc((size_t)tlen_ * 6 + qlen_ + 1, 16); /* :96-98: zero initialised */
	u = mem; v = u + tlen_ * 16; x = v + tlen_ * 16; y = x + tlen_ * 16; s = (int8_t*)(y + tlen_ * 16);
	sf = (uint8_t*)(s + tlen_ * 16); qr = sf + tlen_ * 16;
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)tlen_ * 16 * 4);
		for (t = 0; t < tlen_ * 16; ++t) H[t] = WMO_NEG_INF;
	}
	p = (uint8_t*)malloc(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
	off = (int*)malloc((size_t)(qlen + tlen - 1) * sizeof(int) * 2);
	off_end = off + qlen + tlen - 1;
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, v1;
		uint8_t *qrr = qr + (qlen - 1 - r);
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { ez->zdropped = 1; break; }
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		if (st > 0) { /* :133-139 */
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = (int8_t)x[st - 1], v1 = (int8_t)v[st - 1];
			else x1 = v1 = 0;
		} else x1 = 0, v1 = r ? q : 0;
		if (en >= r) y[r] = 0, u[r] = r ? q : 0;
		for (t = st0; t <= en0; t += 16) { /* :142-160 */
			int kk;
			for (kk = 0; kk < 16; ++kk) {
				uint8_t sq = sf[t + kk], sq2 = qrr[t + kk];
				s[t + kk] = (sq == m - 1 || sq2 == m - 1) ? sc_N : (sq == sq2 ? mat[0] : mat[1]);
			}
		}
		off[r] = st, off_end[r] = en;
		{
			uint8_t cx = (uint8_t)x1, cv = (uint8_t)v1; /* cells t-1 of the previous diagonal */
			uint8_t orx = x1 < 0 ? 0xff : 0, orv = v1 < 0 ? 0xff : 0; /* sign bytes of _mm_cvtsi32_si128(int8_t) */
			uint8_t *pr = p + (size_t)r * n_col_ * 16;
			for (t = st; t <= en; ++t) {
				uint8_t xo = x[t], vo = v[t], ut = u[t], xt1 = cx, vt1 = cv, z, a, b, d, zq;
				if (t - st >= 1 && t - st <= 3) xt1 |= orx, vt1 |= orv;
				z = (uint8_t)((uint8_t)s[t] + qe2_u8);
				a = (uint8_t)(xt1 + vt1);
				b = (uint8_t)(y[t] + ut);
				if (!right) { /* :212-221 */
					d = (int8_t)a > (int8_t)z ? 1 : 0;
					z = (int8_t)z > (int8_t)a ? z : a;
					d = (int8_t)b > (int8_t)z ? 2 : d;
				} else { /* :247-256 */
					d = (int8_t)z > (int8_t)a ? 0 : 1;
					z = (int8_t)z > (int8_t)a ? z : a;
					d = (int8_t)z > (int8_t)b ? d : 2;
				}
				z = z > b ? z : b;                     /* _mm_max_epu8 (:40) */
				z = z < max_sc_u8 ? z : max_sc_u8;     /* _mm_min_epu8 (:41) */
				u[t] = (uint8_t)(z - vt1); v[t] = (uint8_t)(z - ut);
				zq = (uint8_t)(z - (uint8_t)q); a = (uint8_t)(a - zq); b = (uint8_t)(b - zq);
				if (!right) { /* :223-229 */
					x[t] = (int8_t)a > 0 ? a : 0; d |= (int8_t)a > 0 ? 0x08 : 0;
					y[t] = (int8_t)b > 0 ? b : 0; d |= (int8_t)b > 0 ? 0x10 : 0;
				} else { /* :258-264 */
					x[t] = 0 > (int8_t)a ? 0 : a; d |= 0 > (int8_t)a ? 0 : 0x08;
					y[t] = 0 > (int8_t)b ? 0 : b; d |= 0 > (int8_t)b ? 0 : 0x10;
				}
				pr[t - st] = d;
				cx = xo, cv = vo;
			}
		}
		if (!approx_max) { /* :267-321 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] - qe : H[en0] + v[en0] - qe;
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += (int32_t)v[t + i] - qe;
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += (int32_t)v[t] - qe;
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe - qe, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (wmo_apply_zdrop(ez, max_H, r, max_t, zdrop, e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* :322-338 */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t] - qe, d1 = u[last_H0_t + 1] - qe;
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t] - qe;
				} else {
					++last_H0_t, H0 += u[last_H0_t] - qe;
				}
				if ((flag & WMO_EZ_APPROX_DROP) && wmo_apply_zdrop(ez, H0, r, last_H0_t, zdrop, e)) break;
			} else H0 = v[0] - qe - qe, last_H0_t = 0;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(mem); free(H);
	{ /* :343-355 */
		int rev_cigar = !!(flag & WMO_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & WMO_EZ_EXTZ_ONLY)) {
			wmo_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, tlen - 1, qlen - 1, &cig);
		} else if (!ez->zdropped && (flag & WMO_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) {
			ez->reach_end = 1;
			wmo_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, ez->mqe_t, qlen - 1, &cig);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			wmo_backtrack(rev_cigar, p, off, off_end, n_col_ * 16, ez->max_t, ez->max_q, &cig);
		}
	}
	free(p); free(off);
	ez->n_cigar = cig.n;
	if (cigar_out) memcpy(cigar_out, cig.a, (size_t)(cig.n < max_cigar ? cig.n : max_cigar) * 4);
	free(cig.a);
	return ez->n_cigar;
}

/* ------------------------------------------------------------------------------------
 * ksw_exts2_sse: src/ksw2_exts2_sse.c:26-408, the splice-aware extension (reached from mm_align_pair when
 * MM_F_SPLICE is set, src/align.c:326-327).  One affine gap (q, e) plus a long deletion state x2 that opens at q2,
 * extends for free and is entered / left through the donor[] / acceptor[] site costs (:100-165); no band, no
 * end bonus, H as in the dual-affine code (signed bytes).  Restated cell by cell over whole 16-cell blocks
 * with the reference's 8-bit wrap-around arithmetic (SSE4.1 branch).  junc may be null (:113,:126).
 * ---------------------------------------------------------------------------------- */
static void wmo_backtrack_intron(int is_rev, int min_intron_len, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0, wmo_cig_t *c)
{ /* ksw2.h:119-151, is_rot = 1, min_intron_len > 0: the long-gap state reads as N_SKIP (3) */
	int i = i0, j = j0, r, state = 0;
	uint32_t tmp;
	c->n = 0;
	while (i >= 0 && j >= 0) {
		int force_state = -1;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (i > off_end[r]) force_state = 1;
		tmp = force_state < 0 ? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) wmo_push_cigar(c, 0, 1), --i, --j;
		else if (state == 1 || (state == 3 && min_intron_len <= 0)) wmo_push_cigar(c, 2, 1), --i;
		else if (state == 3 && min_intron_len > 0) wmo_push_cigar(c, 3, 1), --i;
		else wmo_push_cigar(c, 1, 1), --j;
	}
	if (i >= 0) wmo_push_cigar(c, min_intron_len > 0 && i >= min_intron_len ? 3 : 2, i + 1);
	if (j >= 0) wmo_push_cigar(c, 1, j + 1);
	if (!is_rev)
		for (i = 0; i < c->n >> 1; ++i) tmp = c->a[i], c->a[i] = c->a[c->n - 1 - i], c->a[c->n - 1 - i] = tmp;
}

int wmo_ksw_exts2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                  int q, int e, int q2, int noncan, int zdrop, int junc_bonus, int flag, const uint8_t *junc,
                  wmo_ez_t *ez, uint32_t *cigar_out, int max_cigar)
{
	const int m = 5;
	int r, t, qe = q + e, n_col_, *off, *off_end, tlen_, qlen_, last_st, last_en, max_sc, min_sc, long_thres, long_diff;
	int approx_max = !!(flag & WMO_EZ_APPROX_MAX), right = !!(flag & WMO_EZ_RIGHT);
	int32_t *H = 0, H0 = 0, last_H0_t = 0;
	int8_t *mem, *u, *v, *x, *y, *x2, *donor, *acceptor, *s, sc_N;
	uint8_t *sf, *qr, *p;
	wmo_cig_t cig = {0, 0, 0};

	wmo_reset_ez(ez);
	if (qlen <= 0 || tlen <= 0 || q2 <= q + e) return 0; /* :61 */
	sc_N = mat[m*m-1] == 0 ? -e : mat[m*m-1]; /* :69 */
	tlen_ = (tlen + 15) / 16;
	n_col_ = ((qlen < tlen ? qlen : tlen) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return 0; /* :79 */
	long_thres = (q2 - q) / e - 1; /* :81-84 */
	if (q2 > q + e + long_thres * e) ++long_thres;
	long_diff = long_thres * e - (q2 - q);

	mem = (int8_t*)calloc((size_t)tlen_ * 9 + qlen_ + 1, 16); /* :86-90 */
	u = mem; v = u + tlen_ * 16; x = v + tlen_ * 16; y = x + tlen_ * 16; x2 = y + tlen_ * 16;
	donor = x2 + tlen_ * 16; acceptor = donor + tlen_ * 16;
	s = acceptor + tlen_ * 16; sf = (uint8_t*)(s + tlen_ * 16); qr = sf + tlen_ * 16;
	memset(u, -q - e, (size_t)tlen_ * 16 * 4);
	memset(x2, -q2, (size_t)tlen_ * 16);
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)tlen_ * 16 * 4);
		for (t = 0; t < tlen_ * 16; ++t) H[t] = WMO_NEG_INF;
	}
	p = (uint8_t*)malloc(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
	off = (int*)malloc((size_t)(qlen + tlen - 1) * sizeof(int) * 2);
	off_end = off + qlen + tlen - 1;
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);

	if (flag & (WMO_EZ_SPLICE_FOR | WMO_EZ_SPLICE_REV)) { /* :106-165 */
		int semi_cost = flag & WMO_EZ_SPLICE_FLANK ? -noncan / 2 : 0;
		int fw = !!(flag & WMO_EZ_SPLICE_FOR), rv = !!(flag & WMO_EZ_SPLICE_REV);
		memset(donor, -noncan, (size_t)tlen_ * 16);
		memset(acceptor, -noncan, (size_t)tlen_ * 16);
		if (!(flag & WMO_EZ_REV_CIGAR)) {
			for (t = 0; t < tlen - 4; ++t) {
				int can_type = 0;
				if (fw && target[t+1] == 2 && target[t+2] == 3) can_type = 1;
				if (rv && target[t+1] == 1 && target[t+2] == 3) can_type = 1;
				if (can_type && (target[t+3] == 0 || target[t+3] == 2)) can_type = 2;
				if (can_type) donor[t] = can_type == 2 ? 0 : semi_cost;
			}
			if (junc)
				for (t = 0; t < tlen - 1; ++t)
					if ((fw && (junc[t+1] & 1)) || (rv && (junc[t+1] & 8))) donor[t] += junc_bonus;
			for (t = 2; t < tlen; ++t) {
				int can_type = 0;
				if (fw && target[t-1] == 0 && target[t] == 2) can_type = 1;
				if (rv && target[t-1] == 0 && target[t] == 1) can_type = 1;
				if (can_type && (target[t-2] == 1 || target[t-2] == 3)) can_type = 2;
				if (can_type) acceptor[t] = can_type == 2 ? 0 : semi_cost;
			}
			if (junc)
				for (t = 0; t < tlen; ++t)
					if ((fw && (junc[t] & 2)) || (rv && (junc[t] & 4))) acceptor[t] += junc_bonus;
		} else {
			for (t = 0; t < tlen - 4; ++t) {
				int can_type = 0;
				if (fw && target[t+1] == 2 && target[t+2] == 0) can_type = 1;
				if (rv && target[t+1] == 1 && target[t+2] == 0) can_type = 1;
				if (can_type && (target[t+3] == 1 || target[t+3] == 3)) can_type = 2;
				if (can_type) donor[t] = can_type == 2 ? 0 : semi_cost;
			}
			if (junc)
				for (t = 0; t < tlen - 1; ++t)
					if ((fw && (junc[t+1] & 2)) || (rv && (junc[t+1] & 4))) donor[t] += junc_bonus;
			for (t = 2; t < tlen; ++t) {
				int can_type = 0;
				if (fw && target[t-1] == 3 && target[t] == 2) can_type = 1;
				if (rv && target[t-1] == 3 && target[t] == 1) can_type = 1;
				if (can_type && (target[t-2] == 0 || target[t-2] == 2)) can_type = 2;
				if (can_type) acceptor[t] = can_type == 2 ? 0 : semi_cost;
			}
			if (junc)
				for (t = 0; t < tlen; ++t)
					if ((fw && (junc[t] & 1)) || (rv && (junc[t] & 8))) acceptor[t] += junc_bonus;
		}
	}

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, x21, v1;
		uint8_t *qrr = qr + (qlen - 1 - r);
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		if (st > 0) { /* :178-186 */
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = -q - e, x21 = -q2, v1 = -q - e;
		} else {
			x1 = -q - e, x21 = -q2;
			v1 = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : 0;
		}
		if (en >= r) {
			y[r] = -q - e;
			u[r] = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : 0;
		}
		if (!(flag & WMO_EZ_GENERIC_SC)) { /* :192-211 */
			for (t = st0; t <= en0; t += 16) {
				int kk;
				for (kk = 0; kk < 16; ++kk) {
					uint8_t sq = sf[t + kk], sq2 = qrr[t + kk];
					s[t + kk] = (sq == m - 1 || sq2 == m - 1) ? sc_N : (sq == sq2 ? mat[0] : mat[1]);
				}
			}
		} else {
			for (t = st0; t <= en0; ++t) s[t] = mat[sf[t] * m + qrr[t]];
		}
		off[r] = st, off_end[r] = en;
		{
			int8_t cx = x1, cx2 = x21, cv = v1; /* cells t - 1 of the previous diagonal (the casts at :214-216 are to uint8_t) */
			uint8_t *pr = p + (size_t)r * n_col_ * 16;
			for (t = st; t <= en; ++t) {
				int8_t xo = x[t], x2o = x2[t], vo = v[t], ut = u[t], z = s[t], a, b, a2, a2a, tmp, dn = donor[t];
				uint8_t d;
				a = (int8_t)(cx + cv); b = (int8_t)(y[t] + ut); a2 = (int8_t)(cx2 + cv); a2a = (int8_t)(a2 + acceptor[t]);
				if (!right) { /* :252-259 */
					d = a > z ? 1 : 0;    z = z > a ? z : a;
					d = b > z ? 2 : d;    z = z > b ? z : b;
					d = a2a > z ? 3 : d;  z = z > a2a ? z : a2a;
				} else { /* :300-307 */
					d = z > a ? 0 : 1;    z = z > a ? z : a;
					d = z > b ? d : 2;    z = z > b ? z : b;
					d = z > a2a ? d : 3;  z = z > a2a ? z : a2a;
				}
				u[t] = (int8_t)(z - cv); v[t] = (int8_t)(z - ut); /* :51-56 */
				tmp = (int8_t)(z - q); a = (int8_t)(a - tmp); b = (int8_t)(b - tmp);
				a2 = (int8_t)(a2 - (int8_t)(z - q2));
				if (!right) { /* :272-290 */
					x[t] = (int8_t)((a > 0 ? a : 0) - qe); d |= a > 0 ? 0x08 : 0;
					y[t] = (int8_t)((b > 0 ? b : 0) - qe); d |= b > 0 ? 0x10 : 0;
					x2[t] = (int8_t)((a2 > dn ? a2 : dn) - q2); d |= a2 > dn ? 0x20 : 0;
				} else { /* :320-338 */
					x[t] = (int8_t)((0 > a ? 0 : a) - qe); d |= 0 > a ? 0 : 0x08;
					y[t] = (int8_t)((0 > b ? 0 : b) - qe); d |= 0 > b ? 0 : 0x10;
					x2[t] = (int8_t)((dn > a2 ? dn : a2) - q2); d |= dn > a2 ? 0 : 0x20;
				}
				pr[t - st] = d;
				cx = xo, cx2 = x2o, cv = vo;
			}
		}
		if (!approx_max) { /* :341-389 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += (int32_t)v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += (int32_t)v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (wmo_apply_zdrop(ez, max_H, r, max_t, zdrop, 0)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* :390-406 */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t];
				} else {
					++last_H0_t, H0 += u[last_H0_t];
				}
			} else H0 = v[0] - qe, last_H0_t = 0;
			if ((flag & WMO_EZ_APPROX_DROP) && wmo_apply_zdrop(ez, H0, r, last_H0_t, zdrop, 0)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(mem); free(H);
	{ /* :411-419: no end-bonus clause here */
		int rev_cigar = !!(flag & WMO_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & WMO_EZ_EXTZ_ONLY))
			wmo_backtrack_intron(rev_cigar, long_thres, p, off, off_end, n_col_ * 16, tlen - 1, qlen - 1, &cig);
		else if (ez->max_t >= 0 && ez->max_q >= 0)
			wmo_backtrack_intron(rev_cigar, long_thres, p, off, off_end, n_col_ * 16, ez->max_t, ez->max_q, &cig);
	}
	free(p); free(off);
	ez->n_cigar = cig.n;
	if (cigar_out) memcpy(cigar_out, cig.a, (size_t)(cig.n < max_cigar ? cig.n : max_cigar) * 4);
	free(cig.a);
	return ez->n_cigar;
}

/* ------------------------------------------------------------------------------------
 * ksw_ll_i16: src/ksw2_ll_sse.c:32-147 (score, qe, te of a striped int16 local SW).
 * Restated as the row-by-row Gotoh recurrence the striped kernel evaluates, over the
 * padded query (slen*8 columns, padding scores 0, :70-75), with the same tie rules:
 * te = last row whose maximum is >= the running best (:138), qe = last striped slot
 * holding the best (:144-145).  Values saturate like adds_epi16 / subs_epu16.
 * ---------------------------------------------------------------------------------- */
int wmo_ksw_ll(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int gapo, int gape, int *qe_, int *te_)
{
	int slen = (qlen + 7) / 8, qlen8 = slen * 8, i, j, gmax = 0, gapoe = gapo + gape;
	int *H0 = (int*)calloc(qlen8 + 1, sizeof(int)), *H1 = (int*)calloc(qlen8 + 1, sizeof(int));
	int *E = (int*)calloc(qlen8 + 1, sizeof(int)), *Hmax = (int*)calloc(qlen8 + 1, sizeof(int));
	*qe_ = *te_ = -1;
	for (i = 0; i < tlen; ++i) {
		int f = 0, imax = 0, *tmp;
		const int8_t *ma = mat + target[i] * 5;
		for (j = 0; j < qlen8; ++j) {
			int sc = j < qlen ? ma[query[j]] : 0;
			int hd = j > 0 ? H0[j - 1] : 0, h, e = E[j], t;
			h = hd + sc; if (h > 32767) h = 32767; if (h < -32768) h = -32768;
			if (h < e) h = e;
			if (h < f) h = f;
			H1[j] = h;
			if (h > imax) imax = h;
			t = h - gapoe; if (t < 0) t = 0;
			e -= gape; if (e < 0) e = 0;
			E[j] = e > t ? e : t;
			f -= gape; if (f < 0) f = 0;
			if (f < t) f = t;
		}
		if (imax >= gmax) { gmax = imax; *te_ = i; memcpy(Hmax, H1, qlen8 * sizeof(int)); }
		tmp = H1; H1 = H0; H0 = tmp;
	}
	{ /* striped order: slot i <-> column (i/8) + (i%8)*slen */
		int best = -1;
		for (i = 0; i < qlen8; ++i) {
			int col = i / 8 + i % 8 * slen;
			if (Hmax[col] == gmax) best = col;
		}
		*qe_ = best;
	}
	free(H0); free(H1); free(E); free(Hmax);
	return gmax;
}
