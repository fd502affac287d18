// Weighted robust-winnowing minimizers on sm_100a: the output of mm_sketch
// (reference src/sketch.c:128-219, weights :70-90, bloom probe ext/bloom/bloom_filter.hpp:303-319,
// :551-565) for a batch of sequences, restructured into three data-parallel passes:
//
//   A  per base: canonical k-mer -> order key (double, murmur64 / 2^64, x^8 when the k-mer is in the
//      down-weight filter) and an "eligible" bit = the key is strictly below the w previous keys.
//   B  the winnowing state machine is history dependent only through ties, so it can be restarted
//      exactly at any eligible position (a strict new minimum is certain there, whatever the
//      history).  One thread per 128-base chunk starts at the first eligible position of its chunk
//      and runs until the first eligible position of a later chunk, flagging emitted minimizers.
//   C  flags -> per-chunk counts -> exclusive scan -> (x,y) records in position order.
//
// For odd k (the defaults 15/19) a k-mer cannot equal its reverse complement, so ring slots and
// bases coincide (src/sketch.c:166 never fires); even k takes a sequential per-sequence kernel.
#include <math.h>
#include <string.h>
#include <vector>
#include "wm_common.cuh"
#include "scan.cuh"
#include "sketch.cuh"

#define WM_SK_TN 1024      // new positions per tile in pass A
#define WM_SK_THREADS 256
#define WM_SK_CH 128       // chunk size in pass B/C

__device__ __forceinline__ uint64_t wm_murmur64(uint64_t key)
{ // src/sketch.c:43-51
	key ^= key >> 33; key *= 0xff51afd7ed558ccdULL;
	key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ULL;
	key ^= key >> 33;
	return key;
}
__device__ __forceinline__ uint64_t wm_hash64(uint64_t key, uint64_t mask)
{ // src/sketch.c:53-63
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}
__device__ __forceinline__ bool wm_bloom_contains(const wm_bloom_dev &bf, uint64_t key)
{ // bloom_filter.hpp:303-319 with hash_ap on 8 bytes (:556-565) and compute_indices (:461-465)
	const uint32_t i1 = (uint32_t)key, i2 = (uint32_t)(key >> 32);
	for (int i = 0; i < bf.n_salt; ++i) {
		uint32_t h = bf.salt[i];
		h ^= (h << 7) ^ (i1 * (h >> 3)) ^ (~((h << 11) + (i2 ^ (h >> 5))));
		uint64_t bit = (uint64_t)h % bf.bits;
		if (!((bf.table[bit >> 3] >> (bit & 7)) & 1)) return false;
	}
	return true;
}
__device__ __forceinline__ double wm_weight(uint64_t kmer, const wm_bloom_dev &bf)
{ // src/sketch.c:70-90; (double)UINT64_MAX is 2^64, so the division is an exact scaling
	const uint64_t h = wm_murmur64(kmer);
	const double x = __ull2double_rn(h) * 5.421010862427522e-20; // 2^-64
	if (wm_bloom_contains(bf, kmer)) {
		double p2 = __dmul_rn(x, x), p4 = __dmul_rn(p2, p2);
		return -__dmul_rn(p4, p4);
	}
	return -x;
}

// canonical k-mer ending at position i of `c` (codes 0..4); returns false if it spans an N / the start
__device__ __forceinline__ bool wm_kmer_at(const uint8_t *c, int k, uint64_t *fw, uint64_t *rv)
{ // c points at the first base of the k-mer
	uint64_t f = 0, r = 0;
	bool ok = true;
	for (int j = 0; j < k; ++j) {
		uint64_t b = c[j];
		ok &= b < 4;
		f = f << 2 | (b & 3);
		r = r >> 2 | (3ULL ^ (b & 3)) << (2 * (k - 1));
	}
	*fw = f, *rv = r;
	return ok;
}

// ---- pass A ----
__global__ void __launch_bounds__(WM_SK_THREADS)
wm_sketch_order_kernel(const uint8_t *__restrict__ codes, const wm_sk_task *__restrict__ tasks, const int64_t *__restrict__ tile_off,
                       const int64_t *__restrict__ base_off, int n_tasks, int w, int k, wm_bloom_dev bf,
                       double *__restrict__ ord, uint8_t *__restrict__ elig)
{
	extern __shared__ __align__(16) uint8_t sm_raw[];
	double *s_ord = (double*)sm_raw;                       // WM_SK_TN + w entries, position p0 - w + j
	uint8_t *s_code = sm_raw + (size_t)(WM_SK_TN + 256) * 8;  // WM_SK_TN + w + k - 1 codes, position p0 - w - (k-1) + j
	// which task does this tile belong to?
	int lo = 0, hi = n_tasks;
	const int64_t tile = blockIdx.x;
	while (hi - lo > 1) { int m = (lo + hi) >> 1; if (tile_off[m] <= tile) lo = m; else hi = m; }
	const wm_sk_task T = tasks[lo];
	const int p0 = (int)(tile - tile_off[lo]) * WM_SK_TN;
	const uint8_t *seq = codes + T.seq_off;
	const int n_code = WM_SK_TN + w + k - 1, c_start = p0 - w - (k - 1);
	for (int j = threadIdx.x; j < n_code; j += WM_SK_THREADS) {
		int p = c_start + j;
		s_code[j] = (p >= 0 && p < T.len) ? seq[p] : 4;
	}
	__syncthreads();
	for (int j = threadIdx.x; j < WM_SK_TN + w; j += WM_SK_THREADS) {
		const int p = p0 - w + j;
		double o = 2.0;
		if (p >= k - 1 && p < T.len) {
			uint64_t f, r;
			if (wm_kmer_at(s_code + j, k, &f, &r) && f != r) o = wm_weight(f < r ? f : r, bf);
		}
		s_ord[j] = o;
	}
	__syncthreads();
	const int64_t gb = base_off[lo];
	for (int j = threadIdx.x; j < WM_SK_TN; j += WM_SK_THREADS) {
		const int p = p0 + j;
		if (p >= T.len) break;
		const double o = s_ord[j + w];
		bool e = o < 2.0;
		if (e) {
			double m = 2.0;
			for (int d = 0; d < w; ++d) m = fmin(m, s_ord[j + d]);
			e = o < m;
		}
		ord[gb + p] = o;
		elig[gb + p] = e ? 1 : 0;
	}
}

// ---- pass B ----
__global__ void wm_sketch_winnow_kernel(const wm_sk_task *__restrict__ tasks, const int64_t *__restrict__ chunk_off,
                                        const int64_t *__restrict__ base_off, int n_tasks, int64_t n_chunks, int w,
                                        const double *__restrict__ ord_all, const uint8_t *__restrict__ elig_all, uint8_t *__restrict__ flag_all)
{
	const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ch >= n_chunks) return;
	int lo = 0, hi = n_tasks;
	while (hi - lo > 1) { int m = (lo + hi) >> 1; if (chunk_off[m] <= ch) lo = m; else hi = m; }
	const int len = tasks[lo].len;
	const int64_t gb = base_off[lo];
	const double *ord = ord_all + gb;
	const uint8_t *elig = elig_all + gb;
	uint8_t *flag = flag_all + gb;
	const int c0 = (int)(ch - chunk_off[lo]) * WM_SK_CH;
	int c1 = c0 + WM_SK_CH; if (c1 > len) c1 = len;
	int min_i = -1, v = 0, i;
	double min_ord = 2.0;
	if (c0 == 0) i = 0;
	else {
		int s = c0;
		while (s < c1 && !elig[s]) ++s;
		if (s >= c1) return; // an earlier thread runs through this chunk
		min_i = s, min_ord = ord[s];
		for (int j = s; j >= 0 && v < w + 1 && ord[j] < 2.0; --j) ++v;
		i = s + 1;
	}
	bool seen = false; // an eligible position was already met in the chunk that contains i
	for (; i < len; ++i) {
		if ((i & (WM_SK_CH - 1)) == 0) seen = false;
		const double o = ord[i];
		v = o < 2.0 ? (v < w + 1 ? v + 1 : v) : 0;
		if (o < min_ord) { // a new minimum (src/sketch.c:180-189)
			if (v >= w + 1 && min_i >= 0) flag[min_i] = 1;
			min_i = i, min_ord = o;
		} else if (min_i >= 0 && i - min_i == w) { // the old minimum left the window (:191-205)
			if (v >= w) flag[min_i] = 1;
			int m = -1; double mo = 2.0;
			for (int j = i - w + 1; j <= i; ++j) {
				double oj = ord[j];
				if (mo >= oj) mo = oj, m = j; // ">=": the closest (rightmost) k-mer wins
			}
			if (mo >= 2.0) m = -1, mo = 2.0;
			min_i = m, min_ord = mo;
		}
		if (elig[i]) {
			if (!seen && i >= c1) return; // the thread of that chunk resumes from here
			seen = true;
		}
	}
	if (min_i >= 0) flag[min_i] = 1; // :208
}

// sequential pass for even k (symmetric k-mers make ring slots != bases): one thread per sequence
__global__ void wm_sketch_seq_kernel(const uint8_t *__restrict__ codes, const wm_sk_task *__restrict__ tasks, const int64_t *__restrict__ base_off,
                                     int n_tasks, int w, int k, wm_bloom_dev bf, uint8_t *__restrict__ flag_all, double *__restrict__ ring_all)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	if (tid >= n_tasks) return;
	const wm_sk_task T = tasks[tid];
	const uint8_t *seq = codes + T.seq_off;
	uint8_t *flag = flag_all + base_off[tid];
	double *buf_ord = ring_all + (size_t)tid * 512; // order keys
	double *buf_pos_ = buf_ord + 256;               // positions, stored as doubles (exact below 2^53)
	const uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1;
	uint64_t kmer0 = 0, kmer1 = 0;
	int l = 0, buf_pos = 0, min_pos = 0, min_i = -1;
	double min_ord = 2.0;
	for (int j = 0; j < w; ++j) buf_ord[j] = 2.0, buf_pos_[j] = -1.0;
	for (int i = 0; i < T.len; ++i) {
		const int c = seq[i];
		double o = 2.0; int oi = -1;
		if (c < 4) {
			kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
			kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
			if (kmer0 == kmer1) continue; // src/sketch.c:166
			++l;
			if (l >= k) o = wm_weight(kmer0 < kmer1 ? kmer0 : kmer1, bf), oi = i;
		} else l = 0;
		buf_ord[buf_pos] = o, buf_pos_[buf_pos] = (double)oi;
		if (o < min_ord) {
			if (l >= w + k && min_i >= 0) flag[min_i] = 1;
			min_i = oi, min_pos = buf_pos, min_ord = o;
		} else if (buf_pos == min_pos) {
			if (l >= w + k - 1 && min_i >= 0) flag[min_i] = 1;
			min_i = -1, min_ord = 2.0;
			for (int j = buf_pos + 1; j < w; ++j) if (min_ord >= buf_ord[j]) min_i = (int)buf_pos_[j], min_pos = j, min_ord = buf_ord[j];
			for (int j = 0; j <= buf_pos; ++j) if (min_ord >= buf_ord[j]) min_i = (int)buf_pos_[j], min_pos = j, min_ord = buf_ord[j];
		}
		if (++buf_pos == w) buf_pos = 0;
	}
	if (min_i >= 0) flag[min_i] = 1;
}

// ---- pass C ----
__global__ void wm_sketch_count_kernel(const wm_sk_task *__restrict__ tasks, const int64_t *__restrict__ chunk_off, const int64_t *__restrict__ base_off,
                                       int n_tasks, int64_t n_chunks, const uint8_t *__restrict__ flag_all, int32_t *__restrict__ cnt)
{
	const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ch >= n_chunks) return;
	int lo = 0, hi = n_tasks;
	while (hi - lo > 1) { int m = (lo + hi) >> 1; if (chunk_off[m] <= ch) lo = m; else hi = m; }
	const int len = tasks[lo].len;
	const uint8_t *flag = flag_all + base_off[lo];
	const int c0 = (int)(ch - chunk_off[lo]) * WM_SK_CH;
	int c1 = c0 + WM_SK_CH; if (c1 > len) c1 = len;
	int n = 0;
	for (int i = c0; i < c1; ++i) n += flag[i];
	cnt[ch] = n;
}

__global__ void wm_sketch_emit_kernel(const uint8_t *__restrict__ codes, const wm_sk_task *__restrict__ tasks, const int64_t *__restrict__ chunk_off,
                                      const int64_t *__restrict__ base_off, int n_tasks, int64_t n_chunks, int k,
                                      const uint8_t *__restrict__ flag_all, const int64_t *__restrict__ rank, wm128_dev *__restrict__ out,
                                      int64_t *__restrict__ mz_off)
{
	const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ch > n_chunks) return;
	if (ch == n_chunks) { mz_off[n_tasks] = rank[n_chunks]; return; }
	int lo = 0, hi = n_tasks;
	while (hi - lo > 1) { int m = (lo + hi) >> 1; if (chunk_off[m] <= ch) lo = m; else hi = m; }
	const wm_sk_task T = tasks[lo];
	const uint8_t *flag = flag_all + base_off[lo];
	const uint8_t *seq = codes + T.seq_off;
	const int c0 = (int)(ch - chunk_off[lo]) * WM_SK_CH;
	int c1 = c0 + WM_SK_CH; if (c1 > T.len) c1 = T.len;
	if (c0 == 0) mz_off[lo] = rank[ch];
	// empty sequences own no chunk: give them the offset of the next non-empty one
	if (c0 == 0) for (int t = lo - 1; t >= 0 && tasks[t].len == 0; --t) mz_off[t] = rank[ch];
	int64_t o = rank[ch];
	const uint64_t mask = (1ULL << 2 * k) - 1;
	for (int i = c0; i < c1; ++i)
		if (flag[i]) {
			uint64_t f, r;
			wm_kmer_at(seq + i - (k - 1), k, &f, &r);
			const int z = f < r ? 0 : 1; // src/sketch.c:167
			wm128_dev m;
			m.x = wm_hash64(z ? r : f, mask) << 8 | (uint64_t)k;   // :171 (span == k once l >= k)
			m.y = (uint64_t)T.rid << 32 | (uint32_t)i << 1 | (uint64_t)z; // :172
			out[o++] = m;
		}
}

__global__ void wm_ascii_to_code_kernel(const char *__restrict__ in, uint8_t *__restrict__ out, int64_t n)
{ // seq_nt4_table (src/sketch.c:19-36): ACGT/acgt (and U/u) -> 0..3, everything else 4
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned char c = in[i];
	uint8_t v = 4;
	switch (c) {
		case 'A': case 'a': v = 0; break;
		case 'C': case 'c': v = 1; break;
		case 'G': case 'g': v = 2; break;
		case 'T': case 't': case 'U': case 'u': v = 3; break;
		default: v = c < 4 ? c : 4; // the table maps bytes 0..3 to themselves
	}
	out[i] = v;
}

void wm_ascii_to_code(const char *d_in, uint8_t *d_out, int64_t n, cudaStream_t st)
{
	if (n <= 0) return;
	wm_count_launch(); wm_ascii_to_code_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, d_out, n);
	WM_CUDA_CHECK(cudaGetLastError());
}

// ---- host-side launcher on device-resident code arrays ----
// tasks (host copy) describe slices of d_codes.  On return *n_mz is the total number of minimizers,
// ws->mz holds them (device) and ws->mz_off (device, n_tasks+1) their per-task offsets.
void wm_sketch_run(wm_sketch_ws *ws, const wm_bloom_dev &bf, const uint8_t *d_codes, const wm_sk_task *h_tasks, int n_tasks,
                   int w, int k, int64_t *n_mz, cudaStream_t st)
{
	*n_mz = 0;
	std::vector<int64_t> h_off(3 * (size_t)(n_tasks + 1));
	int64_t *tile_off = h_off.data(), *chunk_off = tile_off + n_tasks + 1, *base_off = chunk_off + n_tasks + 1;
	tile_off[0] = chunk_off[0] = base_off[0] = 0;
	for (int i = 0; i < n_tasks; ++i) {
		int64_t L = h_tasks[i].len > 0 ? h_tasks[i].len : 0;
		tile_off[i + 1] = tile_off[i] + (L + WM_SK_TN - 1) / WM_SK_TN;
		chunk_off[i + 1] = chunk_off[i] + (L + WM_SK_CH - 1) / WM_SK_CH;
		base_off[i + 1] = base_off[i] + L;
	}
	const int64_t n_tiles = tile_off[n_tasks], n_chunks = chunk_off[n_tasks], n_bases = base_off[n_tasks];
	wm_sk_task *d_tasks = (wm_sk_task*)ws->tasks.need(sizeof(wm_sk_task) * (n_tasks + 1));
	int64_t *d_off = (int64_t*)ws->offs.need(sizeof(int64_t) * h_off.size());
	int64_t *d_mz_off = (int64_t*)ws->mz_off.need(sizeof(int64_t) * (n_tasks + 1));
	WM_CUDA_CHECK(wm_memcpy_async(d_tasks, h_tasks, sizeof(wm_sk_task) * n_tasks, cudaMemcpyHostToDevice, st));
	WM_CUDA_CHECK(wm_memcpy_async(d_off, h_off.data(), sizeof(int64_t) * h_off.size(), cudaMemcpyHostToDevice, st));
	WM_CUDA_CHECK(cudaMemsetAsync(d_mz_off, 0, sizeof(int64_t) * (n_tasks + 1), st));
	if (n_bases == 0 || n_tasks == 0) { wm_stream_sync(st); return; }
	const int64_t *d_tile_off = d_off, *d_chunk_off = d_off + n_tasks + 1, *d_base_off = d_chunk_off + n_tasks + 1;
	uint8_t *d_flag = (uint8_t*)ws->flag.need(n_bases);
	WM_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, n_bases, st));
	if (k & 1) {
		double *d_ord = (double*)ws->ord.need(sizeof(double) * n_bases);
		uint8_t *d_elig = (uint8_t*)ws->elig.need(n_bases);
		const size_t smem = (size_t)(WM_SK_TN + 256) * 8 + WM_SK_TN + 256 + 32;
		wm_count_launch(); wm_sketch_order_kernel<<<(unsigned)n_tiles, WM_SK_THREADS, smem, st>>>(d_codes, d_tasks, d_tile_off, d_base_off, n_tasks, w, k, bf, d_ord, d_elig);
		WM_CUDA_CHECK(cudaGetLastError());
		wm_count_launch(); wm_sketch_winnow_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(d_tasks, d_chunk_off, d_base_off, n_tasks, n_chunks, w, d_ord, d_elig, d_flag);
		WM_CUDA_CHECK(cudaGetLastError());
	} else {
		double *d_ring = (double*)ws->ord.need(sizeof(double) * 512 * (size_t)n_tasks);
		wm_count_launch(); wm_sketch_seq_kernel<<<(n_tasks + 63) / 64, 64, 0, st>>>(d_codes, d_tasks, d_base_off, n_tasks, w, k, bf, d_flag, d_ring);
		WM_CUDA_CHECK(cudaGetLastError());
	}
	int32_t *d_cnt = (int32_t*)ws->cnt.need(sizeof(int32_t) * (n_chunks + 1));
	int64_t *d_rank = (int64_t*)ws->rank.need(sizeof(int64_t) * (n_chunks + 2));
	int64_t *d_tmp = (int64_t*)ws->scan_tmp.need(sizeof(int64_t) * wm_scan_tmp_elems(n_chunks));
	wm_count_launch(); wm_sketch_count_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(d_tasks, d_chunk_off, d_base_off, n_tasks, n_chunks, d_flag, d_cnt);
	WM_CUDA_CHECK(cudaGetLastError());
	wm_exclusive_scan(d_cnt, n_chunks, d_rank, d_tmp, st);
	int64_t total = 0;
	WM_CUDA_CHECK(wm_memcpy_async(&total, d_rank + n_chunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	wm_stream_sync(st);
	wm128_dev *d_mz = (wm128_dev*)ws->mz.need(sizeof(wm128_dev) * (total + 1));
	// trailing empty sequences: their offset is the total
	{
		std::vector<int64_t> fill(n_tasks + 1, total);
		WM_CUDA_CHECK(wm_memcpy_async(d_mz_off, fill.data(), sizeof(int64_t) * (n_tasks + 1), cudaMemcpyHostToDevice, st));
	}
	wm_count_launch(); wm_sketch_emit_kernel<<<(unsigned)((n_chunks + 1 + 127) / 128), 128, 0, st>>>(d_codes, d_tasks, d_chunk_off, d_base_off, n_tasks, n_chunks, k,
	                                                                               d_flag, d_rank, d_mz, d_mz_off);
	WM_CUDA_CHECK(cudaGetLastError());
	*n_mz = total;
}

// ---- down-weight filter construction (host; replaces bloom_filter of src/index.c:404-432) ----
struct wm_bloom_s {
	uint64_t bits; uint32_t salt[2]; int n_salt;
	std::vector<uint8_t> table;
};

static inline uint32_t wm_hash_ap8_host(uint64_t key, uint32_t h)
{
	uint32_t i1 = (uint32_t)key, i2 = (uint32_t)(key >> 32);
	h ^= (h << 7) ^ (i1 * (h >> 3)) ^ (~((h << 11) + (i2 ^ (h >> 5))));
	return h;
}

extern "C" wm_bloom_s *wm_bloom_build(const uint64_t *canon_kmers, int64_t n)
{
	wm_bloom_s *b = new wm_bloom_s();
	// bloom_parameters::compute_optimal_parameters (bloom_filter.hpp:108-147) with
	// projected_element_count = max(n,1000), fpp = 0.001, maximum_number_of_hashes = 2 (index.c:411-414)
	const double cnt = (double)(n > 1000 ? n : 1000), p = 0.001;
	double min_m = INFINITY, min_k = 0.0;
	for (double kk = 1.0; kk < 1000.0; kk += 1.0) {
		double m = (-kk * cnt) / log(1.0 - pow(p, 1.0 / kk));
		if (m < min_m) min_m = m, min_k = kk;
	}
	unsigned nh = (unsigned)min_k;
	uint64_t ts = (uint64_t)min_m;
	ts += (ts % 8) != 0 ? 8 - ts % 8 : 0;
	if (nh < 1) nh = 1; else if (nh > 2) nh = 2;
	b->bits = ts, b->n_salt = (int)nh;
	const uint64_t seed = 0xA5A5A5A55A5A5A5AULL * 0xA5A5A5A5ULL + 1; // bloom_filter.hpp:186
	const uint32_t predef[2] = { 0xAAAAAAAAu, 0x55555555u };          // :477
	for (int i = 0; i < b->n_salt; ++i) b->salt[i] = predef[i];
	for (int i = 0; i < b->n_salt; ++i) b->salt[i] = b->salt[i] * b->salt[(i + 3) % b->n_salt] + (uint32_t)seed; // :519-528
	b->table.assign(ts / 8 + 16, 0);
	for (int64_t i = 0; i < n; ++i)
		for (int s = 0; s < b->n_salt; ++s) {
			uint64_t bit = (uint64_t)wm_hash_ap8_host(canon_kmers[i], b->salt[s]) % ts;
			b->table[bit >> 3] |= (uint8_t)(1u << (bit & 7));
		}
	return b;
}
extern "C" uint64_t wm_bloom_bits(const wm_bloom_s *b) { return b->bits; }
extern "C" const uint8_t *wm_bloom_table(const wm_bloom_s *b) { return b->table.data(); }
extern "C" void wm_bloom_destroy(wm_bloom_s *b) { delete b; }

void wm_bloom_params(const wm_bloom_s *b, uint64_t *bits, uint32_t *salt, int *n_salt)
{
	*bits = b->bits; salt[0] = b->salt[0]; salt[1] = b->salt[1]; *n_salt = b->n_salt;
}

// salts depend only on the number of hashes, which is 2 for every table built by the reference
void wm_bloom_dev_from_table(wm_bloom_dev *d, const uint8_t *d_table, uint64_t bits)
{
	const uint64_t seed = 0xA5A5A5A55A5A5A5AULL * 0xA5A5A5A5ULL + 1;
	d->table = d_table; d->bits = bits; d->n_salt = 2;
	d->salt[0] = 0xAAAAAAAAu; d->salt[1] = 0x55555555u;
	for (int i = 0; i < 2; ++i) d->salt[i] = d->salt[i] * d->salt[(i + 3) % 2] + (uint32_t)seed;
}

// ---- C ABI: batched mm_sketch ----
extern "C" int wm_sketch_batch(const wm_bloom_s *bloom, int n, const char *seq, const int64_t *off, const uint32_t *rid,
                               int w, int k, wm128_dev **out, int64_t **out_off)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
		fprintf(stderr, "[ERROR] wm_sketch_batch: no CUDA device visible; winnowmap-b200 has no CPU fallback\n");
		exit(1);
	}
	if (!(w > 0 && w < 256 && k > 0 && k <= 28)) { // assert at src/sketch.c:140
		fprintf(stderr, "[ERROR] wm_sketch_batch: invalid (w,k)\n");
		return -1;
	}
	*out = 0; *out_off = (int64_t*)calloc(n + 1, sizeof(int64_t));
	if (n <= 0) return 0;
	const int64_t tot = off[n];
	char *d_ascii = wm_dev_alloc<char>(tot + 1);
	uint8_t *d_codes = wm_dev_alloc<uint8_t>(tot + 1);
	WM_CUDA_CHECK(cudaMemcpy(d_ascii, seq, tot, cudaMemcpyHostToDevice));
	wm_ascii_to_code(d_ascii, d_codes, tot, 0);
	uint8_t *d_table = wm_dev_alloc<uint8_t>(bloom->table.size());
	WM_CUDA_CHECK(cudaMemcpy(d_table, bloom->table.data(), bloom->table.size(), cudaMemcpyHostToDevice));
	wm_bloom_dev bf; wm_bloom_dev_from_table(&bf, d_table, bloom->bits);
	bf.n_salt = bloom->n_salt; bf.salt[0] = bloom->salt[0]; bf.salt[1] = bloom->salt[1];
	std::vector<wm_sk_task> tasks(n);
	for (int i = 0; i < n; ++i) tasks[i].seq_off = off[i], tasks[i].len = (int32_t)(off[i + 1] - off[i]), tasks[i].rid = rid ? rid[i] : 0;
	wm_sketch_ws ws;
	int64_t n_mz = 0;
	wm_sketch_run(&ws, bf, d_codes, tasks.data(), n, w, k, &n_mz, 0);
	WM_CUDA_CHECK(cudaDeviceSynchronize());
	*out = (wm128_dev*)malloc(sizeof(wm128_dev) * (n_mz > 0 ? n_mz : 1));
	if (n_mz > 0) WM_CUDA_CHECK(cudaMemcpy(*out, ws.mz.p, sizeof(wm128_dev) * n_mz, cudaMemcpyDeviceToHost));
	WM_CUDA_CHECK(cudaMemcpy(*out_off, ws.mz_off.p, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost));
	ws.release();
	cudaFree(d_ascii); cudaFree(d_codes); cudaFree(d_table);
	return 0;
}
