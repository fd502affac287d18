// TEST INFRASTRUCTURE ONLY. Thin extern "C" wrappers around the REAL reference
// (libwinnowmap.a built by oracle/build_ref.sh from /root/reference, unmodified apart
// from the documented rep_len=0 init) so that python tests can pin oracle/wm_oracle.c
// and the CUDA kernels against the reference's own functions.  Built into
// oracle/_ref/libref_harness.so; nothing in the product path links or loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "minimap.h"
#include "mmpriv.h"
#include "ksw2.h"
#include "kalloc.h"

extern "C" {

// ---- ksw2 (src/ksw2_extd2_sse.c:26, src/ksw2_extz2_sse.c:23, src/ksw2_exts2_sse.c:26) ----
// out_ez: max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, reach_end, n_cigar
int ref_ksw_extd2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, const int8_t *mat,
                  int gapo, int gape, int gapo2, int gape2, int w, int zdrop, int end_bonus, int flag,
                  int *out_ez, uint32_t *cigar, int max_cigar)
{
	ksw_extz_t ez; memset(&ez, 0, sizeof(ez));
	ksw_extd2_sse(0, qlen, q, tlen, t, 5, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, &ez);
	out_ez[0] = ez.max; out_ez[1] = ez.zdropped; out_ez[2] = ez.max_q; out_ez[3] = ez.max_t;
	out_ez[4] = ez.mqe; out_ez[5] = ez.mqe_t; out_ez[6] = ez.mte; out_ez[7] = ez.mte_q;
	out_ez[8] = ez.score; out_ez[9] = ez.reach_end; out_ez[10] = ez.n_cigar;
	int n = ez.n_cigar < max_cigar ? ez.n_cigar : max_cigar;
	if (n > 0) memcpy(cigar, ez.cigar, n * 4);
	kfree(0, ez.cigar);
	return ez.n_cigar;
}

int ref_ksw_extz2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, const int8_t *mat,
                  int gapo, int gape, int w, int zdrop, int end_bonus, int flag,
                  int *out_ez, uint32_t *cigar, int max_cigar)
{
	ksw_extz_t ez; memset(&ez, 0, sizeof(ez));
	ksw_extz2_sse(0, qlen, q, tlen, t, 5, mat, gapo, gape, w, zdrop, end_bonus, flag, &ez);
	out_ez[0] = ez.max; out_ez[1] = ez.zdropped; out_ez[2] = ez.max_q; out_ez[3] = ez.max_t;
	out_ez[4] = ez.mqe; out_ez[5] = ez.mqe_t; out_ez[6] = ez.mte; out_ez[7] = ez.mte_q;
	out_ez[8] = ez.score; out_ez[9] = ez.reach_end; out_ez[10] = ez.n_cigar;
	int n = ez.n_cigar < max_cigar ? ez.n_cigar : max_cigar;
	if (n > 0) memcpy(cigar, ez.cigar, n * 4);
	kfree(0, ez.cigar);
	return ez.n_cigar;
}

// splice-aware extension (src/ksw2_exts2_sse.c:26); junc may be null
int ref_ksw_exts2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, const int8_t *mat,
                  int gapo, int gape, int gapo2, int noncan, int zdrop, int junc_bonus, int flag, const uint8_t *junc,
                  int *out_ez, uint32_t *cigar, int max_cigar)
{
	ksw_extz_t ez; memset(&ez, 0, sizeof(ez));
	ksw_exts2_sse(0, qlen, q, tlen, t, 5, mat, gapo, gape, gapo2, noncan, zdrop, junc_bonus, flag, junc, &ez);
	out_ez[0] = ez.max; out_ez[1] = ez.zdropped; out_ez[2] = ez.max_q; out_ez[3] = ez.max_t;
	out_ez[4] = ez.mqe; out_ez[5] = ez.mqe_t; out_ez[6] = ez.mte; out_ez[7] = ez.mte_q;
	out_ez[8] = ez.score; out_ez[9] = ez.reach_end; out_ez[10] = ez.n_cigar;
	int n = ez.n_cigar < max_cigar ? ez.n_cigar : max_cigar;
	if (n > 0) memcpy(cigar, ez.cigar, n * 4);
	kfree(0, ez.cigar);
	return ez.n_cigar;
}

// ---- ksw_ll (src/ksw2_ll_sse.c:32,80) ----
int ref_ksw_ll(int qlen, const uint8_t *q, int tlen, const uint8_t *t, const int8_t *mat, int gapo, int gape, int *qe, int *te)
{
	void *qp = ksw_ll_qinit(0, 2, qlen, q, 5, mat);
	int sc = ksw_ll_i16(qp, tlen, t, gapo, gape, qe, te);
	kfree(0, qp);
	return sc;
}

// ---- sorts (src/misc.c:156-159) ----
void ref_radix_sort_128x(uint64_t *xy, long n) { radix_sort_128x((mm128_t*)xy, (mm128_t*)xy + n); }
void ref_radix_sort_64(uint64_t *a, long n) { radix_sort_64(a, a + n); }

// ---- bloom + sketch (src/index.c:404-432, src/sketch.c:128) ----
struct ref_sk { mm_idx_t mi; };

void *ref_sketch_ctx(int n_kmers, const uint64_t *canon_kmers)
{
	ref_sk *c = (ref_sk*)calloc(1, sizeof(ref_sk));
	bloom_parameters parameters;
	uint64_t cnt = n_kmers;
	parameters.projected_element_count = cnt > 1000 ? cnt : 1000;
	parameters.false_positive_probability = 0.001;
	parameters.maximum_number_of_hashes = 2;
	parameters.compute_optimal_parameters();
	c->mi.downFilter = new bloom_filter(parameters);
	for (int i = 0; i < n_kmers; ++i) c->mi.downFilter->insert(canon_kmers[i]);
	return c;
}
uint64_t ref_bloom_size(void *ctx) { return ((ref_sk*)ctx)->mi.downFilter->size(); }
void ref_bloom_table(void *ctx, uint8_t *out) { ref_sk *c = (ref_sk*)ctx; memcpy(out, c->mi.downFilter->table(), c->mi.downFilter->size() / 8); }
int ref_bloom_contains(void *ctx, uint64_t key) { return ((ref_sk*)ctx)->mi.downFilter->contains(key); }
void ref_sketch_free(void *ctx) { ref_sk *c = (ref_sk*)ctx; delete c->mi.downFilter; free(c); }

long ref_sketch(void *ctx, const char *seq, int len, int w, int k, uint32_t rid, uint64_t *out_xy, long max_out)
{
	ref_sk *c = (ref_sk*)ctx;
	mm128_v v = {0, 0, 0};
	mm_sketch(0, seq, len, w, k, rid, 0, &v, &c->mi);
	long n = (long)v.n < max_out ? (long)v.n : max_out;
	memcpy(out_xy, v.a, n * 16);
	long tot = v.n;
	kfree(0, v.a);
	return tot;
}

// ---- chaining (src/chain.c:22) ----
// a_xy is copied (mm_chain_dp frees its input); returns n_u, writes u[] and the compacted anchors
int ref_chain_dp(int max_dist_x, int min_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                 float gap_scale, long n, const uint64_t *a_xy, uint64_t *u_out, uint64_t *b_out, long *n_b)
{
	mm128_t *a = (mm128_t*)kmalloc(0, (n > 0 ? n : 1) * 16);
	memcpy(a, a_xy, n * 16);
	int n_u = 0; uint64_t *u = 0;
	mm128_t *b = mm_chain_dp(max_dist_x, min_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, gap_scale, 0, 1, n, a, &n_u, &u, 0);
	long nb = 0;
	for (int i = 0; i < n_u; ++i) { u_out[i] = u[i]; nb += (int32_t)u[i]; }
	if (nb) memcpy(b_out, b, nb * 16);
	*n_b = nb;
	kfree(0, b); kfree(0, u);
	return n_u;
}

// ---- the reference's own index, flattened the way INTEGRATION.md section 3 shows (src/index.c:33-38, :88-105) ----
} // extern "C" (khash instantiation below needs C++ linkage-neutral statics)
#include "khash.h"
#define ref_idx_hash(a) ((a)>>1)
#define ref_idx_eq(a, b) ((a)>>1 == (b)>>1)
KHASH_INIT(refidx, uint64_t, uint64_t, 1, ref_idx_hash, ref_idx_eq) // the table type of src/index.c:25-27 under another name (same layout)
typedef struct ref_bucket_s { // mm_idx_bucket_t, src/index.c:33-38 (opaque in minimap.h)
	mm128_v a;
	int32_t n;
	uint64_t *p;
	void *h;
} ref_bucket_t;
#include <algorithm>
#include <utility>
#include <vector>
struct ref_flat {
	mm_idx_t *mi;
	std::vector<uint64_t> keys, pos_off, pos, seq_off;
	std::vector<uint32_t> seq_len;
	std::vector<const char*> names;
};
extern "C" {

// mm_idx_reader_open / mm_idx_reader_read (src/index.c:688,:713) on a FASTA file, then the bucket walk
void *ref_idx_build_flat(const char *fn, const char *kmer_fn, int w, int k, int n_threads)
{
	mm_idxopt_t io; mm_mapopt_t mo;
	mm_set_opt(0, &io, &mo);
	io.k = k, io.w = w;
	mm_idx_reader_t *r = mm_idx_reader_open(fn, &io, 0);
	if (!r) return 0;
	mm_idx_t *mi = mm_idx_reader_read(r, n_threads, kmer_fn ? kmer_fn : "");
	mm_idx_reader_close(r);
	if (!mi) return 0;
	ref_flat *f = new ref_flat(); f->mi = mi;
	std::vector<std::pair<uint64_t, std::pair<const uint64_t*, int> > > all;
	const ref_bucket_t *B = (const ref_bucket_t*)mi->B;
	for (uint32_t b = 0; b < 1U << mi->b; ++b) {
		khash_t(refidx) *h = (khash_t(refidx)*)B[b].h;
		if (h == 0) continue;
		for (khint_t x = 0; x < kh_end(h); ++x) {
			if (!kh_exist(h, x)) continue;
			const uint64_t minier = (kh_key(h, x) >> 1) << mi->b | b; // inverse of src/index.c:90-96
			if (kh_key(h, x) & 1) all.push_back(std::make_pair(minier, std::make_pair((const uint64_t*)&kh_val(h, x), 1)));
			else all.push_back(std::make_pair(minier, std::make_pair((const uint64_t*)&B[b].p[kh_val(h, x) >> 32], (int)(uint32_t)kh_val(h, x))));
		}
	}
	std::sort(all.begin(), all.end());
	for (size_t i = 0; i < all.size(); ++i) {
		f->keys.push_back(all[i].first); f->pos_off.push_back(f->pos.size());
		f->pos.insert(f->pos.end(), all[i].second.first, all[i].second.first + all[i].second.second);
	}
	f->pos_off.push_back(f->pos.size());
	for (uint32_t i = 0; i < mi->n_seq; ++i) { f->names.push_back(mi->seq[i].name); f->seq_len.push_back(mi->seq[i].len); f->seq_off.push_back(mi->seq[i].offset); }
	return f;
}
// sizes: [n_seq, S_words, n_keys, n_pos, bloom_bits, k, w]
void ref_idx_flat_sizes(void *p, uint64_t *out)
{
	ref_flat *f = (ref_flat*)p;
	uint64_t sum = 0;
	for (uint32_t i = 0; i < f->mi->n_seq; ++i) sum += f->mi->seq[i].len;
	out[0] = f->mi->n_seq, out[1] = (sum + 7) / 8, out[2] = f->keys.size(), out[3] = f->pos.size(), out[4] = f->mi->downFilter->size();
	out[5] = f->mi->k, out[6] = f->mi->w;
}
const uint64_t *ref_idx_flat_keys(void *p) { return ((ref_flat*)p)->keys.data(); }
const uint64_t *ref_idx_flat_pos_off(void *p) { return ((ref_flat*)p)->pos_off.data(); }
const uint64_t *ref_idx_flat_pos(void *p) { return ((ref_flat*)p)->pos.data(); }
const uint32_t *ref_idx_flat_S(void *p) { return ((ref_flat*)p)->mi->S; }
const uint32_t *ref_idx_flat_seq_len(void *p) { return ((ref_flat*)p)->seq_len.data(); }
const uint64_t *ref_idx_flat_seq_off(void *p) { return ((ref_flat*)p)->seq_off.data(); }
const char *const *ref_idx_flat_names(void *p) { return ((ref_flat*)p)->names.data(); }
const uint8_t *ref_idx_flat_bloom(void *p) { return ((ref_flat*)p)->mi->downFilter->table(); }
void ref_idx_flat_free(void *p) { ref_flat *f = (ref_flat*)p; mm_idx_destroy(f->mi); delete f; }

// ---- ABI: sizes and field offsets of the public structs (src/minimap.h:80-176) for tests/test_abi_layout.py ----
#include <stddef.h>
int ref_abi_layout(int64_t *out, int cap)
{
	int n = 0;
#define PUT(v) do { if (n < cap) out[n] = (int64_t)(v); ++n; } while (0)
	PUT(sizeof(mm_mapopt_t)); PUT(sizeof(mm_reg1_t)); PUT(sizeof(mm_extra_t)); PUT(sizeof(mm_idxopt_t));
#define MO(f) PUT(offsetof(mm_mapopt_t, f))
	MO(flag); MO(seed); MO(sdust_thres); MO(max_qlen); MO(bw); MO(max_gap); MO(max_gap_ref); MO(min_gap_ref); MO(max_frag_len); MO(max_chain_skip);
	MO(max_chain_iter); MO(min_cnt); MO(min_chain_score); MO(chain_gap_scale); MO(SVaware); MO(SVawareMinReadLength); MO(suffixSampleOffset);
	MO(min_mapq); MO(min_qcov); MO(minPrefixLength); MO(maxPrefixLength); MO(prefixIncrementFactor); MO(stage2_bw); MO(stage2_zdrop_inv);
	MO(stage2_max_gap); MO(stage2_extension_inc); MO(mask_level); MO(mask_len); MO(pri_ratio); MO(best_n); MO(max_join_long); MO(max_join_short);
	MO(min_join_flank_sc); MO(min_join_flank_ratio); MO(alt_drop); MO(a); MO(b); MO(q); MO(e); MO(q2); MO(e2); MO(sc_ambi); MO(noncan); MO(junc_bonus);
	MO(zdrop); MO(zdrop_inv); MO(end_bonus); MO(min_dp_max); MO(min_ksw_len); MO(anchor_ext_len); MO(anchor_ext_shift); MO(max_clip_ratio);
	MO(pe_ori); MO(pe_bonus); MO(mid_occ_frac); MO(min_mid_occ); MO(mid_occ); MO(max_occ); MO(mini_batch_size); MO(max_sw_mat);
	MO(kmer_freq_filename); MO(split_prefix);
#define RG(f) PUT(offsetof(mm_reg1_t, f))
	RG(id); RG(cnt); RG(rid); RG(score); RG(qs); RG(qe); RG(rs); RG(re); RG(parent); RG(subsc); RG(as); RG(mlen); RG(blen); RG(n_sub); RG(score0); RG(hash); RG(div); RG(p);
#define EX(f) PUT(offsetof(mm_extra_t, f))
	EX(capacity); EX(dp_score); EX(dp_max); EX(dp_max2); EX(n_cigar); EX(cigar);
#define IO(f) PUT(offsetof(mm_idxopt_t, f))
	IO(k); IO(w); IO(flag); IO(bucket_bits); IO(mini_batch_size); IO(batch_size);
	return n;
}

} // extern "C"
