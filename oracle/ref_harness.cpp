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

// ---- ksw2 (src/ksw2_extd2_sse.c:26, src/ksw2_extz2_sse.c:23) ----
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

} // extern "C"
