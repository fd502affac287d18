// Chain -> hit glue of the mapping path, host side: tiny arrays, branchy,
// float/libm arithmetic whose roundings must match the reference's x86-64 build (no FMA), so it stays
// on the CPU by design.  Each function states the reference code whose result it reproduces.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "host_glue.h"
#include "host_sort.h"

namespace wmh {

static inline uint64_t mix64(uint64_t key)
{ // the unmasked Wang-style mixer of src/hit.c:39-49
	key = (~key + (key << 21));
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8));
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4));
	key = key ^ key >> 28;
	key = (key + (key << 31));
	return key;
}

static inline int span_of(const wm_pair_t &p) { return (int)(p.y >> 32 & 0xff); }

// seeded match / block length of a chain (mm_cal_fuzzy_len, src/hit.c:8-21)
static void fuzzy_len(wm_reg1_t *r, const wm_pair_t *a)
{
	r->mlen = r->blen = 0;
	if (r->cnt <= 0) return;
	r->mlen = r->blen = span_of(a[r->as]);
	for (int i = r->as + 1; i < r->as + r->cnt; ++i) {
		const int span = span_of(a[i]);
		const int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x;
		const int ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
		r->blen += tl > ql ? tl : ql;
		r->mlen += (tl > span && ql > span) ? span : (tl < ql ? tl : ql);
	}
}

// mm_reg_set_coor (src/hit.c:23-37)
void set_coor(wm_reg1_t *r, int32_t qlen, const wm_pair_t *a)
{
	const int32_t k = r->as, q_span = span_of(a[k]);
	const wm_pair_t &first = a[k], &last = a[k + r->cnt - 1];
	r->rev = first.x >> 63;
	r->rid = (int32_t)(first.x << 1 >> 33);
	r->rs = (int32_t)first.x + 1 > q_span ? (int32_t)first.x + 1 - q_span : 0;
	r->re = (int32_t)last.x + 1;
	if (!r->rev) {
		r->qs = (int32_t)first.y + 1 - q_span;
		r->qe = (int32_t)last.y + 1;
	} else {
		r->qs = qlen - ((int32_t)last.y + 1);
		r->qe = qlen - ((int32_t)first.y + 1 - q_span);
	}
	fuzzy_len(r, a);
}

// mm_gen_regs (src/hit.c:52-90): chains -> hits ordered by (score, hash) descending
void gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const wm_pair_t *a, std::vector<wm_reg1_t> &regs)
{
	regs.clear();
	if (n_u == 0) return;
	std::vector<wm_pair_t> z(n_u);
	for (int i = 0, k = 0; i < n_u; ++i) {
		const uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ hash);
		z[i].x = u[i] ^ h;
		z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
		k += (int32_t)u[i];
	}
	radix_sort(z.data(), z.data() + n_u);
	std::reverse(z.begin(), z.end());
	regs.resize(n_u);
	for (int i = 0; i < n_u; ++i) {
		wm_reg1_t *ri = &regs[i];
		memset(ri, 0, sizeof(wm_reg1_t));
		ri->id = i;
		ri->parent = WM_PARENT_UNSET;
		ri->score = ri->score0 = (int32_t)(z[i].x >> 32);
		ri->hash = (uint32_t)z[i].x;
		ri->cnt = (int32_t)z[i].y;
		ri->as = (int32_t)(z[i].y >> 32);
		ri->div = -1.0f;
		set_coor(ri, qlen, a);
	}
}

static inline int alt_score(int score, float alt_diff_frac)
{ // mm_alt_score (src/hit.c:100-105)
	if (score < 0) return score;
	score = (int)(score * (1.0 - alt_diff_frac) + .499);
	return score > 0 ? score : 1;
}

// mm_split_reg (src/hit.c:107-123)
void split_reg(wm_reg1_t *r, wm_reg1_t *r2, int n, int qlen, const wm_pair_t *a)
{
	if (n <= 0 || n >= r->cnt) return;
	*r2 = *r;
	r2->id = -1;
	r2->sam_pri = 0;
	r2->p = 0;
	r2->split_inv = 0;
	r2->cnt = r->cnt - n;
	r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
	r2->as = r->as + n;
	if (r->parent == r->id) r2->parent = WM_PARENT_TMP_PRI;
	set_coor(r2, qlen, a);
	r->cnt -= r2->cnt;
	r->score -= r2->score;
	set_coor(r, qlen, a);
	r->split |= 1, r2->split |= 2;
}

// mm_set_parent (src/hit.c:125-186): primary/secondary tree by query overlap; fills subsc / n_sub / dp_max2
void set_parent(float mask_level, int mask_len, int n, wm_reg1_t *r, int sub_diff, int hard_mask_level, float alt_diff_frac)
{
	if (n <= 0) return;
	for (int i = 0; i < n; ++i) r[i].id = i;
	std::vector<uint64_t> cov(n);
	std::vector<int> w(n);
	w[0] = 0, r[0].parent = 0;
	int k = 1;
	for (int i = 1; i < n; ++i) {
		wm_reg1_t *ri = &r[i];
		const int si = ri->qs, ei = ri->qe;
		int n_cov = 0, uncov_len = 0, j = 0;
		bool test_only = false;
		if (!hard_mask_level) {
			for (j = 0; j < k; ++j) { // overlaps with existing primaries
				const wm_reg1_t *rp = &r[w[j]];
				int sj = rp->qs, ej = rp->qe;
				if (ej <= si || sj >= ei) continue;
				if (sj < si) sj = si;
				if (ej > ei) ej = ei;
				cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
			}
			if (n_cov == 0) test_only = true; // a new primary; j == k here
			else { // length of ri not covered by primaries
				int x = si;
				radix_sort(cov.data(), cov.data() + n_cov);
				for (int c = 0; c < n_cov; ++c) {
					if ((int)(cov[c] >> 32) > x) uncov_len += (int)(cov[c] >> 32) - x;
					x = (int32_t)cov[c] > x ? (int32_t)cov[c] : x;
				}
				if (ei > x) uncov_len += ei - x;
			}
		}
		if (!test_only) {
			for (j = 0; j < k; ++j) {
				wm_reg1_t *rp = &r[w[j]];
				const int sj = rp->qs, ej = rp->qe;
				if (ej <= si || sj >= ei) continue;
				const int min = ej - sj < ei - si ? ej - sj : ei - si;
				const int max = ej - sj > ei - si ? ej - sj : ei - si;
				const int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
				if ((float)ol / min - (float)uncov_len / max > mask_level && uncov_len <= mask_len) { // secondary to rp
					int cnt_sub = 0, sci = ri->score;
					ri->parent = rp->parent;
					if (!rp->is_alt && ri->is_alt) sci = alt_score(sci, alt_diff_frac);
					rp->subsc = rp->subsc > sci ? rp->subsc : sci;
					if (ri->cnt >= rp->cnt) cnt_sub = 1;
					if (rp->p && ri->p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != min)) {
						sci = ri->p->dp_max;
						if (!rp->is_alt && ri->is_alt) sci = alt_score(sci, alt_diff_frac);
						rp->p->dp_max2 = rp->p->dp_max2 > sci ? rp->p->dp_max2 : sci;
						if (rp->p->dp_max - ri->p->dp_max <= sub_diff) cnt_sub = 1;
					}
					if (cnt_sub) ++rp->n_sub;
					break;
				}
			}
		}
		if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
}

// mm_set_sam_pri (src/hit.c:220-230)
int set_sam_pri(int n, wm_reg1_t *r)
{
	int n_pri = 0;
	for (int i = 0; i < n; ++i)
		if (r[i].id == r[i].parent) { ++n_pri; r[i].sam_pri = (n_pri == 1); }
		else r[i].sam_pri = 0;
	return n_pri;
}

// mm_sync_regs (src/hit.c:232-253)
void sync_regs(int n_regs, wm_reg1_t *regs)
{
	if (n_regs <= 0) return;
	int max_id = -1;
	for (int i = 0; i < n_regs; ++i) max_id = max_id > regs[i].id ? max_id : regs[i].id;
	std::vector<int> tmp(max_id + 1 > 0 ? max_id + 1 : 0, -1);
	for (int i = 0; i < n_regs; ++i) if (regs[i].id >= 0) tmp[regs[i].id] = i;
	for (int i = 0; i < n_regs; ++i) {
		wm_reg1_t *r = &regs[i];
		r->id = i;
		if (r->parent == WM_PARENT_TMP_PRI) r->parent = i;
		else if (r->parent >= 0 && tmp[r->parent] >= 0) r->parent = tmp[r->parent];
		else r->parent = WM_PARENT_UNSET;
	}
	set_sam_pri(n_regs, regs);
}

// mm_select_sub (src/hit.c:255-272)
void select_sub(float pri_ratio, int min_diff, int best_n, std::vector<wm_reg1_t> &regs)
{
	const int n = (int)regs.size();
	if (!(pri_ratio > 0.0f && n > 0)) return;
	wm_reg1_t *r = regs.data();
	int k = 0, n_2nd = 0;
	for (int i = 0; i < n; ++i) {
		const int p = r[i].parent;
		if (p == i || r[i].inv) r[k++] = r[i];
		else if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
			if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re))
				r[k++] = r[i], ++n_2nd;
			else if (r[i].p) free(r[i].p);
		} else if (r[i].p) free(r[i].p);
	}
	if (k != n) sync_regs(k, r);
	regs.resize(k);
}

// mm_filter_regs (src/hit.c:274-293)
void filter_regs(const wm_mapopt_t *opt, int qlen, std::vector<wm_reg1_t> &regs)
{
	int k = 0;
	for (size_t i = 0; i < regs.size(); ++i) {
		wm_reg1_t *r = &regs[i];
		int flt = 0;
		if (!r->inv && !r->seg_split && r->cnt < opt->min_cnt) flt = 1;
		if (r->p) {
			if (r->mlen < opt->min_chain_score) flt = 1;
			else if (r->p->dp_max < opt->min_dp_max) flt = 1;
			else if (r->qs > qlen * opt->max_clip_ratio && qlen - r->qe > qlen * opt->max_clip_ratio) flt = 1;
			if (flt) free(r->p);
		}
		if (!flt) { if (k < (int)i) regs[k++] = regs[i]; else ++k; }
	}
	regs.resize(k);
}

// mm_squeeze_a (src/hit.c:295-313)
int squeeze_a(std::vector<wm_reg1_t> &regs, wm_pair_t *a)
{
	const int n_regs = (int)regs.size();
	int as = 0;
	std::vector<uint64_t> aux(n_regs);
	for (int i = 0; i < n_regs; ++i) aux[i] = (uint64_t)regs[i].as << 32 | (uint32_t)i;
	radix_sort(aux.data(), aux.data() + n_regs);
	for (int i = 0; i < n_regs; ++i) {
		wm_reg1_t *r = &regs[(int32_t)aux[i]];
		if (r->as != as) {
			memmove(&a[as], &a[r->as], (size_t)r->cnt * 16);
			r->as = as;
		}
		as += r->cnt;
	}
	return as;
}

// mm_join_long (src/hit.c:315-371)
void join_long(const wm_mapopt_t *opt, int qlen, std::vector<wm_reg1_t> &regs, wm_pair_t *a)
{
	const int n_regs = (int)regs.size();
	if (n_regs < 2) return;
	squeeze_a(regs, a);
	std::vector<uint64_t> aux;
	for (int i = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0) aux.push_back((uint64_t)regs[i].as << 32 | (uint32_t)i);
	const int n_aux = (int)aux.size();
	radix_sort(aux.data(), aux.data() + n_aux);
	int n_drop = 0;
	for (int i = n_aux - 1; i >= 1; --i) {
		wm_reg1_t *r0 = &regs[(int32_t)aux[i - 1]], *r1 = &regs[(int32_t)aux[i]];
		if (r0->as + r0->cnt != r1->as) continue;
		if (r0->rid != r1->rid || r0->rev != r1->rev) continue;
		const wm_pair_t *a0e = &a[r0->as + r0->cnt - 1], *a1s = &a[r1->as];
		if (a1s->x <= a0e->x || (int32_t)a1s->y <= (int32_t)a0e->y) continue;
		int max_gap, min_gap;
		max_gap = min_gap = (int32_t)a1s->y - (int32_t)a0e->y;
		max_gap = a0e->x + max_gap > a1s->x ? max_gap : (int)(a1s->x - a0e->x);
		min_gap = a0e->x + min_gap < a1s->x ? min_gap : (int)(a1s->x - a0e->x);
		if (max_gap > opt->max_join_long || min_gap > opt->max_join_short) continue;
		const int sc_thres = (int)((float)opt->min_join_flank_sc / opt->max_join_long * max_gap + .499);
		if (r0->score < sc_thres || r1->score < sc_thres) continue;
		const int min_flank_len = (int)(max_gap * opt->min_join_flank_ratio);
		if (r0->re - r0->rs < min_flank_len || r0->qe - r0->qs < min_flank_len) continue;
		if (r1->re - r1->rs < min_flank_len || r1->qe - r1->qs < min_flank_len) continue;
		a[r1->as].y |= WM_SEED_LONG_JOIN;
		r0->cnt += r1->cnt, r0->score += r1->score;
		set_coor(r0, qlen, a);
		r1->cnt = 0;
		r1->parent = r0->id;
		++n_drop;
	}
	if (n_drop > 0) {
		for (int i = 0; i < n_regs; ++i) {
			wm_reg1_t *r = &regs[i];
			if (r->parent >= 0 && r->id != r->parent)
				if (regs[r->parent].parent >= 0 && regs[r->parent].parent != r->parent)
					r->parent = regs[r->parent].parent;
		}
		filter_regs(opt, qlen, regs);
		sync_regs((int)regs.size(), regs.data());
	}
}

// mm_hit_sort (src/hit.c:188-218)
void hit_sort(std::vector<wm_reg1_t> &regs, float alt_diff_frac)
{
	const int n = (int)regs.size();
	if (n <= 1) return;
	std::vector<wm_pair_t> aux;
	aux.reserve(n);
	for (int i = 0; i < n; ++i) {
		wm_reg1_t &r = regs[i];
		if (r.inv || r.cnt > 0) {
			int score = r.p ? r.p->dp_max : r.score;
			if (r.is_alt) score = alt_score(score, alt_diff_frac);
			wm_pair_t t;
			t.x = (uint64_t)score << 32 | r.hash;
			t.y = (uint64_t)i;
			aux.push_back(t);
		} else if (r.p) { free(r.p); r.p = 0; }
	}
	const int n_aux = (int)aux.size();
	radix_sort(aux.data(), aux.data() + n_aux);
	std::vector<wm_reg1_t> t(n_aux);
	for (int i = n_aux - 1; i >= 0; --i) t[n_aux - 1 - i] = regs[aux[i].y];
	regs.swap(t);
}

// chain_post (src/map.c:256-265) without the multi-segment branch
void chain_post(const wm_mapopt_t *opt, int k, int qlen, std::vector<wm_reg1_t> &regs, wm_pair_t *a)
{
	if (opt->flag & WM_F_ALL_CHAINS) return;
	set_parent(opt->mask_level, opt->mask_len, (int)regs.size(), regs.data(), opt->a * 2 + opt->b, (int)(opt->flag & WM_F_HARD_MLEVEL), opt->alt_drop);
	select_sub(opt->pri_ratio, k * 2, opt->best_n, regs);
	if (!(opt->flag & (WM_F_SPLICE | WM_F_SR | WM_F_NO_LJOIN))) join_long(opt, qlen, regs, a);
}

// mm_est_err (src/esterr.c:30-64): divergence estimate from the fraction of matched minimizers
static inline int32_t forward_qpos(int32_t qlen, const wm_pair_t *a)
{
	int32_t x = (int32_t)a->y;
	if (a->x >> 63) x = qlen - 1 - (x + 1 - span_of(*a));
	return x;
}

void est_err(const wm_host_idx *mi, int qlen, std::vector<wm_reg1_t> &regs, const wm_pair_t *a, int32_t n, const uint64_t *mini_pos)
{
	if (n == 0) return;
	uint64_t sum_k = 0;
	for (int i = 0; i < n; ++i) sum_k += mini_pos[i] >> 32 & 0xff;
	const float avg_k = (float)sum_k / n;
	for (size_t i = 0; i < regs.size(); ++i) {
		wm_reg1_t *r = &regs[i];
		r->div = -1.0f;
		if (r->cnt == 0) continue;
		int32_t st, en, L = 0, R = n - 1;
		{ // binary search of the first chained minimizer (get_mini_idx, esterr.c:15-28)
			const int32_t x = forward_qpos(qlen, r->rev ? &a[r->as + r->cnt - 1] : &a[r->as]);
			st = -1;
			while (L <= R) {
				int32_t m = (int32_t)(((uint64_t)L + R) >> 1), y = (int32_t)mini_pos[m];
				if (y < x) L = m + 1; else if (y > x) R = m - 1; else { st = m; break; }
			}
		}
		en = st;
		if (st < 0) continue; // "logic inconsistency" warning branch of the reference
		const int32_t l_ref = (int32_t)mi->len[r->rid];
		int32_t n_match = 1;
		for (int32_t k = 1, j = st + 1; j < n && k < r->cnt; ++j) {
			const int32_t x = forward_qpos(qlen, r->rev ? &a[r->as + r->cnt - 1 - k] : &a[r->as + k]);
			if (x == (int32_t)mini_pos[j]) ++k, en = j, ++n_match;
		}
		int32_t n_tot = en - st + 1;
		if (r->qs > avg_k && r->rs > avg_k) ++n_tot;
		if (qlen - r->qs > avg_k && l_ref - r->re > avg_k) ++n_tot;
		r->div = n_match >= n_tot ? 0.0f : (float)(1.0 - pow((double)n_match / n_tot, 1.0 / avg_k));
	}
}

// mm_set_mapq (src/hit.c:463-508) incl. mm_set_inv_mapq (:437-461)
void set_mapq(std::vector<wm_reg1_t> &regs_v, int min_chain_sc, int match_sc, int rep_len, int is_sr)
{
	static const float q_coef = 40.0f;
	const int n_regs = (int)regs_v.size();
	wm_reg1_t *regs = regs_v.data();
	if (n_regs == 0) return;
	int64_t sum_sc = 0;
	for (int i = 0; i < n_regs; ++i) if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
	const float uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
	for (int i = 0; i < n_regs; ++i) {
		wm_reg1_t *r = &regs[i];
		if (r->inv) r->mapq = 0;
		else if (r->parent == r->id) {
			int mapq, subsc;
			float pen_s1 = (r->score > 100 ? 1.0f : 0.01f * r->score) * uniq_ratio;
			float pen_cm = r->cnt > 10 ? 1.0f : 0.1f * r->cnt;
			pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
			subsc = r->subsc > min_chain_sc ? r->subsc : min_chain_sc;
			if (r->p && r->p->dp_max2 > 0 && r->p->dp_max > 0) {
				float identity = (float)r->mlen / r->blen;
				float x = (float)r->p->dp_max2 * subsc / r->p->dp_max / r->score0;
				mapq = (int)(identity * pen_cm * q_coef * (1.0f - x * x) * logf((float)r->p->dp_max / match_sc));
				if (!is_sr) {
					int mapq_alt = (int)(6.02f * identity * identity * (r->p->dp_max - r->p->dp_max2) / match_sc + .499f);
					mapq = mapq < mapq_alt ? mapq : mapq_alt;
				}
			} else {
				float x = (float)subsc / r->score0;
				if (r->p) {
					float identity = (float)r->mlen / r->blen;
					mapq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf((float)r->p->dp_max / match_sc));
				} else mapq = (int)(pen_cm * q_coef * (1.0f - x) * logf(r->score));
			}
			mapq -= (int)(4.343f * logf(r->n_sub + 1) + .499f);
			mapq = mapq > 0 ? mapq : 0;
			r->mapq = mapq < 60 ? mapq : 60;
			if (r->p && r->p->dp_max > r->p->dp_max2 && r->mapq == 0) r->mapq = 1;
		} else r->mapq = 0;
	}
	// inversion hits inherit the smaller MAPQ of their flanks
	if (n_regs < 3) return;
	bool any_inv = false;
	for (int i = 0; i < n_regs; ++i) any_inv |= regs[i].inv;
	if (!any_inv) return;
	std::vector<wm_pair_t> aux;
	for (int i = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0) {
			wm_pair_t t; t.y = i; t.x = (uint64_t)regs[i].rid << 32 | (uint32_t)regs[i].rs;
			aux.push_back(t);
		}
	const int n_aux = (int)aux.size();
	radix_sort(aux.data(), aux.data() + n_aux);
	for (int i = 1; i < n_aux - 1; ++i) {
		wm_reg1_t *inv = &regs[aux[i].y];
		if (inv->inv) {
			wm_reg1_t *l = &regs[aux[i - 1].y], *rr = &regs[aux[i + 1].y];
			inv->mapq = l->mapq < rr->mapq ? l->mapq : rr->mapq;
		}
	}
}

} // namespace wmh
