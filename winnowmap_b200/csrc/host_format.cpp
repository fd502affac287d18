// PAF output, byte-identical to the reference writer (src/format.c:266-334; tags :280-306).
#include <stdio.h>
#include <string.h>
#include "host_io.h"

namespace wmh {

static inline void put_int(std::string &s, int c)
{ // the %d of mm_sprintf_lite (src/format.c:36-44)
	char buf[16]; int l = 0;
	unsigned x = c >= 0 ? c : -c;
	do { buf[l++] = x % 10 + '0'; x /= 10; } while (x > 0);
	if (c < 0) buf[l++] = '-';
	while (l > 0) s.push_back(buf[--l]);
}

static double event_identity(const wm_reg1_t *r)
{ // mm_event_identity (src/format.c:267-278)
	int32_t n_gapo = 0, n_gap = 0;
	if (r->p == 0) return -1.0f;
	for (uint32_t i = 0; i < r->p->n_cigar; ++i) {
		int32_t op = r->p->cigar[i] & 0xf, len = r->p->cigar[i] >> 4;
		if (op == 1 || op == 2) ++n_gapo, n_gap += len;
	}
	return (double)r->mlen / (r->blen + r->p->n_ambi - n_gap + n_gapo);
}

static void write_tags(std::string &s, const wm_reg1_t *r)
{ // src/format.c:280-306
	int type;
	if (r->id == r->parent) type = r->inv ? 'I' : 'P';
	else type = r->inv ? 'i' : 'S';
	if (r->p) {
		s += "\tNM:i:"; put_int(s, r->blen - r->mlen + r->p->n_ambi);
		s += "\tms:i:"; put_int(s, r->p->dp_max);
		s += "\tAS:i:"; put_int(s, r->p->dp_score);
		s += "\tnn:i:"; put_int(s, r->p->n_ambi);
		if (r->p->trans_strand == 1 || r->p->trans_strand == 2) { s += "\tts:A:"; s.push_back("?+-?"[r->p->trans_strand]); }
	}
	s += "\ttp:A:"; s.push_back((char)type);
	s += "\tcm:i:"; put_int(s, r->cnt);
	s += "\ts1:i:"; put_int(s, r->score);
	if (r->parent == r->id) { s += "\ts2:i:"; put_int(s, r->subsc); }
	if (r->p) {
		char buf[16];
		double div = 1.0 - event_identity(r);
		if (div == 0.0) buf[0] = '0', buf[1] = 0;
		else snprintf(buf, 16, "%.4f", 1.0 - event_identity(r));
		s += "\tde:f:"; s += buf;
	} else if (r->div >= 0.0f && r->div <= 1.0f) {
		char buf[16];
		if (r->div == 0.0f) buf[0] = '0', buf[1] = 0;
		else snprintf(buf, 16, "%.4f", r->div);
		s += "\tdv:f:"; s += buf;
	}
	if (r->split) { s += "\tzd:i:"; put_int(s, r->split); }
}

void write_paf(std::string &s, const wm_host_idx *mi, const wm_read *t, const wm_reg1_t *r, int64_t opt_flag, int rep_len)
{
	s.clear();
	if (r == 0) {
		s += t->name; s.push_back('\t'); put_int(s, (int)t->seq.size());
		s += "\t0\t0\t*\t*\t0\t0\t0\t0\t0\t0";
		if (rep_len >= 0) { s += "\trl:i:"; put_int(s, rep_len); }
		return;
	}
	s += t->name; s.push_back('\t'); put_int(s, (int)t->seq.size()); s.push_back('\t'); put_int(s, r->qs); s.push_back('\t'); put_int(s, r->qe);
	s.push_back('\t'); s.push_back("+-"[r->rev]); s.push_back('\t');
	s += mi->name[r->rid];
	s.push_back('\t'); put_int(s, (int)mi->len[r->rid]); s.push_back('\t'); put_int(s, r->rs); s.push_back('\t'); put_int(s, r->re);
	s.push_back('\t'); put_int(s, r->mlen); s.push_back('\t'); put_int(s, r->blen);
	s.push_back('\t'); put_int(s, (int)r->mapq);
	write_tags(s, r);
	if (rep_len >= 0) { s += "\trl:i:"; put_int(s, rep_len); }
	if (r->p && (opt_flag & WM_F_OUT_CG)) {
		s += "\tcg:Z:";
		for (uint32_t k = 0; k < r->p->n_cigar; ++k) { put_int(s, (int)(r->p->cigar[k] >> 4)); s.push_back("MIDNSHP=XB"[r->p->cigar[k] & 0xf]); }
	}
	if ((opt_flag & WM_F_COPY_COMMENT) && !t->comment.empty()) { s.push_back('\t'); s += t->comment; }
}

} // namespace wmh
