// Pieces of the banded extension DP shared by the fill sweeps (ksw_extd2.cu, ksw_extd2_v2.cuh), the traceback
// kernel and the CPU emulation harness of the tests (tests/hostsim/kernel_emul.cpp compiles this header for the host).
#pragma once
#include <limits.h>
#include "wm_common.cuh"

__device__ __forceinline__ int wm_band_st(int r, int qlen, int w)
{
	int st = 0;
	if (st < r - qlen + 1) st = r - qlen + 1;
	if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
	return st;
}
__device__ __forceinline__ int wm_band_en(int r, int tlen, int w)
{
	int en = tlen - 1;
	if (en > r) en = r;
	if (en > ((r + w) >> 1)) en = (r + w) >> 1;
	return en;
}

__device__ __forceinline__ int wm_ncol16(int qlen, int tlen, int w)
{ // src/ksw2_extd2_sse.c:84-86, in bytes
	int n = qlen < tlen ? qlen : tlen;
	n = ((n < w + 1 ? n : w + 1) + 15) / 16 + 1;
	return n * 16;
}

// ksw_apply_zdrop (src/ksw2.h:160-176), rotated coordinates
__device__ __forceinline__ bool wm_apply_zdrop(wm_extz_dev &ez, int32_t H, int r, int t, int zdrop, int e)
{
	if (H > ez.max) {
		ez.max = H, ez.max_t = t, ez.max_q = r - t;
	} else if (t >= ez.max_t && r - t >= ez.max_q) {
		int tl = t - ez.max_t, ql = (r - t) - ez.max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && ez.max - H > zdrop + l * e) { ez.zdropped = 1; return true; }
	}
	return false;
}

// ksw_backtrack (src/ksw2.h:119-151, is_rot = 1) + the start-cell choice of
// src/ksw2_extd2_sse.c:379-391.  One thread per job; the band bounds off[]/off_end[] the
// reference stores per diagonal are recomputed from r.
// CIGAR run-length accumulator with ksw_push_cigar semantics (src/ksw2.h:103-113); the pending
// run stays in registers so that the count stays exact even if the buffer is too small.
struct wm_cigar_acc {
	uint32_t *cig; int cap, n, cur_len; uint32_t cur_op; bool pending;
	__device__ __forceinline__ void init(uint32_t *c, int cap_) { cig = c, cap = cap_, n = 0, cur_len = 0, cur_op = 0, pending = false; }
	__device__ __forceinline__ void flush() { if (pending) { if (n < cap) cig[n] = (uint32_t)cur_len << 4 | cur_op; ++n; pending = false; } }
	__device__ __forceinline__ void push(uint32_t op, int len) {
		if (pending && op == cur_op) cur_len += len;
		else { flush(); cur_op = op, cur_len = len, pending = true; }
	}
};

// The score walk of mm_test_zdrop (src/align.c:32-70) over a finished CIGAR, for the jobs that ask for it (flag
// WM_DP_SCAN_ZDROP: the gap fills, whose result the aligner tests before deciding on a second pass, :736).
// out[0] = max_zdrop, out[1..4] = the most-dropped region {t0, t1, q0, q1}.
__device__ void wm_zdrop_scan(const wm_zd_params &Z, const uint8_t *__restrict__ qseq, const uint8_t *__restrict__ tseq, int n_cigar,
                              const uint32_t *__restrict__ cigar, int32_t *out)
{
	int32_t score = 0, mx = INT_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
	int32_t p00 = -1, p01 = -1, p10 = -1, p11 = -1;
	for (int k = 0; k < n_cigar; ++k) {
		const uint32_t op = cigar[k] & 0xf; const int len = (int)(cigar[k] >> 4);
		int steps = 1;
		if (op == 0) steps = len;
		else if (op == 1 || op == 2 || op == 3) { score -= Z.q + Z.e * len; if (op == 1) j += len; else i += len; }
		else continue;
		for (int l = 0; l < steps; ++l) {
			int ii = i, jj = j;
			if (op == 0) { score += Z.mat[tseq[i + l] * 5 + qseq[j + l]]; ii = i + l, jj = j + l; }
			if (score < mx) { // update_max_zdrop :32-45
				const int li = ii - max_i, lj = jj - max_j;
				const int diff = li > lj ? li - lj : lj - li;
				const int z = mx - score - diff * Z.e;
				if (z > max_zdrop) max_zdrop = z, p00 = max_i, p01 = ii, p10 = max_j, p11 = jj;
			} else mx = score, max_i = ii, max_j = jj;
		}
		if (op == 0) i += len, j += len;
	}
	out[0] = max_zdrop, out[1] = p00, out[2] = p01, out[3] = p10, out[4] = p11;
}

// One job of the traceback kernel: *ez_io is the job's fill result (updated: reach_end, n_cigar), zd its five Z-drop words or null.
__device__ __forceinline__ void wm_extd2_backtrack_job(const wm_dp_job &J, wm_extz_dev *ez_io, const uint8_t *__restrict__ bt, uint32_t *__restrict__ cigar_pool,
                                              const uint8_t *__restrict__ seq, const wm_zd_params &Z, int32_t *zd, int splice = 0, int min_intron_len = 0, int early_out = 0)
{ // early_out: the scoring made the reference return before the sweep (wm_dp_params::early_out): no CIGAR.  splice: the job came from ksw_exts2 (no end-bonus clause, src/ksw2_exts2_sse.c:411-419; the long-gap state reads as N_SKIP)
	wm_extz_dev ez = *ez_io;
	const int qlen = J.qlen, tlen = J.tlen;
	int w = J.w;
	const bool scan = zd != 0 && (J.flag & WM_DP_SCAN_ZDROP) != 0;
	if (scan) zd[0] = -1; // "no result": the host falls back to its own walk
	if (qlen <= 0 || tlen <= 0 || early_out) return;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int i0 = -1, j0 = -1;
	if (!ez.zdropped && !(J.flag & 0x40)) i0 = tlen - 1, j0 = qlen - 1;
	else if (!splice && !ez.zdropped && (J.flag & 0x40) && ez.mqe + J.end_bonus > ez.max) ez.reach_end = 1, i0 = ez.mqe_t, j0 = qlen - 1;
	else if (ez.max_t >= 0 && ez.max_q >= 0) i0 = ez.max_t, j0 = ez.max_q;
	int n = 0;
	if (i0 >= 0 && j0 >= 0) {
		const int n_col16 = wm_ncol16(qlen, tlen, w);
		const uint8_t *p = bt + J.p_off;
		uint32_t *cig = cigar_pool + J.cig_off;
		const int cap = J.cig_cap;
		wm_cigar_acc acc; acc.init(cig, cap);
		int i = i0, j = j0, state = 0;
		while (i >= 0 && j >= 0) {
			int r = i + j, force_state = -1;
			int off = wm_band_st(r, qlen, w) / 16 * 16, off_end = (wm_band_en(r, tlen, w) + 16) / 16 * 16 - 1;
			if (i < off) force_state = 2;
			if (i > off_end) force_state = 1;
			uint32_t tmp = force_state < 0 ? p[(size_t)r * n_col16 + i - off] : 0;
			if (state == 0) state = tmp & 7;
			else if (!(tmp >> (state + 2) & 1)) state = 0;
			if (state == 0) state = tmp & 7;
			if (force_state >= 0) state = force_state;
			if (state == 0) acc.push(0, 1), --i, --j;
			else if (state == 1 || (state == 3 && min_intron_len <= 0)) acc.push(2, 1), --i;
			else if (state == 3) acc.push(3, 1), --i; // intron (src/ksw2.h:142)
			else acc.push(1, 1), --j;
		}
		if (i >= 0) acc.push(min_intron_len > 0 && i >= min_intron_len ? 3u : 2u, i + 1);
		if (j >= 0) acc.push(1, j + 1);
		acc.flush();
		n = acc.n;
		if (!(J.flag & 0x80) && n <= cap)
			for (int k = 0; k < n >> 1; ++k) { uint32_t t2 = cig[k]; cig[k] = cig[n - 1 - k]; cig[n - 1 - k] = t2; }
	}
	ez.n_cigar = n;
	*ez_io = ez;
	if (scan && n <= J.cig_cap) wm_zdrop_scan(Z, seq + J.q_off, seq + J.t_off, n, cigar_pool + J.cig_off, zd);
}

