// Single-affine banded extension DP: ksw_extz2_sse (reference src/ksw2_extz2_sse.c:23-304), reached when
// the two gap pairs are equal (src/align.c:328-331).  Included by ksw_extd2.cu and by the CPU emulation harness of the tests.
//
// Same sweep as the dual-affine kernel (one warp per job, anti-diagonals r = i + j, lanes own consecutive target
// positions, the t - 1 operand comes from the neighbouring lane), same rotated direction matrix and traceback.  What
// differs is the arithmetic: the state rows hold UNSIGNED 8-bit offsets that start at 0 (kcalloc, :96), the cell score is
// z = s + 2(q + e) (:27), the three-way maximum compares `a` as signed bytes but takes `b` with an unsigned maximum
// (:40,:163-166), the clamp with mat[0] + 2(q + e) is unsigned (:41), and H accumulates u8 / v8 minus (q + e) (:273-287).
// Every operation here is the reference's 8-bit operation, wrap-around included, so out-of-range scoring sets misbehave
// identically -- down to the sign bytes that _mm_cvtsi32_si128(int8_t) ORs into cells 1..3 of a diagonal's first block (:153).
// One cell per lane and step: this path is rare (no preset uses a single gap pair), exactness comes first.
#pragma once
#include "ksw_extd2_common.cuh"

// S: the job's state slice, 9 bytes per target cell of tlen16 (u, v, x, y, s rows + int32 H)
__device__ void wm_extz2_fill_job(const wm_dp_job &J, const uint8_t *__restrict__ seq, uint8_t *__restrict__ bt,
                                  wm_extz_dev *out, const wm_dp_params &P, int8_t *S, int lane, unsigned long long *cell_ctr)
{
	unsigned long long cells_acc = 0;
	const unsigned FULL = 0xffffffffu;
	const uint8_t *query = seq + J.q_off, *target = seq + J.t_off;
	const int qlen = J.qlen, tlen = J.tlen, flag = J.flag;
	int w = J.w;
	wm_extz_dev ez;
	ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
	ez.max = 0, ez.score = ez.mqe = ez.mte = WM_NEG_INF;
	ez.n_cigar = 0, ez.zdropped = 0, ez.reach_end = 0, ez.reserved = 0;
	if (qlen <= 0 || tlen <= 0 || P.early_out) { if (lane == 0) *out = ez; return; }

	const int q = P.q, e = P.e, qe = q + e;
	const uint8_t qe2_u8 = (uint8_t)(qe * 2), max_sc_u8 = (uint8_t)(P.sc_mch + qe * 2), q_u8 = (uint8_t)q;
	const bool approx_max = (flag & 0x08) != 0, right = (flag & 0x02) != 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	const int tlen16 = (tlen + 15) / 16 * 16;
	const int n_col16 = wm_ncol16(qlen, tlen, w);
	uint8_t *u = (uint8_t*)S, *v = u + tlen16, *x = v + tlen16, *y = x + tlen16;
	int8_t *s = (int8_t*)(y + tlen16);
	int32_t *H = (int32_t*)(s + tlen16);
	for (int i = lane; i < tlen16; i += 32) {
		u[i] = v[i] = x[i] = y[i] = 0; s[i] = 0;
		if (!approx_max) H[i] = WM_NEG_INF;
	}
	__syncwarp();

	int32_t H0 = 0, last_H0_t = 0;
	int last_st = -1, last_en = -1;
	const int n_diag = qlen + tlen - 1;
	for (int r = 0; r < n_diag; ++r) {
		const int st0 = wm_band_st(r, qlen, w), en0 = wm_band_en(r, tlen, w);
		if (st0 > en0) { ez.zdropped = 1; break; }
		const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
		cells_acc += (unsigned long long)(en - st + 1);
		// boundary operands (:133-140); x1 / v1 are int8_t in the reference
		int8_t x1, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = (int8_t)x[st - 1], v1 = (int8_t)v[st - 1];
			else x1 = v1 = 0;
		} else x1 = 0, v1 = (int8_t)(r ? q : 0);
		if (en >= r && lane == 0) { y[r] = 0; u[r] = (uint8_t)(r ? q : 0); }
		{ // score refresh in unaligned 16-cell groups starting at st0 (:142-160)
			const int lim = st0 + ((en0 - st0) / 16 + 1) * 16;
			for (int i = st0 + lane; i < lim; i += 32) {
				if (i < tlen16) {
					const int sq = i < tlen ? target[i] : 0;
					const int j = r - i;
					const int sr = (j >= 0 && j < qlen) ? query[j] : 0;
					s[i] = (int8_t)((sq == 4 || sr == 4) ? P.sc_N : (sq == sr ? P.sc_mch : P.sc_mis));
				}
			}
		}
		__syncwarp();
		{
			uint8_t cx = (uint8_t)x1, cv = (uint8_t)v1;
			const uint8_t orx = x1 < 0 ? 0xff : 0, orv = v1 < 0 ? 0xff : 0;
			uint8_t *pr = bt + J.p_off + (size_t)r * n_col16;
			for (int c = st; c <= en; c += 32) {
				const int t = c + lane;
				const bool act = t <= en;
				uint8_t xo = 0, vo = 0, ut = 0, yo = 0; int8_t sc = 0;
				if (act) sc = s[t], xo = x[t], vo = v[t], ut = u[t], yo = y[t];
				unsigned xl = __shfl_up_sync(FULL, (unsigned)xo, 1), vl = __shfl_up_sync(FULL, (unsigned)vo, 1);
				if (lane == 0) xl = cx, vl = cv;
				cx = (uint8_t)__shfl_sync(FULL, (unsigned)xo, 31), cv = (uint8_t)__shfl_sync(FULL, (unsigned)vo, 31);
				if (act) {
					uint8_t xt1 = (uint8_t)xl, vt1 = (uint8_t)vl;
					if (t - st >= 1 && t - st <= 3) xt1 |= orx, vt1 |= orv;
					uint8_t z = (uint8_t)((uint8_t)sc + qe2_u8), a = (uint8_t)(xt1 + vt1), b = (uint8_t)(yo + ut), d;
					if (!right) { // :212-221
						d = (int8_t)a > (int8_t)z ? 1 : 0;
						z = (int8_t)z > (int8_t)a ? z : a;
						d = (int8_t)b > (int8_t)z ? 2 : d;
					} else { // :247-256
						d = (int8_t)z > (int8_t)a ? 0 : 1;
						z = (int8_t)z > (int8_t)a ? z : a;
						d = (int8_t)z > (int8_t)b ? d : 2;
					}
					z = z > b ? z : b;                 // unsigned maximum (:40)
					z = z < max_sc_u8 ? z : max_sc_u8; // unsigned clamp (:41)
					u[t] = (uint8_t)(z - vt1); v[t] = (uint8_t)(z - ut);
					const uint8_t zq = (uint8_t)(z - q_u8);
					a = (uint8_t)(a - zq); b = (uint8_t)(b - zq);
					if (!right) { // :223-229
						x[t] = (int8_t)a > 0 ? a : 0; d |= (int8_t)a > 0 ? 0x08 : 0;
						y[t] = (int8_t)b > 0 ? b : 0; d |= (int8_t)b > 0 ? 0x10 : 0;
					} else { // :258-264
						x[t] = 0 > (int8_t)a ? 0 : a; d |= 0 > (int8_t)a ? 0 : 0x08;
						y[t] = 0 > (int8_t)b ? 0 : b; d |= 0 > (int8_t)b ? 0 : 0x10;
					}
					pr[t - st] = d;
				}
			}
		}
		__syncwarp();
		if (!approx_max) { // exact max with the 32-bit H row (:267-321); u8 / v8 are unsigned bytes
			int32_t max_H, max_t;
			if (r > 0) {
				const int32_t Hm1 = en0 > 0 ? H[en0 - 1] : 0, Hen = H[en0];
				__syncwarp();
				const int en1 = st0 + (en0 - st0) / 4 * 4;
				long long best = (long long)0x8000000000000000LL;
				for (int t = st0 + lane; t < en0; t += 32) {
					const int32_t h = H[t] + (int32_t)v[t] - qe;
					H[t] = h;
					const uint32_t prio = t < en1 ? 1u + ((uint32_t)((t - st0) & 3) << 24) + (uint32_t)((t - st0) >> 2 << 2)
					                              : (1u << 27) + (uint32_t)(t - st0);
					const long long key = ((long long)h << 32) | (long long)(0xffffffffu - prio);
					best = key > best ? key : best;
				}
				const int32_t Hn = en0 > 0 ? Hm1 + (int32_t)u[en0] - qe : Hen + (int32_t)v[en0] - qe;
				if (lane == 0) H[en0] = Hn;
				{
					const long long key = ((long long)Hn << 32) | (long long)0xffffffffu;
					best = key > best ? key : best;
				}
				#pragma unroll
				for (int o = 16; o; o >>= 1) {
					const long long other = __shfl_xor_sync(FULL, best, o);
					best = other > best ? other : best;
				}
				max_H = (int32_t)(best >> 32);
				const uint32_t prio = 0xffffffffu - (uint32_t)(best & 0xffffffffLL);
				if (prio == 0) max_t = en0;
				else if (prio < (1u << 27)) max_t = st0 + (int)((prio - 1) & 0xffffffu) + (int)((prio - 1) >> 24);
				else max_t = st0 + (int)(prio - (1u << 27));
				__syncwarp();
			} else {
				max_H = (int32_t)v[0] - qe - qe, max_t = 0;
				if (lane == 0) H[0] = max_H;
				__syncwarp();
			}
			const int32_t Hen0 = H[en0], Hst0 = H[st0];
			if (en0 == tlen - 1 && Hen0 > ez.mte) ez.mte = Hen0, ez.mte_q = r - en;
			if (r - st0 == qlen - 1 && Hst0 > ez.mqe) ez.mqe = Hst0, ez.mqe_t = st0;
			if (wm_apply_zdrop(ez, max_H, r, max_t, J.zdrop, e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
		} else { // approximate max (:322-338)
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					const int32_t d0 = (int32_t)v[last_H0_t] - qe, d1 = (int32_t)u[last_H0_t + 1] - qe;
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += (int32_t)v[last_H0_t] - qe;
				} else {
					++last_H0_t, H0 += (int32_t)u[last_H0_t] - qe;
				}
				if ((flag & 0x10) && wm_apply_zdrop(ez, H0, r, last_H0_t, J.zdrop, e)) break;
			} else H0 = (int32_t)v[0] - qe - qe, last_H0_t = 0;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
		}
		last_st = st, last_en = en;
	}
	if (lane == 0) { *out = ez; if (cell_ctr) atomicAdd(cell_ctr, cells_acc); }
}
