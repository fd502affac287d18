// Splice-aware extension DP: ksw_exts2_sse (reference src/ksw2_exts2_sse.c:26-408), the call mm_align_pair makes when
// MM_F_SPLICE is set (src/align.c:326-327).  Included by ksw_extd2.cu and by the CPU emulation harness of the tests.
//
// Same sweep as the other two kernels of the family (one warp per job, anti-diagonals r = i + j, lanes own consecutive target
// positions, the t - 1 operand comes from the neighbouring lane, rotated direction matrix).  What differs: no band; one
// affine gap (q, e) plus a long-deletion state x2 that opens at q2, extends for free and pays the donor[] / acceptor[]
// site costs on the way in and out (:100-165: GT..AG / CT..AC by transcript strand, the GTr / yAG half bonus, optional
// junction annotation); direction bit 5 is the continuation of x2, state 3 reads as N_SKIP in the traceback
// (src/ksw2.h:141-145 with min_intron_len = long_thres); no end bonus.  All cell arithmetic is the reference's signed 8-bit
// arithmetic, wrap-around included.  One cell per lane and step, as in ksw_extz2.cuh: exactness first.
#pragma once
#include "ksw_extd2_common.cuh"

#define WM_EXTS2_CELL_BYTES 12 // u, v, x, y, x2, donor, acceptor, s rows + int32 H

// S: the job's state slice, WM_EXTS2_CELL_BYTES per target cell of tlen16.  junc: junction annotation of the target (one byte
// per target base, src/index.c:780) or null.
__device__ void wm_exts2_fill_job(const wm_dp_job &J, const uint8_t *__restrict__ seq, const uint8_t *__restrict__ junc, uint8_t *__restrict__ bt,
                                  wm_extz_dev *out, const wm_dp_params &P, int8_t *S, int lane, unsigned long long *cell_ctr)
{
	unsigned long long cells_acc = 0;
	const unsigned FULL = 0xffffffffu;
	const uint8_t *query = seq + J.q_off, *target = seq + J.t_off;
	const int qlen = J.qlen, tlen = J.tlen, flag = J.flag;
	wm_extz_dev ez;
	ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
	ez.max = 0, ez.score = ez.mqe = ez.mte = WM_NEG_INF;
	ez.n_cigar = 0, ez.zdropped = 0, ez.reach_end = 0, ez.reserved = 0;
	if (qlen <= 0 || tlen <= 0 || P.early_out) { if (lane == 0) *out = ez; return; }

	const int q = P.q, e = P.e, q2 = P.q2, qe = q + e, long_thres = P.long_thres, long_diff = P.long_diff;
	const bool approx_max = (flag & 0x08) != 0, right = (flag & 0x02) != 0, generic = (flag & 0x04) != 0;
	const int tlen16 = (tlen + 15) / 16 * 16;
	const int n_col16 = wm_ncol16(qlen, tlen, tlen > qlen ? tlen : qlen);
	int8_t *u = S, *v = u + tlen16, *x = v + tlen16, *y = x + tlen16, *x2 = y + tlen16;
	int8_t *donor = x2 + tlen16, *acceptor = donor + tlen16, *s = acceptor + tlen16;
	int32_t *H = (int32_t*)(s + tlen16);
	const bool fw = (flag & 0x100) != 0, rv = (flag & 0x200) != 0, sites = fw || rv;
	for (int i = lane; i < tlen16; i += 32) {
		u[i] = v[i] = x[i] = y[i] = (int8_t)(-q - e); x2[i] = (int8_t)(-q2); s[i] = 0;
		donor[i] = acceptor[i] = sites ? (int8_t)(-P.noncan) : 0;
		if (!approx_max) H[i] = WM_NEG_INF;
	}
	__syncwarp();
	if (sites) { // :106-165; the two orientations differ in which signal sits where
		const int semi_cost = (flag & 0x400) ? -P.noncan / 2 : 0;
		const bool revc = (flag & 0x80) != 0;
		const int d2 = revc ? 0 : 3, d3a = revc ? 1 : 0, d3b = revc ? 3 : 2;     // donor: G T r | G A y (reversed)
		const int a1 = revc ? 3 : 0, a2a_ = revc ? 0 : 1, a2b = revc ? 2 : 3;    // acceptor: y A G | r T G (reversed)
		const int jd_f = revc ? 2 : 1, jd_r = revc ? 4 : 8, ja_f = revc ? 1 : 2, ja_r = revc ? 8 : 4;
		for (int t = lane; t < tlen; t += 32) {
			if (t < tlen - 4) {
				int can = 0;
				if (fw && target[t + 1] == 2 && target[t + 2] == d2) can = 1;
				if (rv && target[t + 1] == 1 && target[t + 2] == d2) can = 1;
				if (can && (target[t + 3] == d3a || target[t + 3] == d3b)) can = 2;
				if (can) donor[t] = (int8_t)(can == 2 ? 0 : semi_cost);
			}
			if (junc && t < tlen - 1)
				if ((fw && (junc[t + 1] & jd_f)) || (rv && (junc[t + 1] & jd_r))) donor[t] = (int8_t)(donor[t] + P.junc_bonus);
			if (t >= 2) {
				int can = 0;
				if (fw && target[t - 1] == a1 && target[t] == 2) can = 1;
				if (rv && target[t - 1] == a1 && target[t] == 1) can = 1;
				if (can && (target[t - 2] == a2a_ || target[t - 2] == a2b)) can = 2;
				if (can) acceptor[t] = (int8_t)(can == 2 ? 0 : semi_cost);
			}
			if (junc)
				if ((fw && (junc[t] & ja_f)) || (rv && (junc[t] & ja_r))) acceptor[t] = (int8_t)(acceptor[t] + P.junc_bonus);
		}
		__syncwarp();
	}

	int32_t H0 = 0, last_H0_t = 0;
	int last_st = -1, last_en = -1;
	const int n_diag = qlen + tlen - 1;
	for (int r = 0; r < n_diag; ++r) {
		int st0 = 0, en0 = tlen - 1;
		if (st0 < r - qlen + 1) st0 = r - qlen + 1;
		if (en0 > r) en0 = r;
		const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
		cells_acc += (unsigned long long)(en - st + 1);
		// boundary operands (:178-191)
		const int8_t edge = (int8_t)(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : 0);
		int8_t x1, x21, v1;
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = (int8_t)(-q - e), x21 = (int8_t)(-q2), v1 = (int8_t)(-q - e);
		} else x1 = (int8_t)(-q - e), x21 = (int8_t)(-q2), v1 = edge;
		if (en >= r && lane == 0) { y[r] = (int8_t)(-q - e); u[r] = edge; }
		if (!generic) { // score refresh in unaligned 16-cell groups starting at st0 (:192-208)
			const int lim = st0 + ((en0 - st0) / 16 + 1) * 16;
			for (int i = st0 + lane; i < lim; i += 32)
				if (i < tlen16) {
					const int sq = i < tlen ? target[i] : 0;
					const int j = r - i;
					const int sr = (j >= 0 && j < qlen) ? query[j] : 0;
					s[i] = (int8_t)((sq == 4 || sr == 4) ? P.sc_N : (sq == sr ? P.sc_mch : P.sc_mis));
				}
		} else { // :209-211
			for (int i = st0 + lane; i <= en0; i += 32) s[i] = P.mat[target[i] * 5 + query[r - i]];
		}
		__syncwarp();
		{
			int8_t cx = x1, cx2 = x21, cv = v1;
			uint8_t *pr = bt + J.p_off + (size_t)r * n_col16;
			for (int c = st; c <= en; c += 32) {
				const int t = c + lane;
				const bool act = t <= en;
				int8_t xo = 0, x2o = 0, vo = 0, ut = 0, yo = 0, z = 0, dn = 0, ac = 0;
				if (act) z = s[t], xo = x[t], x2o = x2[t], vo = v[t], ut = u[t], yo = y[t], dn = donor[t], ac = acceptor[t];
				int xl = __shfl_up_sync(FULL, (int)xo, 1), x2l = __shfl_up_sync(FULL, (int)x2o, 1), vl = __shfl_up_sync(FULL, (int)vo, 1);
				if (lane == 0) xl = cx, x2l = cx2, vl = cv;
				cx = (int8_t)__shfl_sync(FULL, (int)xo, 31), cx2 = (int8_t)__shfl_sync(FULL, (int)x2o, 31), cv = (int8_t)__shfl_sync(FULL, (int)vo, 31);
				if (act) {
					const int8_t xt1 = (int8_t)xl, x2t1 = (int8_t)x2l, vt1 = (int8_t)vl;
					int8_t a = (int8_t)(xt1 + vt1), b = (int8_t)(yo + ut), a2 = (int8_t)(x2t1 + vt1);
					const int8_t a2a = (int8_t)(a2 + ac);
					uint8_t d;
					if (!right) { // :252-259
						d = a > z ? 1 : 0;    z = z > a ? z : a;
						d = b > z ? 2 : d;    z = z > b ? z : b;
						d = a2a > z ? 3 : d;  z = z > a2a ? z : a2a;
					} else { // :300-307
						d = z > a ? 0 : 1;    z = z > a ? z : a;
						d = z > b ? d : 2;    z = z > b ? z : b;
						d = z > a2a ? d : 3;  z = z > a2a ? z : a2a;
					}
					u[t] = (int8_t)(z - vt1); v[t] = (int8_t)(z - ut); // :51-56
					const int8_t zq = (int8_t)(z - q);
					a = (int8_t)(a - zq); b = (int8_t)(b - zq); a2 = (int8_t)(a2 - (int8_t)(z - q2));
					if (!right) { // :272-290
						x[t] = (int8_t)((a > 0 ? a : 0) - qe); d |= a > 0 ? 0x08 : 0;
						y[t] = (int8_t)((b > 0 ? b : 0) - qe); d |= b > 0 ? 0x10 : 0;
						x2[t] = (int8_t)((a2 > dn ? a2 : dn) - q2); d |= a2 > dn ? 0x20 : 0;
					} else { // :320-338
						x[t] = (int8_t)((0 > a ? 0 : a) - qe); d |= 0 > a ? 0 : 0x08;
						y[t] = (int8_t)((0 > b ? 0 : b) - qe); d |= 0 > b ? 0 : 0x10;
						x2[t] = (int8_t)((dn > a2 ? dn : a2) - q2); d |= dn > a2 ? 0 : 0x20;
					}
					pr[t - st] = d;
				}
			}
		}
		__syncwarp();
		if (!approx_max) { // exact max with the 32-bit H row (:341-389); the four-lane SIMD tie order as in ksw_extz2.cuh
			int32_t max_H, max_t;
			if (r > 0) {
				const int32_t Hm1 = en0 > 0 ? H[en0 - 1] : 0, Hen = H[en0];
				__syncwarp();
				const int en1 = st0 + (en0 - st0) / 4 * 4;
				long long best = (long long)0x8000000000000000LL;
				for (int t = st0 + lane; t < en0; t += 32) {
					const int32_t h = H[t] + (int32_t)v[t];
					H[t] = h;
					const uint32_t prio = t < en1 ? 1u + ((uint32_t)((t - st0) & 3) << 24) + (uint32_t)((t - st0) >> 2 << 2)
					                              : (1u << 27) + (uint32_t)(t - st0);
					const long long key = ((long long)h << 32) | (long long)(0xffffffffu - prio);
					best = key > best ? key : best;
				}
				const int32_t Hn = en0 > 0 ? Hm1 + (int32_t)u[en0] : Hen + (int32_t)v[en0];
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
				max_H = (int32_t)v[0] - qe, max_t = 0;
				if (lane == 0) H[0] = max_H;
				__syncwarp();
			}
			const int32_t Hen0 = H[en0], Hst0 = H[st0];
			if (en0 == tlen - 1 && Hen0 > ez.mte) ez.mte = Hen0, ez.mte_q = r - en;
			if (r - st0 == qlen - 1 && Hst0 > ez.mqe) ez.mqe = Hst0, ez.mqe_t = st0;
			if (wm_apply_zdrop(ez, max_H, r, max_t, J.zdrop, 0)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
		} else { // approximate max (:390-406)
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					const int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t];
				} else {
					++last_H0_t, H0 += u[last_H0_t];
				}
			} else H0 = (int32_t)v[0] - qe, last_H0_t = 0;
			if ((flag & 0x10) && wm_apply_zdrop(ez, H0, r, last_H0_t, J.zdrop, 0)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
		}
		last_st = st, last_en = en;
	}
	if (lane == 0) { *out = ez; if (cell_ctr) atomicAdd(cell_ctr, cells_acc); }
}
