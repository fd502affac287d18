// TEST INFRASTRUCTURE.  The DP fill sweep (csrc/ksw_extd2_v2.cuh) and the traceback (csrc/ksw_extd2_common.cuh) of
// the product, compiled for the host on top of cuda_emul.h: one job at a time on a software warp of 32 threads.  The
// tests compare the result with the CPU oracle -- the kernels' arithmetic is checked without a GPU; the real kernels
// are checked on the device by the -m gpu tests.
#include <stdlib.h>
#include <vector>
#include "cuda_emul.h"
#include "../../winnowmap_b200/csrc/wm_common.cuh"
#include "../../winnowmap_b200/csrc/ksw_extd2_common.cuh"
#include "../../winnowmap_b200/csrc/ksw_extd2_v2.cuh"

namespace wm_emul {
thread_local Warp *warp = 0; thread_local int lane = 0;
static void fiber_entry()
{
	Warp *w = warp;
	const int me = w->cur;
	lane = me;
	w->body(me, w->arg);
	w->done[me] = true;
	// a finished lane never comes back: continue with another unfinished lane, or return to the caller
	for (int k = 1; k <= 32; ++k) {
		const int nxt = (me + k) & 31;
		if (!w->done[nxt]) { w->cur = nxt; setcontext(&w->ctx[nxt]); }
	}
	setcontext(&w->main_ctx);
}
void run_warp(void (*body)(int lane, void *arg), void *arg)
{
	Warp W;
	memset(W.done, 0, sizeof(W.done));
	W.cur = 0, W.arrived = 0, W.gen = 0, W.body = body, W.arg = arg;
	const size_t stack_bytes = (size_t)1 << 18;
	for (int l = 0; l < 32; ++l) {
		W.stack[l] = (char*)malloc(stack_bytes);
		getcontext(&W.ctx[l]);
		W.ctx[l].uc_stack.ss_sp = W.stack[l], W.ctx[l].uc_stack.ss_size = stack_bytes, W.ctx[l].uc_link = 0;
		makecontext(&W.ctx[l], (void (*)())fiber_entry, 0);
	}
	Warp *saved = warp;
	warp = &W;
	swapcontext(&W.main_ctx, &W.ctx[0]);
	warp = saved;
	for (int l = 0; l < 32; ++l) free(W.stack[l]);
}
}
// defined by the CUDA library (ksw_extd2.cu); not used by the sweep itself
wm_prof_t g_wm_prof;
thread_local cudaStream_t wm_dbuf_stream = 0;
thread_local bool wm_dbuf_async = false;

// Same arithmetic as wm_dp_params_init / wm_extd2_bt_bytes of csrc/ksw_extd2.cu (host code of the CUDA library, not
// linkable here); src/ksw2_extd2_sse.c:61-97 and :84-86,114.
static void params_init(wm_dp_params *P, const int8_t *mat, int q, int e, int q2, int e2)
{
	P->qe_h = q + e;
	if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
	P->q = q, P->e = e, P->q2 = q2, P->e2 = e2;
	P->sc_mch = mat[0], P->sc_mis = mat[1];
	P->sc_N = mat[24] == 0 ? -e2 : mat[24];
	int min_sc = mat[1];
	for (int t = 1; t < 25; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
	P->early_out = -min_sc > 2 * (q + e);
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	P->long_thres = long_thres;
	P->long_diff = long_thres * (e - e2) - (q2 - q) - e2;
}

// one ksw_extd2 call through the product's sweep + traceback; global_state != 0 runs the instantiation that keeps
// the state rows in a global slice (the path of jobs that do not fit shared memory)
extern "C" int wmt_emul_extd2(const uint8_t *query, int qlen, const uint8_t *target, int tlen, const int8_t *mat, int q, int e, int q2, int e2,
                              int w, int zdrop, int end_bonus, int flag, int global_state, int32_t *ez_out, uint32_t *cigar, int cig_cap, int32_t *zd_out)
{
	wm_dp_params P; params_init(&P, mat, q, e, q2, e2);
	std::vector<uint8_t> seq((size_t)qlen + tlen + 64, 0);
	if (qlen > 0) memcpy(seq.data(), query, qlen);
	if (tlen > 0) memcpy(seq.data() + qlen, target, tlen);
	wm_dp_job J; memset(&J, 0, sizeof(J));
	J.q_off = 0, J.t_off = qlen, J.p_off = 0, J.cig_off = 0;
	J.qlen = qlen, J.tlen = tlen, J.w = w, J.zdrop = zdrop, J.end_bonus = end_bonus, J.flag = flag, J.cig_cap = cig_cap, J.pad = global_state ? 0 : -1;
	wm_extz_dev ez; memset(&ez, 0, sizeof(ez));
	const int tlen16 = (tlen + 15) / 16 * 16;
	if (!global_state && (tlen16 > WM_V2_T || qlen > WM_V2_Q)) return -1;
	if (qlen <= 0 || tlen <= 0 || P.early_out) { // what wm_extd2_fill_kernel's first-generation branch returns for these
		ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
		ez.max = 0, ez.score = ez.mqe = ez.mte = WM_NEG_INF;
	} else {
		int ww = w < 0 ? (tlen > qlen ? tlen : qlen) : w;
		const size_t bt_bytes = ((size_t)(qlen + tlen - 1) * (size_t)(wm_ncol16(qlen, tlen, ww) / 16) + 1) * 16;
		std::vector<uint8_t> bt(bt_bytes + 64, 0);
		const size_t slice = global_state ? wm_v2_slice_bytes(tlen16, qlen) : (size_t)WM_V2_SLICE + 64;
		std::vector<uint64_t> state(slice / 8 + 16, 0); // 8-byte aligned like the device slices
		struct Args { const wm_dp_job *J; const uint8_t *seq; uint8_t *bt; wm_extz_dev *ez; const wm_dp_params *P; uint8_t *state; int tlen16, qlen, global_state; }
			A = { &J, seq.data(), bt.data(), &ez, &P, (uint8_t*)state.data(), tlen16, qlen, global_state };
		wm_emul::run_warp([](int l, void *p) {
			Args &a = *(Args*)p;
			if (a.global_state) wm_extd2_fill_job_v2<false>(*a.J, a.seq, a.bt, a.ez, *a.P, a.state, a.tlen16, a.qlen, l, 0);
			else wm_extd2_fill_job_v2<true>(*a.J, a.seq, a.bt, a.ez, *a.P, a.state, 0, 0, l, 0);
		}, &A);
		wm_zd_params Z; memset(&Z, 0, sizeof(Z));
		Z.q = q, Z.e = e; memcpy(Z.mat, mat, 25);
		wm_extd2_backtrack_job(J, &ez, bt.data(), cigar, seq.data(), Z, zd_out);
	}
	memcpy(ez_out, &ez, sizeof(ez));
	return 0;
}
