// C-ABI kernel-level batch entry points (include/winnowmap_b200.h): host buffers in, host
// buffers out.  These are what the parity tests drive; the mapping pipeline uses the same
// device launchers on device-resident data.
#include <vector>
#include <string.h>
#include "wm_common.cuh"
#include "../../include/winnowmap_b200.h"

// wm_extd2_ws / wm_extd2_plan / wm_extd2_launch: wm_common.cuh

extern "C" const char *wm_version(void) { return "winnowmap-b200 0.1 (Winnowmap 2.03 semantics)"; }

extern "C" int wm_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" int wm_set_device(int device) { WM_CUDA_CHECK(cudaSetDevice(device)); return 0; }

static void wm_require_device(const char *who)
{
	if (wm_device_count() <= 0) {
		fprintf(stderr, "[ERROR] %s: no CUDA device visible; winnowmap-b200 has no CPU fallback\n", who);
		exit(1);
	}
}

extern "C" int wm_ksw_extd2_batch(int n, const uint8_t *qseq, const int64_t *qoff, const uint8_t *tseq, const int64_t *toff,
                                  const int8_t *mat, int q, int e, int q2, int e2,
                                  const int32_t *w, const int32_t *zdrop, const int32_t *end_bonus, const int32_t *flag,
                                  wm_extz_t *ez, uint32_t *cigar, const int64_t *cigar_off)
{
	wm_require_device("wm_ksw_extd2_batch");
	if (n <= 0) return 0;
	static_assert(sizeof(wm_extz_t) == sizeof(wm_extz_dev), "ez layout");
	// the sequence pool in the layout of the mapping pipeline (gpu_backend.cu): every sequence on a 16-byte boundary, zero padded
	// (the fill kernel stages aligned sequences with bulk copies; odd-numbered jobs are shifted by one byte here so that the
	// plain staging loop is exercised as well)
	std::vector<wm_dp_job> jobs(n);
	std::vector<uint8_t> pool;
	int64_t p_off = 0;
	for (int i = 0; i < n; ++i) {
		wm_dp_job &J = jobs[i];
		J.qlen = (int32_t)(qoff[i + 1] - qoff[i]); J.tlen = (int32_t)(toff[i + 1] - toff[i]);
		const size_t shift = (i & 1) ? 1 : 0;
		pool.resize((pool.size() + 15) / 16 * 16 + shift, 0);
		J.q_off = (int64_t)pool.size(); pool.insert(pool.end(), qseq + qoff[i], qseq + qoff[i + 1]);
		pool.resize((pool.size() + 15) / 16 * 16 + shift, 0);
		J.t_off = (int64_t)pool.size(); pool.insert(pool.end(), tseq + toff[i], tseq + toff[i + 1]);
		J.w = w[i]; J.zdrop = zdrop[i]; J.end_bonus = end_bonus[i]; J.flag = flag[i];
		J.p_off = p_off; p_off += (int64_t)wm_extd2_bt_bytes(J.qlen, J.tlen, J.w);
		J.cig_off = cigar_off[i]; J.cig_cap = (int32_t)(cigar_off[i + 1] - cigar_off[i]); J.pad = -1;
	}
	pool.resize((pool.size() + 15) / 16 * 16 + 32, 0);
	const wm_extd2_plan_t plan = wm_extd2_plan(jobs.data(), n, q == q2 && e == e2);
	std::vector<int32_t> coop; // the big jobs go to the CTA-cooperative sweep, as in the mapping pipeline (gpu_backend.cu)
	if (!(q == q2 && e == e2) && !(getenv("WM_DP_COOP") && *getenv("WM_DP_COOP") == '0') && !(getenv("WM_DP_V1") && *getenv("WM_DP_V1") == '1'))
		for (int i = 0; i < n; ++i)
			if (jobs[i].pad >= 0 && wm_dp_is_coop(jobs[i].qlen, jobs[i].tlen, jobs[i].w)) { jobs[i].flag |= WM_DP_COOP; coop.push_back(i); }
	int32_t *d_coop = wm_dev_alloc<int32_t>(coop.size() + 1);
	if (!coop.empty()) WM_CUDA_CHECK(cudaMemcpy(d_coop, coop.data(), sizeof(int32_t) * coop.size(), cudaMemcpyHostToDevice));
	uint8_t *d_seq = wm_dev_alloc<uint8_t>(pool.size() + 16);
	uint8_t *d_bt = wm_dev_alloc<uint8_t>(p_off + 16);
	wm_dp_job *d_jobs = wm_dev_alloc<wm_dp_job>(n);
	wm_extz_dev *d_ez = wm_dev_alloc<wm_extz_dev>(n);
	uint32_t *d_cig = wm_dev_alloc<uint32_t>(cigar_off[n] + 1);
	WM_CUDA_CHECK(cudaMemcpy(d_seq, pool.data(), pool.size(), cudaMemcpyHostToDevice));
	WM_CUDA_CHECK(cudaMemcpy(d_jobs, jobs.data(), sizeof(wm_dp_job) * n, cudaMemcpyHostToDevice));
	wm_dp_params P; wm_dp_params_init(&P, mat, q, e, q2, e2);
	wm_extd2_ws ws;
	wm_extd2_launch(&ws, d_jobs, n, plan, d_seq, d_bt, d_ez, d_cig, P, 0, 0, 0, d_coop, (int)coop.size());
	WM_CUDA_CHECK(cudaDeviceSynchronize());
	WM_CUDA_CHECK(cudaMemcpy(ez, d_ez, sizeof(wm_extz_dev) * n, cudaMemcpyDeviceToHost));
	if (cigar_off[n] > 0) WM_CUDA_CHECK(cudaMemcpy(cigar, d_cig, sizeof(uint32_t) * cigar_off[n], cudaMemcpyDeviceToHost));
	cudaFree(d_seq); cudaFree(d_bt); cudaFree(d_jobs); cudaFree(d_ez); cudaFree(d_cig); cudaFree(d_coop);
	return 0;
}

// Batched ksw_exts2_sse (src/ksw2_exts2_sse.c:26): the splice-aware extension.  flag carries KSW_EZ_SPLICE_FOR / _REV / _FLANK
// (0x100 / 0x200 / 0x400) beside the usual bits; junc: one annotation byte per target base (src/index.c:780), or null.
extern "C" int wm_ksw_exts2_batch(int n, const uint8_t *qseq, const int64_t *qoff, const uint8_t *tseq, const int64_t *toff, const uint8_t *junc,
                                  const int8_t *mat, int q, int e, int q2, int noncan, int junc_bonus,
                                  const int32_t *zdrop, const int32_t *flag, wm_extz_t *ez, uint32_t *cigar, const int64_t *cigar_off)
{
	wm_require_device("wm_ksw_exts2_batch");
	if (n <= 0) return 0;
	std::vector<wm_dp_job> jobs(n);
	std::vector<uint8_t> pool, jpool;
	int64_t p_off = 0;
	for (int i = 0; i < n; ++i) {
		wm_dp_job &J = jobs[i];
		J.qlen = (int32_t)(qoff[i + 1] - qoff[i]); J.tlen = (int32_t)(toff[i + 1] - toff[i]);
		pool.resize((pool.size() + 15) / 16 * 16, 0);
		J.q_off = (int64_t)pool.size(); pool.insert(pool.end(), qseq + qoff[i], qseq + qoff[i + 1]);
		pool.resize((pool.size() + 15) / 16 * 16, 0);
		J.t_off = (int64_t)pool.size(); pool.insert(pool.end(), tseq + toff[i], tseq + toff[i + 1]);
		if (junc) { jpool.resize(J.t_off, 0); jpool.insert(jpool.end(), junc + toff[i], junc + toff[i + 1]); }
		J.w = -1; J.zdrop = zdrop[i]; J.end_bonus = -1; J.flag = flag[i]; // no band, no end bonus on this path
		J.p_off = p_off; p_off += (int64_t)wm_extd2_bt_bytes(J.qlen, J.tlen, -1);
		J.cig_off = cigar_off[i]; J.cig_cap = (int32_t)(cigar_off[i + 1] - cigar_off[i]); J.pad = -1;
	}
	pool.resize((pool.size() + 15) / 16 * 16 + 32, 0);
	if (junc) jpool.resize(pool.size(), 0);
	const wm_extd2_plan_t plan = wm_extd2_plan(jobs.data(), n, false, true);
	uint8_t *d_seq = wm_dev_alloc<uint8_t>(pool.size() + 16), *d_junc = junc ? wm_dev_alloc<uint8_t>(pool.size() + 16) : 0;
	uint8_t *d_bt = wm_dev_alloc<uint8_t>(p_off + 16);
	wm_dp_job *d_jobs = wm_dev_alloc<wm_dp_job>(n);
	wm_extz_dev *d_ez = wm_dev_alloc<wm_extz_dev>(n);
	uint32_t *d_cig = wm_dev_alloc<uint32_t>(cigar_off[n] + 1);
	WM_CUDA_CHECK(cudaMemcpy(d_seq, pool.data(), pool.size(), cudaMemcpyHostToDevice));
	if (junc) WM_CUDA_CHECK(cudaMemcpy(d_junc, jpool.data(), jpool.size(), cudaMemcpyHostToDevice));
	WM_CUDA_CHECK(cudaMemcpy(d_jobs, jobs.data(), sizeof(wm_dp_job) * n, cudaMemcpyHostToDevice));
	wm_dp_params P; wm_dp_params_init_splice(&P, mat, q, e, q2, noncan, junc_bonus);
	wm_extd2_ws ws;
	wm_extd2_launch(&ws, d_jobs, n, plan, d_seq, d_bt, d_ez, d_cig, P, 0, 0, 0, 0, 0, d_junc);
	WM_CUDA_CHECK(cudaDeviceSynchronize());
	WM_CUDA_CHECK(cudaMemcpy(ez, d_ez, sizeof(wm_extz_dev) * n, cudaMemcpyDeviceToHost));
	if (cigar_off[n] > 0) WM_CUDA_CHECK(cudaMemcpy(cigar, d_cig, sizeof(uint32_t) * cigar_off[n], cudaMemcpyDeviceToHost));
	cudaFree(d_seq); cudaFree(d_junc); cudaFree(d_bt); cudaFree(d_jobs); cudaFree(d_ez); cudaFree(d_cig);
	return 0;
}
