// Local Smith-Waterman score with the tie rules of ksw_ll_i16 (reference src/ksw2_ll_sse.c:80-147, profile
// built by ksw_ll_qinit :32-78).  The striped SSE kernel evaluates the Gotoh recurrence over the query padded
// to a multiple of 8 (padding columns score 0 against everything, :70-75); the reported target end is the last
// row whose maximum is >= the running best (:138) and the query end is the last *striped* slot holding the
// best (:144-145).  One thread per job: these calls are rare and small (inversion tests on Z-drop regions).
#include <vector>
#include "wm_common.cuh"

struct wm_ll_job { int64_t q_off, t_off; int64_t s_off; int32_t qlen, tlen; };

__global__ void wm_ksw_ll_kernel(const wm_ll_job *__restrict__ jobs, int n, const uint8_t *__restrict__ seq, const int8_t *__restrict__ mat_g,
                                 int gapo, int gape, int32_t *__restrict__ scratch, int32_t *__restrict__ out)
{
	const int id = blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= n) return;
	int8_t mat[25];
	for (int i = 0; i < 25; ++i) mat[i] = mat_g[i];
	const wm_ll_job J = jobs[id];
	const uint8_t *query = seq + J.q_off, *target = seq + J.t_off;
	const int qlen = J.qlen, tlen = J.tlen;
	const int slen = (qlen + 7) / 8, qlen8 = slen * 8, gapoe = gapo + gape;
	int32_t *H0 = scratch + J.s_off, *H1 = H0 + qlen8, *E = H1 + qlen8, *Hmax = E + qlen8;
	for (int j = 0; j < qlen8; ++j) H0[j] = 0, E[j] = 0, Hmax[j] = 0;
	int gmax = 0, te = -1, qe = -1;
	for (int i = 0; i < tlen; ++i) {
		int f = 0, imax = 0, hd = 0;
		const int8_t *ma = mat + target[i] * 5;
		for (int j = 0; j < qlen8; ++j) {
			const int sc = j < qlen ? ma[query[j]] : 0;
			int e = E[j], h = hd + sc, t;
			hd = H0[j];
			if (h > 32767) h = 32767;
			if (h < e) h = e;
			if (h < f) h = f;
			H1[j] = h;
			if (h > imax) imax = h;
			t = h - gapoe; if (t < 0) t = 0;
			e -= gape; if (e < 0) e = 0;
			E[j] = e > t ? e : t;
			f -= gape; if (f < 0) f = 0;
			if (f < t) f = t;
		}
		if (imax >= gmax) {
			gmax = imax; te = i;
			for (int j = 0; j < qlen8; ++j) Hmax[j] = H1[j];
		}
		int32_t *tmp = H1; H1 = H0; H0 = tmp;
	}
	for (int i = 0; i < qlen8; ++i) { // striped slot i holds column i/8 + (i%8)*slen
		const int col = i / 8 + i % 8 * slen;
		if (Hmax[col] == gmax) qe = col;
	}
	out[3 * id] = gmax, out[3 * id + 1] = qe, out[3 * id + 2] = te;
}

// jobs/seq device resident; out: 3 ints per job (score, qe, te)
void wm_ksw_ll_launch(const wm_ll_job *d_jobs, int n, const uint8_t *d_seq, const int8_t *d_mat, int gapo, int gape, int32_t *d_scratch, int32_t *d_out, cudaStream_t st)
{
	if (n <= 0) return;
	wm_count_launch(); wm_ksw_ll_kernel<<<(n + 63) / 64, 64, 0, st>>>(d_jobs, n, d_seq, d_mat, gapo, gape, d_scratch, d_out);
	WM_CUDA_CHECK(cudaGetLastError());
}

extern "C" int wm_ksw_ll_batch(int n, const uint8_t *qseq, const int64_t *qoff, const uint8_t *tseq, const int64_t *toff,
                               const int8_t *mat, int gapo, int gape, int32_t *score, int32_t *qe, int32_t *te)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
		fprintf(stderr, "[ERROR] wm_ksw_ll_batch: no CUDA device visible; winnowmap-b200 has no CPU fallback\n");
		exit(1);
	}
	if (n <= 0) return 0;
	const int64_t qtot = qoff[n], ttot = toff[n];
	std::vector<wm_ll_job> jobs(n);
	int64_t s_off = 0;
	for (int i = 0; i < n; ++i) {
		jobs[i].q_off = qoff[i], jobs[i].t_off = qtot + toff[i];
		jobs[i].qlen = (int32_t)(qoff[i + 1] - qoff[i]), jobs[i].tlen = (int32_t)(toff[i + 1] - toff[i]);
		jobs[i].s_off = s_off;
		s_off += 4 * (int64_t)((jobs[i].qlen + 7) / 8 * 8);
	}
	uint8_t *d_seq = wm_dev_alloc<uint8_t>(qtot + ttot + 16);
	wm_ll_job *d_jobs = wm_dev_alloc<wm_ll_job>(n);
	int32_t *d_scr = wm_dev_alloc<int32_t>(s_off + 4), *d_out = wm_dev_alloc<int32_t>(3 * (size_t)n);
	int8_t *d_mat = wm_dev_alloc<int8_t>(32);
	WM_CUDA_CHECK(cudaMemcpy(d_seq, qseq, qtot, cudaMemcpyHostToDevice));
	WM_CUDA_CHECK(cudaMemcpy(d_seq + qtot, tseq, ttot, cudaMemcpyHostToDevice));
	WM_CUDA_CHECK(cudaMemcpy(d_jobs, jobs.data(), sizeof(wm_ll_job) * n, cudaMemcpyHostToDevice));
	WM_CUDA_CHECK(cudaMemcpy(d_mat, mat, 25, cudaMemcpyHostToDevice));
	wm_ksw_ll_launch(d_jobs, n, d_seq, d_mat, gapo, gape, d_scr, d_out, 0);
	std::vector<int32_t> out(3 * (size_t)n);
	WM_CUDA_CHECK(cudaMemcpy(out.data(), d_out, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost));
	for (int i = 0; i < n; ++i) score[i] = out[3 * i], qe[i] = out[3 * i + 1], te[i] = out[3 * i + 2];
	cudaFree(d_seq); cudaFree(d_jobs); cudaFree(d_scr); cudaFree(d_out); cudaFree(d_mat);
	return 0;
}
