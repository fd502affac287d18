// Anchor chaining on sm_100a: mm_chain_dp (reference src/chain.c:22-167) for n_segs == 1 and
// is_cdna == 0, one warp per task.
//
// Forward pass (:45-90): anchors are taken in order; for anchor i the 32 lanes score 32 predecessors
// j = i-1, i-2, ... at a time.  The reference's inner loop is order dependent (running maximum with
// strict ">", the n_skip counter fed by t[j]==i marks, "break" above max_skip, and the t[p[j]] = i side
// effects); a chunk is resolved exactly by (1) letting every candidate lane publish t[p[j]] = i, which
// can only touch indices below every j still to be examined, (2) a prefix maximum over the lanes to
// find the record-setting lanes, and (3) replaying the n_skip arithmetic over the ballot masks.
// Backtracking (:92-165) is a second warp-per-task kernel: the data-parallel sweeps use all lanes, the
// greedy claim walk (:118-135) is inherently serial and runs on lane 0.
#include <vector>
#include <algorithm>
#include <limits.h>
#include <stdlib.h>
#include "wm_common.cuh"
#include "sketch.cuh"
#include "rsort.cuh"
#include "chain.cuh"

#define WM_CHAIN_WARPS 4
#define WM_CHAIN_DENSE_MIN 1024 // tasks with more anchors use the dense-candidate scan when WM_CHAIN_DENSE=1

#include "chain_dev.cuh"

// Tasks come largest first (order[]), one warp per task.
__global__ void __launch_bounds__(WM_CHAIN_WARPS * 32)
wm_chain_fill_kernel(const wm128_dev *__restrict__ a_all, const int64_t *__restrict__ off, const int32_t *__restrict__ order, int n_tasks,
                     wm_chain_params2 PP, const uint8_t *__restrict__ set_id, int32_t *__restrict__ f_all, int32_t *__restrict__ p_all, int32_t *__restrict__ t_all, int32_t *__restrict__ v_all,
                     int *counter)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	for (;;) {
		int ti = 0;
		if (lane == 0) ti = atomicAdd(counter, 1);
		ti = __shfl_sync(FULL, ti, 0);
		if (ti >= n_tasks) break;
		const int task = order ? order[ti] : ti;
		const int64_t base = off[task];
		const int n = (int)(off[task + 1] - base);
		if (n <= 0) continue;
		wm_chain_fill_warp(a_all + base, n, PP.p[set_id ? set_id[task] : 0], f_all + base, p_all + base, t_all + base, v_all + base, lane);
	}
}

// The same forward pass with the dense-candidate scan of chain_dev.cuh (selected with WM_CHAIN_DENSE=1).
__global__ void __launch_bounds__(WM_CHAIN_WARPS * 32)
wm_chain_fill_dense_kernel(const wm128_dev *__restrict__ a_all, const int64_t *__restrict__ off, const int32_t *__restrict__ order, int n_tasks,
                           wm_chain_params2 PP, const uint8_t *__restrict__ set_id, int32_t *__restrict__ f_all, int32_t *__restrict__ p_all, int32_t *__restrict__ t_all,
                           int32_t *__restrict__ v_all, int *counter)
{
	const unsigned FULL = 0xffffffffu;
	__shared__ int32_t D[WM_CHAIN_WARPS][WM_CHAIN_DENSE_CAP];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	for (;;) {
		int ti = 0;
		if (lane == 0) ti = atomicAdd(counter, 1);
		ti = __shfl_sync(FULL, ti, 0);
		if (ti >= n_tasks) break;
		const int task = order ? order[ti] : ti;
		const int64_t base = off[task];
		const int n = (int)(off[task + 1] - base);
		if (n <= 0) continue;
		// the dense scan pays off where windows are long (big tasks: repeats); ordinary tasks keep the plain chunk loop
		if (n > WM_CHAIN_DENSE_MIN) wm_chain_fill_warp_dense(a_all + base, n, PP.p[set_id ? set_id[task] : 0], f_all + base, p_all + base, t_all + base, v_all + base, D[wid], lane);
		else wm_chain_fill_warp(a_all + base, n, PP.p[set_id ? set_id[task] : 0], f_all + base, p_all + base, t_all + base, v_all + base, lane);
	}
}

__global__ void __launch_bounds__(WM_CHAIN_WARPS * 32)
wm_chain_backtrack_kernel(wm128_dev *__restrict__ a_all, const int64_t *__restrict__ off, int n_tasks, wm_chain_params2 PP, const uint8_t *__restrict__ set_id,
                          int32_t *__restrict__ f_all, int32_t *__restrict__ p_all, int32_t *__restrict__ t_all, int32_t *__restrict__ v_all,
                          uint64_t *__restrict__ u_all, uint64_t *__restrict__ u2_all, wm128_dev *__restrict__ w_all, wm128_dev *__restrict__ b_all,
                          int32_t *__restrict__ n_u_out, int64_t *__restrict__ n_b_out, wm_rs_stack *__restrict__ stacks, int *counter)
{
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int wslot = blockIdx.x * WM_CHAIN_WARPS + (threadIdx.x >> 5);
	for (;;) {
		int task = 0;
		if (lane == 0) task = atomicAdd(counter, 1);
		task = __shfl_sync(FULL, task, 0);
		if (task >= n_tasks) break;
		const int64_t base = off[task];
		const int n = (int)(off[task + 1] - base);
		if (lane == 0) n_u_out[task] = 0, n_b_out[task] = 0;
		if (n <= 0) continue;
		__syncwarp();
		wm_chain_backtrack_warp(a_all + base, n, PP.p[set_id ? set_id[task] : 0], f_all + base, p_all + base, t_all + base, v_all + base,
		                        u_all + 2 * base, u2_all + base, w_all + base, b_all + base, stacks + wslot, n_u_out + task, n_b_out + task, lane);
	}
}

// Chains for n_tasks anchor arrays a[off[t]..off[t+1]) (device, sorted).  Results are left in place:
// d_a holds the chained anchors of task t at off[t].. (n_b[t] of them), ws->u2 the (score<<32|cnt) words
// at off[t].. (n_u[t] of them).
void wm_chain_run(wm_chain_ws *ws, wm128_dev *d_a, const int64_t *d_off, const int64_t *h_off, int n_tasks, const wm_chain_params2 &PP, const uint8_t *d_set_id, cudaStream_t st)
{
	if (n_tasks <= 0) return;
	const int64_t n_a = h_off[n_tasks];
	int32_t *f = (int32_t*)ws->f.need(sizeof(int32_t) * (n_a + 1)), *p = (int32_t*)ws->p.need(sizeof(int32_t) * (n_a + 1));
	int32_t *t = (int32_t*)ws->t.need(sizeof(int32_t) * (n_a + 1)), *v = (int32_t*)ws->v.need(sizeof(int32_t) * (n_a + 1));
	uint64_t *u = (uint64_t*)ws->u.need(sizeof(uint64_t) * (2 * n_a + 2)), *u2 = (uint64_t*)ws->u2.need(sizeof(uint64_t) * (n_a + 1));
	wm128_dev *w = (wm128_dev*)ws->w.need(sizeof(wm128_dev) * (n_a + 1)), *b = (wm128_dev*)ws->b.need(sizeof(wm128_dev) * (n_a + 1));
	int32_t *n_u = (int32_t*)ws->n_u.need(sizeof(int32_t) * (n_tasks + 1));
	int64_t *n_b = (int64_t*)ws->n_b.need(sizeof(int64_t) * (n_tasks + 1));
	int *counter = (int*)ws->counter.need(2 * sizeof(int));
	WM_CUDA_CHECK(cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));
	// largest tasks first
	std::vector<int32_t> order(n_tasks);
	for (int i = 0; i < n_tasks; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return h_off[x + 1] - h_off[x] > h_off[y + 1] - h_off[y]; });
	int32_t *d_order = (int32_t*)ws->order.need(sizeof(int32_t) * n_tasks);
	WM_CUDA_CHECK(wm_memcpy_async(d_order, order.data(), sizeof(int32_t) * n_tasks, cudaMemcpyHostToDevice, st));
	int dev = 0, n_sm = 148;
	WM_CUDA_CHECK(cudaGetDevice(&dev));
	WM_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
	int grid = n_sm * (32 / WM_CHAIN_WARPS);
	const int need = (n_tasks + WM_CHAIN_WARPS - 1) / WM_CHAIN_WARPS;
	if (grid > need) grid = need;
	wm_rs_stack *stk = (wm_rs_stack*)ws->stacks.need(sizeof(wm_rs_stack) * (size_t)grid * WM_CHAIN_WARPS);
	static int dense = -1; // experimental formulation, off unless WM_CHAIN_DENSE=1 (same results, see chain_dev.cuh)
	if (dense < 0) { const char *e = getenv("WM_CHAIN_DENSE"); dense = (e && *e == '1') ? 1 : 0; }
	wm_count_launch();
	if (dense) wm_chain_fill_dense_kernel<<<grid, WM_CHAIN_WARPS * 32, 0, st>>>(d_a, d_off, d_order, n_tasks, PP, d_set_id, f, p, t, v, counter);
	else wm_chain_fill_kernel<<<grid, WM_CHAIN_WARPS * 32, 0, st>>>(d_a, d_off, d_order, n_tasks, PP, d_set_id, f, p, t, v, counter);
	WM_CUDA_CHECK(cudaGetLastError());
	wm_count_launch(); wm_chain_backtrack_kernel<<<grid, WM_CHAIN_WARPS * 32, 0, st>>>(d_a, d_off, n_tasks, PP, d_set_id, f, p, t, v, u, u2, w, b, n_u, n_b, stk, counter + 1);
	WM_CUDA_CHECK(cudaGetLastError());
}

// ---- C ABI ----
extern "C" int wm_chain_dp_batch(int n_tasks, const wm128_dev *a, const int64_t *off,
                                 int max_dist_x, int min_dist_x, int max_dist_y, int bw, int max_skip, int max_iter,
                                 int min_cnt, int min_sc, float gap_scale,
                                 int32_t *n_u, uint64_t *u, wm128_dev *b, int64_t *n_b)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
		fprintf(stderr, "[ERROR] wm_chain_dp_batch: no CUDA device visible; winnowmap-b200 has no CPU fallback\n");
		exit(1);
	}
	if (n_tasks <= 0) return 0;
	const int64_t n = off[n_tasks];
	wm128_dev *d_a = wm_dev_alloc<wm128_dev>(n + 1);
	int64_t *d_off = wm_dev_alloc<int64_t>(n_tasks + 1);
	WM_CUDA_CHECK(cudaMemcpy(d_a, a, sizeof(wm128_dev) * n, cudaMemcpyHostToDevice));
	WM_CUDA_CHECK(cudaMemcpy(d_off, off, sizeof(int64_t) * (n_tasks + 1), cudaMemcpyHostToDevice));
	wm_chain_params2 PP;
	wm_chain_params &P = PP.p[0];
	P.max_dist_x = max_dist_x, P.min_dist_x = min_dist_x, P.max_dist_y = max_dist_y, P.bw = bw, P.max_skip = max_skip, P.max_iter = max_iter;
	P.min_cnt = min_cnt, P.min_sc = min_sc, P.gap_scale = gap_scale;
	PP.p[1] = P;
	wm_chain_ws ws;
	wm_chain_run(&ws, d_a, d_off, off, n_tasks, PP, 0, 0);
	WM_CUDA_CHECK(cudaDeviceSynchronize());
	WM_CUDA_CHECK(cudaMemcpy(n_u, ws.n_u.p, sizeof(int32_t) * n_tasks, cudaMemcpyDeviceToHost));
	WM_CUDA_CHECK(cudaMemcpy(n_b, ws.n_b.p, sizeof(int64_t) * n_tasks, cudaMemcpyDeviceToHost));
	if (n > 0) {
		WM_CUDA_CHECK(cudaMemcpy(u, ws.u2.p, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
		WM_CUDA_CHECK(cudaMemcpy(b, d_a, sizeof(wm128_dev) * n, cudaMemcpyDeviceToHost));
	}
	ws.release();
	cudaFree(d_a); cudaFree(d_off);
	return 0;
}
