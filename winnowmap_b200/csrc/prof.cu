// Bench instrumentation (see wm_common.cuh): per-launch CUDA event pairs for the kernel classes whose roofline
// bench.py reports.  Lanes launch concurrently, so besides the sum of the launch durations the time during which at
// least one kernel of a class was running (union of the launch intervals) is kept: it is the denominator of
// `roofline.achieved`.
#include <algorithm>
#include <mutex>
#include <string.h>
#include <utility>
#include <vector>
#include "wm_common.cuh"

wm_prof_t g_wm_prof;
thread_local cudaStream_t wm_dbuf_stream = 0;
thread_local bool wm_dbuf_async = false;

#define WM_PROF_SLOTS 65536
struct wm_prof_launch { cudaEvent_t e0, e1; int kind; };
static std::vector<wm_prof_launch> g_prof_launches;
static std::mutex g_prof_mu;
static cudaEvent_t g_prof_base = 0;
static unsigned long long *g_prof_ctr = 0; // two device counters per slot

int wm_prof_launch_begin(int kind, cudaStream_t st, cudaStream_t zero_st, unsigned long long **ctr)
{
	if (ctr) *ctr = 0;
	if (!g_wm_prof.enabled) return -1;
	int slot = -1;
	wm_prof_launch pl; pl.e0 = pl.e1 = 0; pl.kind = kind;
	{
		std::lock_guard<std::mutex> lk(g_prof_mu);
		if (!g_prof_ctr) {
			WM_CUDA_CHECK(cudaMalloc((void**)&g_prof_ctr, sizeof(unsigned long long) * 2 * WM_PROF_SLOTS));
			WM_CUDA_CHECK(cudaMemset(g_prof_ctr, 0, sizeof(unsigned long long) * 2 * WM_PROF_SLOTS));
		}
		if ((int)g_prof_launches.size() >= WM_PROF_SLOTS) return -1;
		slot = (int)g_prof_launches.size();
		WM_CUDA_CHECK(cudaEventCreate(&pl.e0)); WM_CUDA_CHECK(cudaEventCreate(&pl.e1));
		g_prof_launches.push_back(pl);
	}
	if (ctr) {
		*ctr = g_prof_ctr + 2 * slot;
		WM_CUDA_CHECK(cudaMemsetAsync(*ctr, 0, 2 * sizeof(unsigned long long), zero_st));
	}
	WM_CUDA_CHECK(cudaEventRecord(pl.e0, st));
	return slot;
}

void wm_prof_launch_end(int slot, cudaStream_t st)
{
	if (slot < 0) return;
	cudaEvent_t e1;
	{ std::lock_guard<std::mutex> lk(g_prof_mu); e1 = g_prof_launches[slot].e1; }
	WM_CUDA_CHECK(cudaEventRecord(e1, st));
}

void wm_prof_add(int kind, double alg_bytes, double units, double units2)
{
	if (!g_wm_prof.enabled) return;
	std::lock_guard<std::mutex> lk(g_prof_mu);
	g_wm_prof.k[kind].alg_bytes += alg_bytes, g_wm_prof.k[kind].units += units, g_wm_prof.k[kind].units2 += units2;
}

// start of a profiled region: a base event on the legacy stream gives all launches a common time axis
void wm_prof_region_begin(void)
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	for (auto &l : g_prof_launches) { cudaEventDestroy(l.e0); cudaEventDestroy(l.e1); }
	g_prof_launches.clear();
	if (!g_prof_base) WM_CUDA_CHECK(cudaEventCreate(&g_prof_base));
	WM_CUDA_CHECK(cudaDeviceSynchronize());
	WM_CUDA_CHECK(cudaEventRecord(g_prof_base, 0));
	WM_CUDA_CHECK(cudaEventSynchronize(g_prof_base));
}

void wm_prof_collect(void)
{
	std::lock_guard<std::mutex> lk(g_prof_mu);
	if (g_prof_launches.empty()) return;
	WM_CUDA_CHECK(cudaDeviceSynchronize());
	std::vector<std::pair<float, float>> iv[WM_PK_N];
	std::vector<unsigned long long> ctr(2 * g_prof_launches.size());
	WM_CUDA_CHECK(cudaMemcpy(ctr.data(), g_prof_ctr, sizeof(unsigned long long) * ctr.size(), cudaMemcpyDeviceToHost));
	for (size_t i = 0; i < g_prof_launches.size(); ++i) {
		wm_prof_launch &l = g_prof_launches[i];
		float t0 = 0.f, t1 = 0.f;
		WM_CUDA_CHECK(cudaEventElapsedTime(&t0, g_prof_base, l.e0));
		WM_CUDA_CHECK(cudaEventElapsedTime(&t1, g_prof_base, l.e1));
		wm_prof_kind &K = g_wm_prof.k[l.kind];
		iv[l.kind].push_back(std::make_pair(t0, t1));
		K.ms += t1 - t0; ++K.launches;
		if (l.kind == WM_PK_FILL) { // the kernel counted its block cells: 1 B of backtrack per block cell
			const double c = (double)(ctr[2 * i] + ctr[2 * i + 1]);
			K.units += c; K.alg_bytes += c;
		}
		cudaEventDestroy(l.e0); cudaEventDestroy(l.e1);
	}
	g_prof_launches.clear();
	for (int k = 0; k < WM_PK_N; ++k) {
		if (iv[k].empty()) continue;
		std::sort(iv[k].begin(), iv[k].end());
		float cur0 = iv[k][0].first, cur1 = iv[k][0].second; double uni = 0;
		for (size_t i = 1; i < iv[k].size(); ++i) {
			if (iv[k][i].first > cur1) { uni += cur1 - cur0; cur0 = iv[k][i].first; cur1 = iv[k][i].second; }
			else if (iv[k][i].second > cur1) cur1 = iv[k][i].second;
		}
		uni += cur1 - cur0;
		g_wm_prof.k[k].union_ms += uni;
	}
}
