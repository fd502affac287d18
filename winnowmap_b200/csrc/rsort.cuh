// Tie-exact emulation of the reference's in-place MSD radix sort (src/ksort.h:98-151,
// instantiated at src/misc.c:156-159).  The sort is NOT stable: equal keys end up in the order
// the cycle-leader permutation (:126-138) leaves them, buckets of <= 64 elements are finished by a
// stable insertion sort (:105-115).  That order reaches the output (which predecessor chaining sees
// first), so it is reproduced step by step; one thread walks one array.  Passes in which every key
// falls into one bucket are identity permutations and are skipped.
#pragma once
#include "wm_common.cuh"
#include "sketch.cuh"

#define WM_RS_MIN_SIZE 64

struct wm_rs_frame { int beg, end, s, k; int e[256]; };
struct wm_rs_stack { wm_rs_frame f[8]; int b[256]; };

template <typename T> struct wm_rs_key;
template <> struct wm_rs_key<wm128_dev> { __device__ static __forceinline__ uint64_t get(const wm128_dev &a) { return a.x; } };
template <> struct wm_rs_key<uint64_t> { __device__ static __forceinline__ uint64_t get(const uint64_t &a) { return a; } };

template <typename T>
__device__ void wm_rs_insertsort(T *beg, T *end)
{ // ksort.h:105-115
	for (T *i = beg + 1; i < end; ++i)
		if (wm_rs_key<T>::get(*i) < wm_rs_key<T>::get(*(i - 1))) {
			T *j, tmp = *i;
			for (j = i; j > beg && wm_rs_key<T>::get(tmp) < wm_rs_key<T>::get(*(j - 1)); --j) *j = *(j - 1);
			*j = tmp;
		}
}

// one partition pass of rs_sort (ksort.h:121-139) on a[beg,end) at shift s; fills F.e[] with bucket ends
template <typename T>
__device__ void wm_rs_pass(T *a, wm_rs_frame &F, int *b)
{
	const int beg = F.beg, end = F.end, s = F.s;
	int *e = F.e;
	for (int k = 0; k < 256; ++k) e[k] = 0;
	for (int i = beg; i < end; ++i) ++e[wm_rs_key<T>::get(a[i]) >> s & 255];
	int single = -1;
	for (int k = 0; k < 256; ++k) if (e[k] == end - beg) single = k;
	{ // prefix: b[k] = start, e[k] = end of bucket k
		int acc = beg;
		for (int k = 0; k < 256; ++k) { b[k] = acc; acc += e[k]; e[k] = acc; }
	}
	if (single >= 0) return; // identity permutation
	for (int k = 0; k < 256;) {
		if (b[k] != e[k]) {
			int l = (int)(wm_rs_key<T>::get(a[b[k]]) >> s & 255);
			if (l != k) {
				T tmp = a[b[k]], swap;
				do {
					swap = tmp; tmp = a[b[l]]; a[b[l]++] = swap;
					l = (int)(wm_rs_key<T>::get(tmp) >> s & 255);
				} while (l != k);
				a[b[k]++] = tmp;
			} else ++b[k];
		} else ++k;
	}
}

// ---- warp-cooperative variant ----
// The cycle-leader walk of one pass is a serial dependence chain and stays on lane 0, but everything
// around it is data parallel: the byte histogram (shared-memory atomics), the bucket prefix, and the
// <= 64-element insertion sorts that finish the leaf buckets (disjoint ranges, one per lane).  Buckets
// that need another pass go to a work list; the ranges are disjoint, so the order in which they are
// processed cannot change the result.  `a` may point to shared memory (the caller staged the array)
// or to global memory.
struct wm_rs_range { int beg, end, s; };
struct wm_rs_warp_ws { int e[256], b[256]; };

// s0: the byte the first pass looks at (56 for a whole array; lower when the caller has done the upper passes itself)
template <typename T>
__device__ void wm_radix_sort_warp_from(T *a, int n, int s0, wm_rs_warp_ws *W, wm_rs_range *wl, int lane)
{
	const unsigned FULL = 0xffffffffu;
	if (n <= WM_RS_MIN_SIZE) { if (lane == 0) wm_rs_insertsort(a, a + n); __syncwarp(); return; }
	int n_wl = 0;
	int beg = 0, end = n, s = s0;
	for (;;) {
		// histogram of byte s>>3 (ksort.h:121-122)
		#pragma unroll
		for (int k = 0; k < 8; ++k) W->e[lane * 8 + k] = 0;
		__syncwarp();
		for (int i = beg + lane; i < end; i += 32) atomicAdd(&W->e[wm_rs_key<T>::get(a[i]) >> s & 255], 1);
		__syncwarp();
		int cnt[8], sum = 0; bool single = false;
		#pragma unroll
		for (int k = 0; k < 8; ++k) { cnt[k] = W->e[lane * 8 + k]; sum += cnt[k]; single |= cnt[k] == end - beg; }
		int incl = sum;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
		int acc = beg + incl - sum;
		#pragma unroll
		for (int k = 0; k < 8; ++k) { W->b[lane * 8 + k] = acc; acc += cnt[k]; W->e[lane * 8 + k] = acc; }
		const bool one_bucket = __any_sync(FULL, single);
		__syncwarp();
		if (one_bucket) { // identity permutation: straight to the next byte
			if (s > 0) { s = s > 8 ? s - 8 : 0; continue; }
		} else {
			if (lane == 0) { // ksort.h:126-138
				int *b = W->b; const int *e = W->e;
				for (int k = 0; k < 256;) {
					const int bk = b[k];
					if (bk != e[k]) {
						int l = (int)(wm_rs_key<T>::get(a[bk]) >> s & 255);
						if (l != k) {
							T tmp = a[bk], swap;
							do {
								const int bl = b[l];
								swap = tmp; tmp = a[bl]; a[bl] = swap; b[l] = bl + 1;
								l = (int)(wm_rs_key<T>::get(tmp) >> s & 255);
							} while (l != k);
							a[bk] = tmp; b[k] = bk + 1;
						} else b[k] = bk + 1;
					} else ++k;
				}
			}
			__syncwarp();
			if (s > 0) { // ksort.h:140-145: finish or queue the sub-buckets
				const int ns = s > 8 ? s - 8 : 0;
				#pragma unroll 1
				for (int kb = 0; kb < 256; kb += 32) {
					const int k = kb + lane;
					const int cb = k ? W->e[k - 1] : beg, ce = W->e[k];
					const bool big = ce - cb > WM_RS_MIN_SIZE;
					const unsigned m = __ballot_sync(FULL, big);
					if (big) { wm_rs_range r; r.beg = cb, r.end = ce, r.s = ns; wl[n_wl + __popc(m & ((1u << lane) - 1u))] = r; }
					else if (ce - cb > 1) wm_rs_insertsort(a + cb, a + ce);
					n_wl += __popc(m);
				}
				__syncwarp();
			}
		}
		if (n_wl == 0) break;
		const wm_rs_range r = wl[--n_wl];
		beg = r.beg, end = r.end, s = r.s;
	}
	__syncwarp();
}

template <typename T>
__device__ __forceinline__ void wm_radix_sort_warp(T *a, int n, wm_rs_warp_ws *W, wm_rs_range *wl, int lane) { wm_radix_sort_warp_from(a, n, 56, W, wl, lane); }

// radix_sort_##name (ksort.h:146-150) on a[0,n)
template <typename T>
__device__ void wm_radix_sort_emul(T *a, int n, wm_rs_stack *stk)
{
	if (n <= WM_RS_MIN_SIZE) { wm_rs_insertsort(a, a + n); return; }
	int depth = 0;
	stk->f[0].beg = 0, stk->f[0].end = n, stk->f[0].s = 56, stk->f[0].k = -1;
	while (depth >= 0) {
		wm_rs_frame &F = stk->f[depth];
		if (F.k < 0) { wm_rs_pass(a, F, stk->b); F.k = 0; if (F.s == 0) { --depth; continue; } }
		bool pushed = false;
		while (F.k < 256) {
			const int k = F.k++;
			const int cb = k ? F.e[k - 1] : F.beg, ce = F.e[k];
			if (ce - cb > WM_RS_MIN_SIZE) {
				wm_rs_frame &G = stk->f[depth + 1];
				G.beg = cb, G.end = ce, G.s = F.s > 8 ? F.s - 8 : 0, G.k = -1;
				++depth; pushed = true;
				break;
			} else if (ce - cb > 1) wm_rs_insertsort(a + cb, a + ce);
		}
		if (!pushed) --depth;
	}
}
