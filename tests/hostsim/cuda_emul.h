// TEST INFRASTRUCTURE.  Host emulation of the CUDA device features the DP sweeps use, so that the kernels' arithmetic
// (csrc/ksw_extd2_v2.cuh, csrc/ksw_extd2_common.cuh) can be compiled with g++ and checked against the oracle without a
// GPU: a warp is 32 lockstep fibers of one host thread, shuffles and __syncwarp are barriers around a shared exchange buffer,
// the 16x2 SIMD intrinsics and prmt/funnel-shift are restated from the PTX ISA.
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <cuda_runtime.h> // types only (uint2, cudaStream_t); __device__ / __forceinline__ expand to host-neutral forms

#define WM_HOST_EMUL 1

namespace wm_emul {
// The 32 lanes are fibers (ucontext) of one host thread, scheduled round-robin; a barrier is "yield until everybody
// has arrived", so a warp-level operation costs a few context switches instead of an OS barrier.
struct Warp {
	ucontext_t main_ctx, ctx[32];
	char *stack[32];
	bool done[32];
	int cur, arrived, gen;
	uint64_t xchg[32];
	void (*body)(int lane, void *arg); void *arg;
};
extern thread_local Warp *warp;
extern thread_local int lane;
inline void yield_lane()
{ // hand over to the next unfinished lane; returns when this lane is scheduled again
	Warp *w = warp;
	const int me = w->cur;
	int nxt = me;
	do { nxt = (nxt + 1) & 31; } while (w->done[nxt] && nxt != me);
	if (nxt == me) return;
	w->cur = nxt;
	swapcontext(&w->ctx[me], &w->ctx[nxt]);
	lane = me; // (thread_local shared by all fibers)
}
inline void sync()
{
	Warp *w = warp;
	const int gen = w->gen;
	if (++w->arrived == 32) { w->arrived = 0; ++w->gen; return; }
	while (w->gen == gen) yield_lane();
}
template <typename T> inline T exchange(T v, int src_lane)
{
	uint64_t u = 0; memcpy(&u, &v, sizeof(T));
	warp->xchg[lane] = u;
	sync();
	const uint64_t r = warp->xchg[src_lane & 31];
	sync();
	T out; memcpy(&out, &r, sizeof(T));
	return out;
}
void run_warp(void (*body)(int lane, void *arg), void *arg); // kernel_emul.cpp: runs body on 32 lockstep lanes
} // namespace wm_emul

inline void __syncwarp(unsigned = 0xffffffffu) { wm_emul::sync(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return wm_emul::exchange(v, src); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32)
{ const int src = wm_emul::lane - (int)d; const T r = wm_emul::exchange(v, src < 0 ? wm_emul::lane : src); return src < 0 ? v : r; }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return wm_emul::exchange(v, wm_emul::lane ^ m); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

// ---- integer SIMD-in-a-word intrinsics (CUDA math API semantics) ----
inline uint32_t wm_emul_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
inline int wm_emul_lo(uint32_t a) { return (int16_t)(a & 0xffffu); }
inline int wm_emul_hi(uint32_t a) { return (int16_t)(a >> 16); }
inline unsigned __vadd2(unsigned a, unsigned b) { return wm_emul_pack(wm_emul_lo(a) + wm_emul_lo(b), wm_emul_hi(a) + wm_emul_hi(b)); }
inline unsigned __vsub2(unsigned a, unsigned b) { return wm_emul_pack(wm_emul_lo(a) - wm_emul_lo(b), wm_emul_hi(a) - wm_emul_hi(b)); }
inline unsigned __vmins2(unsigned a, unsigned b)
{ return wm_emul_pack(wm_emul_lo(a) < wm_emul_lo(b) ? wm_emul_lo(a) : wm_emul_lo(b), wm_emul_hi(a) < wm_emul_hi(b) ? wm_emul_hi(a) : wm_emul_hi(b)); }
inline unsigned __vmaxs2(unsigned a, unsigned b)
{ return wm_emul_pack(wm_emul_lo(a) > wm_emul_lo(b) ? wm_emul_lo(a) : wm_emul_lo(b), wm_emul_hi(a) > wm_emul_hi(b) ? wm_emul_hi(a) : wm_emul_hi(b)); }
// max(a + b, c) per signed 16-bit half; the sum wraps to 16 bits first
inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{ const uint64_t v = (uint64_t)hi << 32 | lo; return (uint32_t)(v >> (shift & 31)); }
inline unsigned long long __brevll(unsigned long long x)
{ unsigned long long r = 0; for (int i = 0; i < 64; ++i) r |= (x >> i & 1ULL) << (63 - i); return r; }
// prmt.b32, default mode: byte i of the result is byte (sel nibble i & 7) of {b,a}; nibble bit 3 replicates that byte's sign
inline uint32_t wm_emul_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	const uint64_t v = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		const unsigned n = sel >> (4 * i) & 0xf;
		uint32_t byte = (uint32_t)(v >> (8 * (n & 7))) & 0xffu;
		if (n & 8) byte = (byte & 0x80u) ? 0xffu : 0u;
		r |= byte << (8 * i);
	}
	return r;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) { return wm_emul_prmt(a, b, sel & 0x7777u); }

// ---- warp votes, bit utilities and the rounding-mode intrinsics of the chaining kernels ----
inline unsigned __ballot_sync(unsigned, bool pred)
{ // two rounds: everybody publishes, everybody collects
	wm_emul::warp->xchg[wm_emul::lane] = pred ? 1 : 0;
	wm_emul::sync();
	unsigned m = 0;
	for (int l = 0; l < 32; ++l) m |= (unsigned)(wm_emul::warp->xchg[l] & 1) << l;
	wm_emul::sync();
	return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
// compiled with -ffp-contract=off: plain C operations round to nearest even, once, like the _rn intrinsics
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __ull2float_rn(unsigned long long x) { return (float)x; }
inline float __ll2float_rn(long long x) { return (float)x; }
using std::max;
using std::min;
