// Packed read pool: the bases of a batch live in HBM as 2 bits per base plus a 1-bit "ambiguous" mask, not as one
// byte per base.  Base i of the pool is bits [2(i & 15), 2(i & 15) + 2) of pk[i >> 4] (A, C, G, T = 0..3, the order of
// seq_nt4_table, reference src/sketch.c:19-36) and bit (i & 31) of nm[i >> 5] (set: the table gave 4).  Consumers never
// expand the pool: a k-mer (k <= 28) is ONE unaligned 64-bit window of pk -- its complement is the reverse-strand k-mer,
// its 2-bit-group reversal the forward one (src/sketch.c:162-163 build the same two words base by base) -- and a DP
// query slice is 16 bases per 32-bit window.  Strand 1 of a read is never materialised: a slice of it is the
// complemented forward window read backwards (src/align.c:874-876).
//
// Both arrays are over-allocated by WM_PK_SLACK words so that the aligned 128-bit loads around the last base stay
// inside the allocation.
#pragma once
#include <stdint.h>

#define WM_PK_SLACK 64

struct wm_pkseq { const uint32_t *pk, *nm; };

// 64-bit window of 2-bit codes starting at base b (b >= 0): bits [2j, 2j + 2) = base b + j, j < 32
__device__ __forceinline__ uint64_t wm_pk_window(const uint32_t *__restrict__ pk, int64_t b)
{
	const uint32_t *w = pk + (b >> 4);
	const unsigned sh = (unsigned)(b & 15) * 2;
	const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
	return (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
}
// 32 ambiguity flags starting at base b
__device__ __forceinline__ uint32_t wm_pk_nwindow(const uint32_t *__restrict__ nm, int64_t b)
{
	const uint32_t *w = nm + (b >> 5);
	return __funnelshift_r(w[0], w[1], (unsigned)(b & 31));
}
// reverses the order of the 32 2-bit groups of x
__device__ __forceinline__ uint64_t wm_pk_rev2(uint64_t x)
{
	x = __brevll(x);
	return (x >> 1 & 0x5555555555555555ULL) | (x & 0x5555555555555555ULL) << 1;
}
// forward and reverse-complement k-mer words (src/sketch.c:162-163) of the k bases in the low 2k bits of window v
__device__ __forceinline__ void wm_pk_kmer(uint64_t v, int k, uint64_t *fw, uint64_t *rv)
{
	*fw = wm_pk_rev2(v) >> (64 - 2 * k);
	*rv = ~v & ((1ULL << 2 * k) - 1);
}
// one base as a 0..4 code
__device__ __forceinline__ int wm_pk_get(const wm_pkseq &s, int64_t b)
{
	if (s.nm[b >> 5] >> (b & 31) & 1) return 4;
	return (int)(s.pk[b >> 4] >> ((b & 15) * 2) & 3);
}

// ---- 32 bases of ASCII -> one group of the pool (lut: seq_nt4_table, 256 entries) ----
__device__ __forceinline__ void wm_pk_pack32(const uint32_t raw[8], const uint8_t *lut, uint64_t *pk, uint32_t *nm)
{
	uint64_t p = 0; uint32_t m = 0;
	#pragma unroll
	for (int j = 0; j < 32; ++j) {
		const uint32_t c = lut[raw[j >> 2] >> 8 * (j & 3) & 255];
		p |= (uint64_t)(c & 3) << 2 * j;
		m |= (c >> 2) << j;
	}
	*pk = p, *nm = m;
}

// ---- a window of a read with masked intervals, as a packed sequence of its own (src/map.c:795-801: covered bases
// become ambiguous) ----
struct wm_mask_task { int64_t src_off, dst_off, mask_off; int32_t len, n_mask; }; // dst_off: a multiple of 32 bases

// bases [p0, p0 + 32) of the window of task T; mask_pool: (start, end) pairs relative to the window, sorted by start
__device__ __forceinline__ void wm_pk_mask32(const wm_pkseq &seq, const wm_mask_task &T, int p0, const int32_t *__restrict__ mask_pool, uint64_t *pk, uint32_t *nm)
{
	const uint64_t v = wm_pk_window(seq.pk, T.src_off + p0);
	uint32_t m = wm_pk_nwindow(seq.nm, T.src_off + p0);
	const int32_t *iv = mask_pool + 2 * T.mask_off;
	int a = 0, b = T.n_mask; // intervals with start <= p0
	while (a < b) { int mid = (a + b) >> 1; if (iv[2 * mid] <= p0) a = mid + 1; else b = mid; }
	for (int j = 0; j < 32; ++j) {
		const int p = p0 + j;
		while (a < T.n_mask && iv[2 * a] <= p) ++a;
		if (p >= T.len || (a > 0 && p < iv[2 * (a - 1) + 1])) m |= 1u << j; // covered by the last interval that starts at or before p
	}
	*pk = v, *nm = m;
}

// ---- DP sequences: 16 bytes of 0..4 codes from one unaligned window ----
// kind 0: read strand 0 (src_off: pool base of the slice's first base), 2: read strand 1 (src_off: pool base of the FORWARD
// base that is the slice's first base: the slice runs down from there, complemented), 1: 4-bit packed reference
struct wm_gather_job { int64_t src_off, dst_off; int32_t len, kind, reversed, pad; };

// bytes [p0, p0 + 16) of the slice of job J (zero past its end) as four little-endian words
__device__ __forceinline__ void wm_pk_gather16(const wm_gather_job &J, int p0, const wm_pkseq &rd, const uint32_t *__restrict__ S, uint32_t out[4])
{
	const int nv = J.len - p0 < 16 ? J.len - p0 : 16;
	out[0] = out[1] = out[2] = out[3] = 0u;
	if (nv <= 0) return;
	// byte p of the slice is pool base A + d * p
	const int sg = J.kind == 2 ? -1 : 1, d = J.reversed ? -sg : sg;
	const int64_t A = J.src_off + (J.reversed ? sg * (int64_t)(J.len - 1) : 0);
	const int64_t b0 = d > 0 ? A + p0 : A - (p0 + nv - 1); // lowest base of the 16
	if (J.kind == 1) {
		const uint32_t *wd = S + (b0 >> 3);
		const unsigned sh = (unsigned)(b0 & 7) * 4;
		const uint32_t w0 = wd[0], w1 = wd[1], w2 = wd[2];
		const uint64_t x = (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
		#pragma unroll
		for (int j = 0; j < 16; ++j)
			if (j < nv) {
				uint32_t c = (uint32_t)(x >> 4 * (d > 0 ? j : nv - 1 - j)) & 15u;
				if (c > 4) c = 4;
				out[j >> 2] |= c << 8 * (j & 3);
			}
	} else {
		const uint32_t v = (uint32_t)wm_pk_window(rd.pk, b0), m = wm_pk_nwindow(rd.nm, b0);
		#pragma unroll
		for (int j = 0; j < 16; ++j)
			if (j < nv) {
				const int sl = d > 0 ? j : nv - 1 - j;
				uint32_t c = v >> 2 * sl & 3u;
				if (J.kind == 2) c = 3u - c;
				if (m >> sl & 1u) c = 4;
				out[j >> 2] |= c << 8 * (j & 3);
			}
	}
}
