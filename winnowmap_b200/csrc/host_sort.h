// Host-side emulation of the reference's unstable in-place radix sort (src/ksort.h:98-151; instances
// radix_sort_128x / radix_sort_64 at src/misc.c:156-159).  The glue code sorts tiny arrays (hits,
// chain descriptors) with it and the tie order is observable, so the permutation is reproduced
// exactly; same algorithm as the device version in rsort.cuh, written recursively here.
#pragma once
#include <stdint.h>
#include "host_types.h"

namespace wmh {

inline uint64_t sort_key(const wm_pair_t &a) { return a.x; }
inline uint64_t sort_key(const uint64_t &a) { return a; }

template <typename T>
void insertion_by_key(T *beg, T *end)
{
	for (T *i = beg + 1; i < end; ++i)
		if (sort_key(*i) < sort_key(*(i - 1))) {
			T tmp = *i, *j;
			for (j = i; j > beg && sort_key(tmp) < sort_key(*(j - 1)); --j) *j = *(j - 1);
			*j = tmp;
		}
}

template <typename T>
void flag_sort_level(T *beg, T *end, int shift)
{
	T *head[256], *tail[256];
	size_t cnt[256] = {0};
	for (T *i = beg; i != end; ++i) ++cnt[sort_key(*i) >> shift & 255];
	T *cur = beg;
	for (int k = 0; k < 256; ++k) { head[k] = cur; cur += cnt[k]; tail[k] = cur; }
	for (int k = 0; k < 256;) { // cycle-leader permutation, buckets filled in arrival order
		if (head[k] == tail[k]) { ++k; continue; }
		int l = (int)(sort_key(*head[k]) >> shift & 255);
		if (l == k) { ++head[k]; continue; }
		T carry = *head[k];
		do {
			T displaced = *head[l];
			*head[l]++ = carry;
			carry = displaced;
			l = (int)(sort_key(carry) >> shift & 255);
		} while (l != k);
		*head[k]++ = carry;
	}
	if (shift == 0) return;
	const int next = shift > 8 ? shift - 8 : 0;
	T *b = beg;
	for (int k = 0; k < 256; ++k) {
		T *e = tail[k];
		if (e - b > 64) flag_sort_level(b, e, next);
		else if (e - b > 1) insertion_by_key(b, e);
		b = e;
	}
}

template <typename T>
void radix_sort(T *beg, T *end)
{
	if (end - beg <= 64) insertion_by_key(beg, end);
	else flag_sort_level(beg, end, 56);
}

} // namespace wmh
