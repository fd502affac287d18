// Exclusive prefix sum over int32 counts -> int64 offsets (n+1 outputs), three small kernels.
// Used for stream compaction (minimizers, anchors); bandwidth-trivial next to the DP.
#pragma once
#include "wm_common.cuh"

#define WM_SCAN_BLOCK 256
#define WM_SCAN_ITEMS 8   // items per thread -> 2048 per block

static __global__ void wm_scan_block_sums(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ block_sums)
{
	__shared__ long long warp_sums[WM_SCAN_BLOCK / 32];
	const int64_t base = (int64_t)blockIdx.x * WM_SCAN_BLOCK * WM_SCAN_ITEMS;
	long long s = 0;
	for (int k = 0; k < WM_SCAN_ITEMS; ++k) {
		int64_t i = base + (int64_t)k * WM_SCAN_BLOCK + threadIdx.x;
		if (i < n) s += in[i];
	}
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		long long t = 0;
		for (int i = 0; i < WM_SCAN_BLOCK / 32; ++i) t += warp_sums[i];
		block_sums[blockIdx.x] = t;
	}
}

// single block: exclusive scan of block sums in place; writes the grand total to *total
static __global__ void wm_scan_top(int64_t *__restrict__ block_sums, int64_t n_blocks, int64_t *__restrict__ total)
{
	__shared__ long long sh[1024];
	__shared__ long long carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int64_t base = 0; base < n_blocks; base += 1024) {
		int64_t i = base + threadIdx.x;
		long long v = i < n_blocks ? block_sums[i] : 0;
		sh[threadIdx.x] = v;
		__syncthreads();
		for (int o = 1; o < 1024; o <<= 1) {
			long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
			__syncthreads();
			sh[threadIdx.x] += t;
			__syncthreads();
		}
		long long incl = sh[threadIdx.x];
		if (i < n_blocks) block_sums[i] = carry + incl - v;
		__syncthreads();
		if (threadIdx.x == 1023) carry += incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry;
}

static __global__ void wm_scan_apply(const int32_t *__restrict__ in, int64_t n, const int64_t *__restrict__ block_sums, int64_t *__restrict__ out)
{
	// each thread owns WM_SCAN_ITEMS consecutive items so that a block scans a contiguous range
	__shared__ long long sh[WM_SCAN_BLOCK];
	const int64_t base = (int64_t)blockIdx.x * WM_SCAN_BLOCK * WM_SCAN_ITEMS + (int64_t)threadIdx.x * WM_SCAN_ITEMS;
	long long loc[WM_SCAN_ITEMS], s = 0;
	for (int k = 0; k < WM_SCAN_ITEMS; ++k) {
		int64_t i = base + k;
		loc[k] = s;
		if (i < n) s += in[i];
	}
	sh[threadIdx.x] = s;
	__syncthreads();
	for (int o = 1; o < WM_SCAN_BLOCK; o <<= 1) {
		long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
		__syncthreads();
		sh[threadIdx.x] += t;
		__syncthreads();
	}
	long long excl = sh[threadIdx.x] - s + block_sums[blockIdx.x];
	for (int k = 0; k < WM_SCAN_ITEMS; ++k) {
		int64_t i = base + k;
		if (i < n) out[i] = excl + loc[k];
	}
}

// out has n+1 entries (out[n] = total); tmp needs (n/2048 + 2) int64.
static inline size_t wm_scan_tmp_elems(int64_t n) { return (size_t)(n / (WM_SCAN_BLOCK * WM_SCAN_ITEMS) + 2); }

static inline void wm_exclusive_scan(const int32_t *d_in, int64_t n, int64_t *d_out, int64_t *d_tmp, cudaStream_t st)
{
	if (n <= 0) { WM_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(int64_t), st)); return; }
	int64_t nb = (n + WM_SCAN_BLOCK * WM_SCAN_ITEMS - 1) / (WM_SCAN_BLOCK * WM_SCAN_ITEMS);
	wm_count_launch(); wm_scan_block_sums<<<(unsigned)nb, WM_SCAN_BLOCK, 0, st>>>(d_in, n, d_tmp);
	wm_count_launch(); wm_scan_top<<<1, 1024, 0, st>>>(d_tmp, nb, d_out + n);
	wm_count_launch(); wm_scan_apply<<<(unsigned)nb, WM_SCAN_BLOCK, 0, st>>>(d_in, n, d_tmp, d_out);
	WM_CUDA_CHECK(cudaGetLastError());
}
