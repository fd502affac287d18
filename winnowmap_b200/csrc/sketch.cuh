#pragma once
#include "wm_common.cuh"

struct wm128_dev { uint64_t x, y; };

// device view of the down-weight filter (ext/bloom/bloom_filter.hpp table + salts)
struct wm_bloom_dev {
	const uint8_t *table;
	uint64_t bits;
	uint32_t salt[2];
	int n_salt;
};

// one sequence to sketch: a slice of a device code array (0..3 = ACGT, 4 = ambiguous)
struct wm_sk_task {
	int64_t seq_off;
	int32_t len;
	uint32_t rid;
};

struct wm_sketch_ws {
	wm_dbuf tasks, offs, ord, elig, flag, cnt, rank, scan_tmp, mz, mz_off;
	void release() { tasks.release(); offs.release(); ord.release(); elig.release(); flag.release(); cnt.release(); rank.release(); scan_tmp.release(); mz.release(); mz_off.release(); }
};

struct wm_bloom_s;
void wm_ascii_to_code(const char *d_in, uint8_t *d_out, int64_t n, cudaStream_t st);
void wm_sketch_run(wm_sketch_ws *ws, const wm_bloom_dev &bf, const uint8_t *d_codes, const wm_sk_task *h_tasks, int n_tasks,
                   int w, int k, int64_t *n_mz, cudaStream_t st);
void wm_bloom_dev_from_table(wm_bloom_dev *d, const uint8_t *d_table, uint64_t bits);
void wm_bloom_params(const wm_bloom_s *b, uint64_t *bits, uint32_t *salt, int *n_salt);
