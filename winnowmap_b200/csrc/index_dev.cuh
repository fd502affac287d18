#pragma once
#include "wm_common.cuh"
#include "sketch.cuh"

// Device-side index construction (index_dev.cu): sorts the n (minimizer, position) pairs of d_a (consumed) by minimizer hash,
// positions ascending, and builds keys / pos_off / pos on the device.
void wm_index_build_dev(wm128_dev *d_a, int64_t n, int k, uint64_t **d_keys_out, uint64_t **d_pos_off_out, uint64_t **d_pos_out, int64_t *n_keys_out, cudaStream_t st);
