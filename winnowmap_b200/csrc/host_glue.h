#pragma once
#include <vector>
#include "host_types.h"

namespace wmh {
void set_coor(wm_reg1_t *r, int32_t qlen, const wm_pair_t *a);
void gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const wm_pair_t *a, std::vector<wm_reg1_t> &regs);
void split_reg(wm_reg1_t *r, wm_reg1_t *r2, int n, int qlen, const wm_pair_t *a);
void set_parent(float mask_level, int mask_len, int n, wm_reg1_t *r, int sub_diff, int hard_mask_level, float alt_diff_frac);
int set_sam_pri(int n, wm_reg1_t *r);
void sync_regs(int n_regs, wm_reg1_t *regs);
void select_sub(float pri_ratio, int min_diff, int best_n, std::vector<wm_reg1_t> &regs);
void filter_regs(const wm_mapopt_t *opt, int qlen, std::vector<wm_reg1_t> &regs);
int squeeze_a(std::vector<wm_reg1_t> &regs, wm_pair_t *a);
void join_long(const wm_mapopt_t *opt, int qlen, std::vector<wm_reg1_t> &regs, wm_pair_t *a);
void hit_sort(std::vector<wm_reg1_t> &regs, float alt_diff_frac);
void chain_post(const wm_mapopt_t *opt, int k, int qlen, std::vector<wm_reg1_t> &regs, wm_pair_t *a);
void est_err(const wm_host_idx *mi, int qlen, std::vector<wm_reg1_t> &regs, const wm_pair_t *a, int32_t n, const uint64_t *mini_pos);
void set_mapq(std::vector<wm_reg1_t> &regs, int min_chain_sc, int match_sc, int rep_len, int is_sr);
}
