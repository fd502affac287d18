"""Python face of the drop-in boundary: index construction and file/batch mapping through the C ABI."""
import ctypes as C
import os

from ._lib import lib


class IdxOpt(C.Structure):  # wm_idxopt_t == mm_idxopt_t (reference src/minimap.h:106-110)
    _fields_ = [("k", C.c_short), ("w", C.c_short), ("flag", C.c_short), ("bucket_bits", C.c_short),
                ("mini_batch_size", C.c_int), ("batch_size", C.c_uint64)]


class MapOpt(C.Structure):  # wm_mapopt_t == mm_mapopt_t (reference src/minimap.h:112-176)
    _fields_ = [("flag", C.c_int64), ("seed", C.c_int), ("sdust_thres", C.c_int), ("max_qlen", C.c_int), ("bw", C.c_int),
                ("max_gap", C.c_int), ("max_gap_ref", C.c_int), ("min_gap_ref", C.c_int), ("max_frag_len", C.c_int),
                ("max_chain_skip", C.c_int), ("max_chain_iter", C.c_int), ("min_cnt", C.c_int), ("min_chain_score", C.c_int),
                ("chain_gap_scale", C.c_float), ("SVaware", C.c_bool), ("SVawareMinReadLength", C.c_int), ("suffixSampleOffset", C.c_int),
                ("min_mapq", C.c_int), ("min_qcov", C.c_float), ("minPrefixLength", C.c_int), ("maxPrefixLength", C.c_int),
                ("prefixIncrementFactor", C.c_float), ("stage2_bw", C.c_int), ("stage2_zdrop_inv", C.c_int), ("stage2_max_gap", C.c_int),
                ("stage2_extension_inc", C.c_int), ("mask_level", C.c_float), ("mask_len", C.c_int), ("pri_ratio", C.c_float),
                ("best_n", C.c_int), ("max_join_long", C.c_int), ("max_join_short", C.c_int), ("min_join_flank_sc", C.c_int),
                ("min_join_flank_ratio", C.c_float), ("alt_drop", C.c_float), ("a", C.c_int), ("b", C.c_int), ("q", C.c_int),
                ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int), ("sc_ambi", C.c_int), ("noncan", C.c_int), ("junc_bonus", C.c_int),
                ("zdrop", C.c_int), ("zdrop_inv", C.c_int), ("end_bonus", C.c_int), ("min_dp_max", C.c_int), ("min_ksw_len", C.c_int),
                ("anchor_ext_len", C.c_int), ("anchor_ext_shift", C.c_int), ("max_clip_ratio", C.c_float), ("pe_ori", C.c_int),
                ("pe_bonus", C.c_int), ("mid_occ_frac", C.c_float), ("min_mid_occ", C.c_int32), ("mid_occ", C.c_int32),
                ("max_occ", C.c_int32), ("mini_batch_size", C.c_int), ("max_sw_mat", C.c_int64), ("kmer_freq_filename", C.c_char_p),
                ("split_prefix", C.c_char_p)]


F_CIGAR, F_OUT_SAM, F_OUT_CG, F_NO_PRINT_2ND, F_PAF_NO_HIT = 0x004, 0x008, 0x020, 0x4000, 0x8000000

STAT_NAMES = ("n_reads", "n_bases", "n_minimaps", "n_chained", "n_dp_jobs", "n_ll_jobs", "n_rounds", "t_seed", "t_dp", "t_host",
              "t_index", "t_map", "n_keys", "n_pos")


def _setup(L):
    if getattr(L, "_wm_mapper_ready", False):
        return L
    L.wm_set_opt.argtypes = [C.c_char_p, C.POINTER(IdxOpt), C.POINTER(MapOpt)]
    L.wm_check_opt.argtypes = [C.POINTER(IdxOpt), C.POINTER(MapOpt)]
    L.wm_index_build.restype = C.c_void_p
    L.wm_index_build.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.wm_gpu_destroy.argtypes = [C.c_void_p]
    L.wm_idx_blob_size.restype = C.c_int64
    L.wm_idx_blob_size.argtypes = [C.c_void_p]
    L.wm_idx_blob_write.argtypes = [C.c_void_p, C.c_void_p]
    L.wm_idx_blob_load.restype = C.c_void_p
    L.wm_idx_blob_load.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    L.wm_map_file.argtypes = [C.c_void_p, C.POINTER(MapOpt), C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
    L.wm_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    L.wm_reset_stats.argtypes = [C.c_void_p]
    assert L.wm_sizeof_mapopt() == C.sizeof(MapOpt), (L.wm_sizeof_mapopt(), C.sizeof(MapOpt))
    L._wm_mapper_ready = True
    return L


F_OUT_SAM = 0x008


def make_options(preset=None, cigar=True, sam=False):
    """mm_set_opt(0) then mm_set_opt(preset) (reference main.c:144-159); -c sets MM_F_OUT_CG|MM_F_CIGAR, -a sets
    MM_F_OUT_SAM|MM_F_CIGAR (main.c)."""
    L = _setup(lib())
    io, mo = IdxOpt(), MapOpt()
    L.wm_set_opt(None, C.byref(io), C.byref(mo))
    if preset is not None and L.wm_set_opt(preset.encode(), C.byref(io), C.byref(mo)) != 0:
        raise ValueError(f"unknown preset {preset}")
    if sam:
        mo.flag |= F_OUT_SAM | F_CIGAR
    elif cigar:
        mo.flag |= F_OUT_CG | F_CIGAR
    rc = L.wm_check_opt(C.byref(io), C.byref(mo))
    if rc < 0:
        raise ValueError(f"mm_check_opt-style validation failed: {rc}")
    return io, mo


class Mapper:
    """winnowmap [-W rep.txt] -x preset -c ref.fa reads.fa  on one GPU."""

    def __init__(self, ref, kmer_freq=None, preset="map-ont", cigar=True, device=0, n_threads=None, blob=None, sam=False):
        self.L = _setup(lib())
        self.io, self.mo = make_options(preset, cigar, sam)
        self.n_threads = n_threads or max(1, min(64, (os.cpu_count() or 2) // 2))
        if blob is not None:  # index received from another rank (numpy uint8 array)
            self._blob_keep = blob
            self.ctx = self.L.wm_idx_blob_load(blob.ctypes.data, blob.nbytes, device)
        else:
            self.ctx = self.L.wm_index_build(ref.encode(), kmer_freq.encode() if kmer_freq else None, self.io.k, self.io.w, device)
        if not self.ctx:
            raise RuntimeError("index construction failed")

    def index_blob(self):
        """The flattened index as one numpy uint8 array (for the one-time NCCL fan-out)."""
        import numpy as np
        n = self.L.wm_idx_blob_size(self.ctx)
        buf = np.empty(n, dtype=np.uint8)
        self.L.wm_idx_blob_write(self.ctx, buf.ctypes.data)
        return buf

    def map_file(self, reads, out, rank=0, world=1, tag_order=False, max_batch_bases=200_000_000):
        rc = self.L.wm_map_file(self.ctx, C.byref(self.mo), reads.encode(), out.encode(), self.n_threads, rank, world, int(tag_order), max_batch_bases)
        if rc != 0:
            raise RuntimeError("wm_map_file failed")

    def stats(self):
        v = (C.c_double * len(STAT_NAMES))()
        self.L.wm_get_stats(self.ctx, v, len(STAT_NAMES))
        return dict(zip(STAT_NAMES, list(v)))

    def reset_stats(self):
        self.L.wm_reset_stats(self.ctx)

    def close(self):
        if self.ctx:
            self.L.wm_gpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
