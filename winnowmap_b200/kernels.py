"""numpy-facing wrappers over the kernel-level C-ABI entry points."""
import ctypes as C

import numpy as np

from ._lib import ExtZ, i8p, i32p, i64p, lib, u8p, u32p, u64p

EZ_FIELDS = ("max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end", "n_cigar")


def _p(a, t):
    return a.ctypes.data_as(t)


def _concat(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) and off[-1] else np.zeros(0, np.uint8)
    return np.ascontiguousarray(buf), off


def ksw_extd2_batch(queries, targets, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, cigar_cap=None):
    """Batched ksw_extd2 (reference src/ksw2_extd2_sse.c:26).  Returns (ez[n,11] int32, [cigar arrays])."""
    n = len(queries)
    qb, qoff = _concat(queries)
    tb, toff = _concat(targets)
    as32 = lambda v: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.int32), (n,)))
    w, zdrop, end_bonus, flag = as32(w), as32(zdrop), as32(end_bonus), as32(flag)
    if cigar_cap is None:
        cap = np.array([len(a) + len(b) + 2 for a, b in zip(queries, targets)], dtype=np.int64)
    else:
        cap = np.broadcast_to(np.asarray(cigar_cap, dtype=np.int64), (n,))
    coff = np.zeros(n + 1, dtype=np.int64)
    coff[1:] = np.cumsum(cap)
    ez = (ExtZ * max(n, 1))()
    cig = np.zeros(max(int(coff[-1]), 1), dtype=np.uint32)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    if len(qb) == 0:
        qb = np.zeros(1, np.uint8)
    if len(tb) == 0:
        tb = np.zeros(1, np.uint8)
    lib().wm_ksw_extd2_batch(n, _p(qb, u8p), _p(qoff, i64p), _p(tb, u8p), _p(toff, i64p), _p(mat, i8p), q, e, q2, e2,
                             _p(w, i32p), _p(zdrop, i32p), _p(end_bonus, i32p), _p(flag, i32p), ez, _p(cig, u32p), _p(coff, i64p))
    out = np.array([[getattr(ez[i], f) for f in EZ_FIELDS] for i in range(n)], dtype=np.int32).reshape(n, len(EZ_FIELDS))
    cigs = [cig[coff[i]: coff[i] + min(out[i, 10], cap[i])].copy() for i in range(n)]
    return out, cigs
