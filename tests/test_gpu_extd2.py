"""GPU parity: the CUDA ksw_extd2 (through the C ABI) against the CPU oracle, bit-exact on
the full ksw_extz_t and the CIGAR."""
import numpy as np
import pytest

import oracle_lib as ol
from test_oracle_vs_ref import FLAGS, rand_pair

pytestmark = pytest.mark.gpu


def _check(queries, targets, mat, prm, w, zdrop, eb, flag):
    from winnowmap_b200 import kernels
    ez, cigs = kernels.ksw_extd2_batch(queries, targets, mat, *prm, w, zdrop, eb, flag)
    for i in range(len(queries)):
        e0, c0 = ol.oracle_extd2(queries[i], targets[i], mat, *prm, int(w[i]), int(zdrop[i]), int(eb[i]), int(flag[i]))
        assert np.array_equal(e0, ez[i]), (i, len(queries[i]), len(targets[i]), int(w[i]), int(flag[i]), e0, ez[i])
        assert np.array_equal(c0, cigs[i]), (i, len(queries[i]), len(targets[i]), int(w[i]), int(flag[i]))


@pytest.mark.parametrize("seed", range(3))
def test_extd2_random_batch(seed):
    rng = np.random.default_rng(900 + seed)
    mat = ol.simple_mat()
    Q, T, W, Z, E, F = [], [], [], [], [], []
    for it in range(400):
        tlen = int(rng.choice([1, 5, 16, 17, 33, 100, 250, 300, 700, 1100, 2500]))
        q, t = rand_pair(rng, tlen, err=float(rng.choice([0.02, 0.1, 0.3])), drift=int(rng.choice([0, 0, 30, 120, 400])),
                         n_runs=int(rng.integers(0, 3)))
        Q.append(q); T.append(t)
        W.append(int(rng.choice([5, 20, 50, 100, 751, 3001]))); Z.append(int(rng.choice([400, 200, 50, -1])))
        E.append(int(rng.choice([-1, 0, 10]))); F.append(FLAGS[int(rng.integers(0, len(FLAGS)))])
    _check(Q, T, mat, (4, 2, 24, 1), np.array(W), np.array(Z), np.array(E), np.array(F))


def test_extd2_asm_scoring_and_swapped_gaps():
    rng = np.random.default_rng(77)
    for a, b, q, e, q2, e2 in [(1, 4, 6, 2, 26, 1), (2, 4, 24, 1, 4, 2)]:
        mat = ol.simple_mat(a, b, 1)
        Q, T = zip(*[rand_pair(rng, int(rng.integers(20, 900)), err=0.05, drift=int(rng.choice([0, 50]))) for _ in range(64)])
        n = len(Q)
        _check(list(Q), list(T), mat, (q, e, q2, e2), np.full(n, 200), np.full(n, 200), np.full(n, -1),
               np.array([FLAGS[i % 4] for i in range(n)]))


def test_extd2_long_end_extension_and_empty():
    rng = np.random.default_rng(5)
    mat = ol.simple_mat()
    q, t = rand_pair(rng, 9000, err=0.08, drift=300)
    Q = [q, q[:1], np.zeros(0, np.uint8), q[:300]]
    T = [t, t[:1], t[:10], np.zeros(0, np.uint8)]
    n = len(Q)
    _check(Q, T, mat, (4, 2, 24, 1), np.array([3001, 751, 751, 751]), np.full(n, 400), np.full(n, -1), np.array([0x40, 0x40, 0, 0]))


def test_cigar_capacity_overflow_is_reported():
    from winnowmap_b200 import kernels
    rng = np.random.default_rng(6)
    mat = ol.simple_mat()
    q, t = rand_pair(rng, 800, err=0.2)
    e0, c0 = ol.oracle_extd2(q, t, mat, 4, 2, 24, 1, 751, 400, -1, 0)
    ez, cigs = kernels.ksw_extd2_batch([q], [t], mat, 4, 2, 24, 1, 751, 400, -1, 0, cigar_cap=4)
    assert ez[0, 10] == len(c0) and len(c0) > 4


@pytest.mark.parametrize("seed", range(2))
def test_extz2_single_gap_pair_batch(seed):
    """Equal gap pairs route to the single-affine kernel (ksw_extz2_sse, src/ksw2_extz2_sse.c:23; dispatch src/align.c:328-331):
    the extd2 matrix of sizes / bands / flags, several scoring sets incl. ones that strain the unsigned-offset encoding,
    a 9 kb end extension (state rows in the global slice) and empty inputs, against the oracle's extz2."""
    from winnowmap_b200 import kernels
    rng = np.random.default_rng(950 + seed)
    for a, b, go, ge in [(2, 4, 4, 2), (1, 4, 6, 2), (1, 9, 16, 2), (5, 4, 40, 20)]:
        mat = ol.simple_mat(a, b, 1)
        Q, T, W, Z, E, F = [], [], [], [], [], []
        for it in range(150):
            tlen = int(rng.choice([1, 5, 16, 17, 33, 100, 250, 300, 700, 1100, 2500]))
            q, t = rand_pair(rng, tlen, err=float(rng.choice([0.02, 0.1, 0.3])), drift=int(rng.choice([0, 0, 30, 120, 400])), n_runs=int(rng.integers(0, 3)))
            Q.append(q); T.append(t)
            W.append(int(rng.choice([5, 20, 50, 100, 751, 3001, -1]))); Z.append(int(rng.choice([400, 200, 50, -1])))
            E.append(int(rng.choice([-1, 0, 10]))); F.append(FLAGS[int(rng.integers(0, len(FLAGS)))] | (0x10 if it % 7 == 0 else 0))
        if seed == 0 and (a, b) == (2, 4):
            q, t = rand_pair(rng, 9000, err=0.08, drift=300)
            Q += [q, np.zeros(0, np.uint8), q[:300]]; T += [t, t[:10], np.zeros(0, np.uint8)]
            W += [3001, 751, 751]; Z += [400, 400, 400]; E += [-1, -1, -1]; F += [0x40, 0, 0]
        ez, cigs = kernels.ksw_extd2_batch(Q, T, mat, go, ge, go, ge, np.array(W), np.array(Z), np.array(E), np.array(F))
        for i in range(len(Q)):
            e0, c0 = ol.oracle_extz2(Q[i], T[i], mat, go, ge, W[i], Z[i], E[i], F[i])
            assert np.array_equal(e0, ez[i]), (i, (a, b, go, ge), len(Q[i]), len(T[i]), W[i], hex(F[i]), e0, ez[i])
            assert np.array_equal(c0, cigs[i]), (i, (a, b, go, ge), len(Q[i]), len(T[i]), W[i], hex(F[i]))


def test_exts2_splice_batch():
    """The splice-aware extension kernel (csrc/ksw_exts2.cuh = ksw_exts2_sse, src/ksw2_exts2_sse.c:26) through the C ABI against the
    oracle's restatement (pinned to the reference's function by tests/test_oracle_vs_ref.py): the splice presets' scorings, both
    transcript strands, flank bonus, junction annotation, reversed inputs, Z-drop, generic scoring, a target whose state rows
    live in the global slice, empty inputs and a scoring the reference refuses (q2 <= q + e)."""
    from winnowmap_b200 import kernels
    from test_kernel_emulation import _splice_cases
    from test_oracle_vs_ref import spliced_pair
    rng = np.random.default_rng(990)
    cases = _splice_cases(rng, 200)
    n_intron = 0
    for k in range(5):
        sub = [c for i, c in enumerate(cases) if i % 5 == k]
        a, b, go, ge, go2, noncan, jb = sub[0][4]
        if k == 0:  # a long one (global state slice) and empty inputs
            q, t = spliced_pair(rng, 12, err=0.05)
            sub.append((q, np.concatenate([t, rng.integers(0, 4, size=1500, dtype=np.uint8)]), 0x100 | 0x400 | 0x40, 200, sub[0][4], None))
            sub.append((np.zeros(0, np.uint8), t[:10].copy(), 0x100, 200, sub[0][4], None))
            sub.append((q[:30].copy(), np.zeros(0, np.uint8), 0x100, 200, sub[0][4], None))
        mat = ol.simple_mat(a, b, 1)
        any_junc = any(c[5] is not None for c in sub)
        ez, cigs = kernels.ksw_exts2_batch([c[0] for c in sub], [c[1] for c in sub], mat, go, ge, go2, noncan, jb, np.array([c[3] for c in sub]),
                                           np.array([c[2] for c in sub]), juncs=[c[5] for c in sub] if any_junc else None)
        for i, (qq, tt, flag, zdrop, _, junc) in enumerate(sub):
            e0, c0 = ol.oracle_exts2(qq, tt, mat, go, ge, go2, noncan, zdrop, jb, flag, junc=junc)
            assert np.array_equal(e0, ez[i]), (k, i, len(qq), len(tt), hex(flag), e0, ez[i])
            assert np.array_equal(c0, cigs[i]), (k, i, len(qq), len(tt), hex(flag))
            n_intron += int((c0 & 0xf == 3).any())
    assert n_intron > 20
