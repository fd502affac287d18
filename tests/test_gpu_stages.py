"""GPU parity for the sketch / sort / chain kernels (through the C ABI) against the CPU oracle."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from test_oracle_vs_ref import make_anchors

pytestmark = pytest.mark.gpu
TESTS_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT_DIR = os.path.dirname(TESTS_DIR)


def _rand_seq(rng, n):
    return bytes(b"ACGT"[c] for c in rng.integers(0, 4, size=n, dtype=np.uint8))


def _seq_set(rng, kmers, k):
    base = _rand_seq(rng, 40000)
    seqs = [base, base[:2000], base[5000:5130], base[:14], b"ACGTACGTAC", b""]
    s2 = bytearray(base[:9000])
    for p in range(500, 8000, 900):
        L = int(rng.integers(1, 80))
        s2[p:p + L] = b"N" * L
    seqs.append(bytes(s2))
    unit = _rand_seq(rng, 7)
    seqs.append(base[:300] + unit * 60 + base[300:900] + b"A" * 120 + base[900:1500] + (b"AC" * 70) + base[1500:2000] + b"acgtn" * 30 + base[2000:2500])
    if len(kmers):
        planted = bytearray(base[:6000])
        for j in range(min(40, len(kmers))):
            km = int(kmers[j])
            st = "".join("ACGT"[(km >> (2 * (k - 1 - i))) & 3] for i in range(k)).encode()
            planted[100 + j * 110: 100 + j * 110 + k] = st
        seqs.append(bytes(planted))
    # many windows of one read, as in stage 1
    seqs += [base[i:i + 2000] for i in range(0, 20000, 2000)]
    return seqs


@pytest.mark.parametrize("n_k,k,w", [(0, 15, 50), (255, 15, 50), (5000, 15, 50), (300, 19, 50), (255, 15, 10), (0, 16, 20), (0, 6, 20)])
def test_sketch_matches_oracle(n_k, k, w):
    from winnowmap_b200 import kernels
    rng = np.random.default_rng(1000 + n_k + k + w)
    kmers = rng.integers(0, 1 << (2 * k), size=n_k, dtype=np.uint64)
    ob = ol.OracleBloom(kmers)
    gb = kernels.Bloom(kmers)
    assert gb.bits() == ob.bits()
    assert np.array_equal(gb.table(), ob.table())
    seqs = _seq_set(rng, kmers, k)
    if k % 2 == 0:
        seqs.append(seqs[0][:1000] + b"ACGTACGTACGTACGTAATT" * 5 + seqs[0][1000:3000])
    rids = np.arange(len(seqs)) % 3
    got = kernels.sketch_batch(gb, seqs, w, k, rids)
    for i, s in enumerate(seqs):
        exp = ol.oracle_sketch(s, w, k, int(rids[i]), ob) if len(s) else np.zeros((0, 2), np.uint64)
        assert np.array_equal(exp, got[i]), (i, len(s), len(exp), len(got[i]))


def test_sort_matches_oracle_with_ties():
    from winnowmap_b200 import kernels
    rng = np.random.default_rng(3)
    arrays = []
    for n in [0, 1, 2, 63, 64, 65, 200, 1000, 5000, 70000]:
        for key_bits in [3, 12, 28, 64]:
            hi = (1 << key_bits) - 1
            x = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
            if key_bits == 64 and n:
                x[::3] = x[0]
            arrays.append(np.stack([x, np.arange(n, dtype=np.uint64)], axis=1))
    got = kernels.radix_sort_128x_batch(arrays)
    for a, g in zip(arrays, got):
        assert np.array_equal(ol.oracle_sort128(a), g), len(a)


@pytest.mark.parametrize("seed", range(3))
def test_chain_matches_oracle(seed):
    from winnowmap_b200 import kernels
    rng = np.random.default_rng(500 + seed)
    arrays = [make_anchors(rng, n, repeats=bool((seed + i) & 1)) for i, n in enumerate([0, 1, 3, 10, 40, 100, 100, 700, 3000, 9000] * 3)]
    for prm in [dict(max_dist_x=5000, min_dist_x=1000, max_dist_y=5000, bw=500), dict(max_dist_x=16000, min_dist_x=1000, max_dist_y=16000, bw=2000),
                dict(max_dist_x=5000, min_dist_x=50, max_dist_y=5000, bw=500, max_iter=20, max_skip=3)]:
        got = kernels.chain_dp_batch(arrays, **prm)
        for a, (u, b) in zip(arrays, got):
            ue, be = ol.oracle_chain(a, prm["max_dist_x"], prm["min_dist_x"], prm["max_dist_y"], prm["bw"],
                                     max_skip=prm.get("max_skip", 25), max_iter=prm.get("max_iter", 5000))
            assert np.array_equal(ue, u), (len(a), len(ue), len(u))
            assert np.array_equal(be, b), len(a)


@pytest.mark.xfail(strict=False, reason="experimental formulation (WM_CHAIN_DENSE=1), verified on the CPU software warp "
                                        "(tests/test_kernel_emulation.py), not yet run on hardware")
def test_chain_dense_formulation_matches_oracle(tmp_path):
    """The dense-candidate forward pass (csrc/chain_dev.cuh) through the C ABI, in a child process: the switch is read
    once per process, and a CUDA fault must not take the test session with it."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import oracle_lib as ol
from winnowmap_b200 import kernels
from test_oracle_vs_ref import make_anchors
rng = np.random.default_rng(77)
arrays = [make_anchors(rng, n, repeats=bool(i & 1)) for i, n in enumerate([0, 5, 100, 1500, 3000, 9000, 9000])]
for prm in [dict(max_dist_x=5000, min_dist_x=1000, max_dist_y=5000, bw=500), dict(max_dist_x=5000, min_dist_x=50, max_dist_y=5000, bw=500, max_iter=20, max_skip=3)]:
    got = kernels.chain_dp_batch(arrays, **prm)
    for a, (u, b) in zip(arrays, got):
        ue, be = ol.oracle_chain(a, prm["max_dist_x"], prm["min_dist_x"], prm["max_dist_y"], prm["bw"], max_skip=prm.get("max_skip", 25), max_iter=prm.get("max_iter", 5000))
        assert np.array_equal(ue, u) and np.array_equal(be, b), len(a)
print("dense ok")
''' % (ROOT_DIR, TESTS_DIR)
    env = dict(os.environ, WM_CHAIN_DENSE="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert out.returncode == 0 and b"dense ok" in out.stdout, out.stdout[-2000:]
