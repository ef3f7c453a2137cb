"""CPU tests of the oracle itself (oracle/knn_oracle.c) -- run with -m "not gpu"."""
import numpy as np
import pytest

from pq_huffman_b200 import synth
from tests import oracle_ffi


def scalar_rows(base, queries, k, self_join):
    """Pure-Python/numpy restatement of the restatement, for small cases:
    float32 sequential sum of squared differences, (dist, id) order."""
    nq, d = queries.shape
    acc = np.zeros((nq, base.shape[0]), dtype=np.float32)
    for i in range(d):
        t = (queries[:, i][:, None] - base[:, i][None, :]).astype(np.float32)
        acc = (acc + (t * t).astype(np.float32)).astype(np.float32)
    ids = np.empty((nq, k), dtype=np.int32)
    dis = np.empty((nq, k), dtype=np.float32)
    for q in range(nq):
        order = sorted(range(base.shape[0]), key=lambda j: (acc[q, j], j))
        if self_join:
            order.remove(q)
            order = [q] + order
        for c in range(k):
            if c < len(order):
                ids[q, c] = order[c]
                dis[q, c] = 0.0 if (self_join and c == 0) else acc[q, order[c]]
            else:
                ids[q, c] = -1
                dis[q, c] = np.inf
    return ids, dis


@pytest.mark.parametrize("n,d,k", [(300, 7, 5), (257, 33, 17), (64, 128, 64), (100, 1, 3)])
def test_oracle_matches_scalar_python(oracle, n, d, k):
    x = synth.uniform(n, d, seed=n + d)
    a, dist = oracle.knn(x, None, k)
    ra, rd = scalar_rows(x, x, k, True)
    assert np.array_equal(a, ra)
    assert np.array_equal(dist.view(np.uint32), rd.view(np.uint32))


def test_oracle_distinct_queries(oracle):
    b = synth.uniform(500, 24, seed=1)
    v = synth.uniform(77, 24, seed=2)
    a, dist = oracle.knn(b, v, 9)
    ra, rd = scalar_rows(b, v, 9, False)
    assert np.array_equal(a, ra)
    assert np.array_equal(dist.view(np.uint32), rd.view(np.uint32))


def test_oracle_threads_agree(oracle):
    x = synth.sift_like(3000, 32, seed=5)
    a1, d1 = oracle.knn(x, None, 17, n_thread=1)
    a8, d8 = oracle.knn(x, None, 17, n_thread=8)
    assert np.array_equal(a1, a8) and np.array_equal(d1.view(np.uint32), d8.view(np.uint32))


def test_oracle_vs_float64_sets(oracle):
    """On integer-valued (SIFT-like) data FP32 is exact, so the oracle must equal float64."""
    x = synth.sift_like(4000, 128, seed=11)
    a, dist = oracle.knn(x, None, 17)
    ia, dd = oracle_ffi.brute_force_f64(x, x, 17, True)
    assert np.array_equal(a, ia)
    dd[:, 0] = 0.0
    assert np.array_equal(dist.astype(np.float64), dd)


def test_oracle_ties_and_duplicates(oracle):
    x = synth.with_duplicates_and_ties(synth.sift_like(1500, 16, seed=3))
    a, dist = oracle.knn(x, None, 9)
    assert np.array_equal(a[:, 0], np.arange(1500))          # self first, even with duplicates
    assert np.all(dist[:, 0] == 0.0)
    assert np.all(np.diff(dist, axis=1) >= 0)                # ascending
    same = np.diff(dist, axis=1) == 0
    assert same.any()                                         # planted ties exist
    # equal distances are ordered by position (column 0 is the forced self)
    assert np.all(np.diff(a, axis=1)[:, 1:][same[:, 1:]] > 0)
    for r in range(1500):
        assert len(set(a[r])) == 9


def test_oracle_k_larger_than_nb(oracle):
    x = synth.uniform(5, 4, seed=2)
    a, dist = oracle.knn(x, None, 8)
    assert np.all(a[:, 5:] == -1) and np.all(np.isinf(dist[:, 5:]))
    assert np.all(a[:, :5] >= 0)


def test_oracle_rows_slice(oracle):
    x = synth.uniform(700, 12, seed=8)
    a, dist = oracle.knn(x, None, 6)
    sa, sd = oracle.knn_rows(x, None, 6, 130, 391)
    assert np.array_equal(sa, a[130:391]) and np.array_equal(sd, dist[130:391])


def test_blas_baseline_agrees_on_integer_data():
    """The timed CPU arm must solve the same problem (ids equal where FP32 is exact)."""
    blas = oracle_ffi.BlasBaseline()
    orc = oracle_ffi.Oracle()
    x = synth.sift_like(3000, 64, seed=21)
    a, dist = orc.knn(x, None, 11)
    ba, bd = blas.knn_rows(x, None, 11, 0, 3000)
    # integer data: expansion is exact too, only the order among equal distances may differ
    assert np.array_equal(dist, bd)
    assert (a != ba).mean() < 0.01
