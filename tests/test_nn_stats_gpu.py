"""SURVEY 8(f4), second half: the estimator's joint code histograms on the device
(pqh_nn_joint_histograms) against the restatement of collect_stats_at_kth_nn that
tests/test_oracle_pins.py pins to the reference's own function.  Counts are integers: bit-exact."""
import numpy as np
import pytest

from pq_huffman_b200 import knn
from tests import ref_semantics

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,num_nn,k_star", [(60_000, 8, 6, 256), (1001, 3, 2, 16), (1, 1, 1, 256)])
def test_joint_histograms_match_the_reference_semantics(n, m, num_nn, k_star):
    rng = np.random.default_rng(n + m)
    pq = rng.integers(0, k_star, (n, m)).astype(np.uint8)
    nn = rng.integers(0, n, (n, num_nn)).astype(np.uint32)
    nn[rng.random((n, num_nn)) < 0.1] = 0xFFFFFFFF
    stats, known = knn.nn_joint_histograms(pq, nn, k_star=k_star)
    assert stats.shape == (num_nn, m, k_star, k_star)
    for j in range(num_nn):
        exp, exp_known = ref_semantics.joint_histograms(pq, nn, k_star, j)
        assert known[j] == exp_known
        assert np.array_equal(stats[j].view(np.uint64), exp.view(np.uint64))
    # a sub-range of ranks gives the same tables
    if num_nn > 2:
        part, pk = knn.nn_joint_histograms(pq, nn, k_star=k_star, nn_first=1, nn_count=2)
        assert np.array_equal(part, stats[1:3]) and np.array_equal(pk, known[1:3])


def test_bad_neighbour_index_is_reported():
    pq = np.zeros((100, 2), dtype=np.uint8)
    nn = np.zeros((100, 2), dtype=np.uint32)
    nn[7, 1] = 100          # not a row
    with pytest.raises(knn.PqhError):
        knn.nn_joint_histograms(pq, nn)
