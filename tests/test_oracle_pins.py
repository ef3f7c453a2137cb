"""Pin the oracle against the reference's OWN compiled sources (oracle/_ref/libref_cnf.so).

The kNN numbers themselves cannot be pinned (Yael is absent: "parity unpinned",
see oracle/knn_oracle.c); what the reference does define in-tree is pinned here:
the distance (get_real_dist, compute_nn_fast.c:304-311) and the merge order
(fast_nn_heap_push / fast_nn_heap_sort, fast_nn_temp_file.c:11-63), plus the
whole CLI linked against the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pq_huffman_b200 import synth, vecs_io
from tests import oracle_ffi, ref_semantics


def test_get_real_dist_bit_exact(oracle, refcnf):
    rng = np.random.default_rng(7)
    for d in (1, 3, 16, 96, 128, 960, 1000):
        for scale in (1.0, 255.0, 1e-3):
            a = (rng.normal(size=d) * scale).astype(np.float32)
            b = (rng.normal(size=d) * scale).astype(np.float32)
            ref = C.c_float(refcnf.get_real_dist(a.ctypes.data, b.ctypes.data, d)).value
            mine = oracle.dist(a, b)
            assert np.float32(ref).view(np.uint32) == np.float32(mine).view(np.uint32)


def test_knn_distances_equal_get_real_dist(oracle, refcnf):
    x = synth.gist_like(400, 96, seed=3)
    a, dist = oracle.knn(x, None, 9)
    for r in range(0, 400, 37):
        for c in range(1, 9):
            ref = refcnf.get_real_dist(x[r].ctypes.data, x[a[r, c]].ctypes.data, 96)
            assert np.float32(ref).view(np.uint32) == dist[r, c].view(np.uint32)


class NNItem(C.Structure):
    _fields_ = [("index", C.c_uint), ("dist", C.c_float)]


def test_heap_semantics_match_reference(refcnf):
    rng = np.random.default_rng(5)
    for num_nn in (1, 2, 5, 16):
        for trial in range(20):
            heap_c = (NNItem * num_nn)()
            heap_py = []
            for i in range(num_nn):
                heap_c[i].index = 0xFFFFFFFF
                heap_c[i].dist = np.inf
                heap_py.append([0xFFFFFFFF, np.float32(np.inf)])
            n_items = int(rng.integers(1, 60))
            ids = rng.integers(0, 25, size=n_items)            # repeats -> duplicate rejection
            dists = rng.integers(0, 12, size=n_items).astype(np.float32)  # ties
            for i, dv in zip(ids, dists):
                item = NNItem(int(i), float(dv))
                refcnf.fast_nn_heap_push(heap_c, C.byref(item), num_nn)
                ref_semantics.heap_push(heap_py, [int(i), np.float32(dv)], num_nn)
            assert [(h.index, h.dist) for h in heap_c] == [(h[0], float(h[1])) for h in heap_py]
            oi = (C.c_uint * num_nn)()
            od = (C.c_float * num_nn)()
            refcnf.fast_nn_heap_sort(heap_c, num_nn, oi, od)
            pi, pd = ref_semantics.heap_sort(heap_py, num_nn)
            assert list(oi) == pi and list(od) == [float(v) for v in pd]


@pytest.mark.parametrize("gen,n,d,k", [("sift", 1000, 32, 16), ("gist", 600, 48, 5)])
def test_reference_cli_with_oracle(tmp_path, oracle, gen, n, d, k):
    """Reference CLI (unmodified objects) + oracle == Python restatement of its merge."""
    cli = os.path.join(oracle_ffi.REF_DIR, "compute_nn_fast_oracle")
    if not os.path.exists(cli):
        pytest.skip("oracle/_ref not built")
    x = synth.sift_like(n, d, seed=13) if gen == "sift" else synth.gist_like(n, d, seed=13)
    x = synth.with_duplicates_and_ties(x) if gen == "sift" else x
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    tpl = str(tmp_path / "out_")
    subprocess.check_call([cli, inp, tpl, str(k), "--num-dims", "0", "--temp-file",
                           str(tmp_path / "tmp.vecs"), "--num-threads", "4"],
                          stdout=subprocess.DEVNULL)
    got_i, got_d = vecs_io.read_nn_result(tpl)
    a, dist = oracle.knn(x, None, k + 1)
    exp_i, exp_d = ref_semantics.single_block_result(a, dist, k)
    assert np.array_equal(got_i, exp_i)
    assert np.array_equal(got_d.view(np.uint32), exp_d.view(np.uint32))


def test_histogram_restatement_matches_the_reference_function():
    """tests/ref_semantics.joint_histograms (the checker of pqh_nn_joint_histograms) against the
    reference's own collect_stats_at_kth_nn, which is static in nn_huffman_estimator.c and
    reached through oracle/ref_est_shim.c (the source included where it lies)."""
    import ctypes as C
    path = os.path.join(oracle_ffi.REF_DIR, "libref_est.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_est.so not built (needs /root/reference)")
    lib = C.CDLL(path)
    lib.ref_collect_stats.restype = C.c_longlong
    lib.ref_collect_stats.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(11)
    n, m, num_nn, k_star = 5000, 4, 6, 256
    pq = rng.integers(0, 40, (n, m)).astype(np.uint8)         # few codes: well filled cells
    nn = rng.integers(0, n, (n, num_nn)).astype(np.uint32)
    nn[rng.random((n, num_nn)) < 0.05] = 0xFFFFFFFF            # unknown neighbours
    for j in (0, 3, 5):
        ref = np.empty((m, k_star, k_star), dtype=np.float64)
        known = lib.ref_collect_stats(n, m, k_star, num_nn, pq.ctypes.data, nn.ctypes.data,
                                      ref.ctypes.data, j)
        got, got_known = ref_semantics.joint_histograms(pq, nn, k_star, j)
        assert known == got_known
        assert np.array_equal(ref.view(np.uint64), got.view(np.uint64))
