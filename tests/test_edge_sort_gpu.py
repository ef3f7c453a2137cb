"""the MST builder's two big edge sorts on the device.

pqh_sort_edges must give what glibc's (stable, merge-sort) qsort gives with the reference's
comparators (mst.c:15-32) wherever those are an order, and must refuse -- leaving the records
alone -- where mst_dist_comparator is not one.  The end-to-end check runs the reference's OWN
mst_builder twice on the same kNN graph, once as built from its sources and once with its qsort
calls renamed to pqh_qsort_edges at compile time (oracle/Makefile, mst_builder_gpu): the tree file
must come out byte for byte the same.
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from pq_huffman_b200 import knn, synth, vecs_io
from tests import oracle_ffi

pytestmark = pytest.mark.gpu

REF = oracle_ffi.REF_DIR


def _edges(n, seed, integer=True):
    rng = np.random.default_rng(seed)
    e = np.empty(n, dtype=knn.EDGE)
    e["source"] = rng.integers(0, max(2, n // 16), n)
    e["target"] = rng.integers(0, 1 << 30, n)
    if integer:   # squared distances of uint8-valued data: integers with many ties
        e["dist"] = rng.integers(0, 5000, n).astype(np.float32)
    else:
        e["dist"] = (rng.random(n, dtype=np.float32) * 4 + 0.01).astype(np.float32)
    return e


@pytest.mark.parametrize("n", [2, 255, 70_001, 1_000_003])
@pytest.mark.parametrize("integer", [True, False])
def test_sort_by_distance_is_the_stable_order(n, integer):
    e = _edges(n, seed=n, integer=integer)
    if not integer:   # keep different distances at least 1e-9 apart so the comparator is an order
        e["dist"] = np.round(e["dist"] * 1e5).astype(np.float32) / np.float32(1e5) + np.float32(1.0)
    want = e[np.argsort(e["dist"], kind="stable")]
    got = e.copy()
    assert knn.sort_edges(got, knn.SORT_BY_DIST)
    assert got.tobytes() == want.tobytes()


def test_sort_by_source_is_the_stable_order():
    e = _edges(300_000, seed=5)
    want = e[np.argsort(e["source"], kind="stable")]
    got = e.copy()
    assert knn.sort_edges(got, knn.SORT_BY_SOURCE)
    assert got.tobytes() == want.tobytes()


def test_signed_zero_and_duplicates_keep_input_order():
    e = _edges(100_000, seed=6)
    e["dist"][::3] = 0.0
    e["dist"][1::7] = -0.0   # the comparator cannot tell -0 from +0: they stay in input order
    order = np.argsort(np.where(e["dist"] == 0, np.float32(0), e["dist"]), kind="stable")
    got = e.copy()
    assert knn.sort_edges(got, knn.SORT_BY_DIST)
    assert got.tobytes() == e[order].tobytes()


@pytest.mark.parametrize("poison", ["nan", "close"])
def test_refuses_when_the_comparator_is_not_an_order(poison):
    e = _edges(80_000, seed=7)
    if poison == "nan":
        e["dist"][12345] = np.nan
    else:            # two different distances 1e-12 apart: "equal" to the comparator, not to the bits
        e["dist"][100] = np.float32(1e-12)
        e["dist"][70_000] = np.float32(2e-12)
    before = e.tobytes()
    assert knn.sort_edges(e, knn.SORT_BY_DIST) is False
    assert e.tobytes() == before
    assert knn.sort_edges(e, knn.SORT_BY_SOURCE)      # the source order is always an order


@pytest.mark.skipif(not all(os.path.exists(os.path.join(REF, b)) for b in
                            ("mst_builder", "mst_builder_gpu")),
                    reason="oracle/_ref binaries not built")
def test_reference_mst_builder_with_device_sorts_writes_the_same_tree(tmp_path):
    """200k x 64 SIFT-shaped rows, k = 16: 3.2M edge records by distance, ~400k forest edges by
    source -- both above PQH_EDGE_SORT_MIN, so both sorts of mst.c run on the device in
    mst_builder_gpu and in libc in mst_builder."""
    n, d, k, m = 200_000, 64, 16, 8
    x = synth.sift_like(n, d)
    a, dist = knn.knn_full_thread(x, None, k + 1)
    out = {}
    pq_tpl = str(tmp_path) + "/pq_"
    vecs_io.write_light(pq_tpl + "pq_indices.bvecsl", synth.pq_codes_stand_in(x, m))
    for binary in ("mst_builder", "mst_builder_gpu"):
        tdir = tmp_path / binary
        tdir.mkdir()
        tpl = str(tdir) + "/"
        vecs_io.write_light(tpl + "nn_indices.ivecsl", np.ascontiguousarray(a[:, 1:]))
        vecs_io.write_light(tpl + "nn_dist.fvecsl", np.ascontiguousarray(dist[:, 1:]))
        p = subprocess.run([os.path.join(REF, binary), tpl, tpl, str(k), "--pq-template", pq_tpl],
                           cwd=str(tmp_path), capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-500:]
        out[binary] = hashlib.sha256(open(tpl + "mst.tree", "rb").read()).hexdigest()
    assert out["mst_builder"] == out["mst_builder_gpu"]
