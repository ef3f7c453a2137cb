"""Wall time of the reference's mst_builder on a 1M x 16 graph: libc qsort vs the device sorts
(oracle/_ref/mst_builder vs mst_builder_gpu, the same sources, qsort renamed at compile time)."""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from pq_huffman_b200 import knn, synth, vecs_io  # noqa: E402

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
n, d, k, m = 1_000_000, 128, 16, 8
x = synth.sift_like(n, d)
a, dist = knn.knn_full_thread(x, None, k + 1)
e = np.empty(n * k, dtype=knn.EDGE)
e["source"] = np.repeat(np.arange(n, dtype=np.int32), k)
e["target"] = a[:, 1:].reshape(-1)
e["dist"] = dist[:, 1:].reshape(-1)
for rep in range(3):
    s = e.copy()
    t0 = time.perf_counter()
    ok = knn.sort_edges(s, knn.SORT_BY_DIST)
    print("pqh_sort_edges(%d records by distance): %.1f ms, taken=%s" % (n * k, (time.perf_counter() - t0) * 1e3, ok), flush=True)
t0 = time.perf_counter()
want = e[np.argsort(e["dist"], kind="stable")]
print("numpy stable argsort + gather on the host: %.1f ms, identical=%s" % ((time.perf_counter() - t0) * 1e3, want.tobytes() == s.tobytes()))
knn.shutdown() if hasattr(knn, "shutdown") else None
with tempfile.TemporaryDirectory() as tmp:
    pq_tpl = tmp + "/pq_"
    vecs_io.write_light(pq_tpl + "pq_indices.bvecsl", synth.pq_codes_stand_in(x, m))
    sha = {}
    for binary in ("mst_builder", "mst_builder_gpu", "mst_builder", "mst_builder_gpu"):
        tpl = tmp + "/" + binary + "_"
        vecs_io.write_light(tpl + "nn_indices.ivecsl", np.ascontiguousarray(a[:, 1:]))
        vecs_io.write_light(tpl + "nn_dist.fvecsl", np.ascontiguousarray(dist[:, 1:]))
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(REF, binary), tpl, tpl, str(k), "--pq-template", pq_tpl],
                           cwd=tmp, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr[-400:]
        sha[binary] = hashlib.sha256(open(tpl + "mst.tree", "rb").read()).hexdigest()
        print("%-16s wall %.2f s" % (binary, dt), flush=True)
    print("mst.tree identical:", sha["mst_builder"] == sha["mst_builder_gpu"])
