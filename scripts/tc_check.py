"""GPU bring-up check for the tensor-core path: parity vs oracle + how rows were resolved."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pq_huffman_b200 import knn, synth  # noqa: E402
from tests import oracle_ffi  # noqa: E402

orc = oracle_ffi.Oracle()
cases = [("sift", 5000, 128, 17), ("sift", 20000, 128, 17), ("deep", 8000, 96, 33),
         ("uniform", 4096, 16, 9), ("gist", 3000, 144, 17), ("sift", 3000, 100, 5),
         ("gist", 3000, 960, 17), ("gist", 5000, 200, 9), ("uniform", 4000, 272, 17),
         ("deep", 20000, 96, 33)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
gens = {"sift": synth.sift_like, "gist": synth.gist_like, "deep": synth.deep_like,
        "uniform": synth.uniform}
for gen, n, d, k in cases:
    x = gens[gen](n, d, seed=n + d)
    knn.set_algo(knn.ALGO_AUTO)
    t = time.time()
    a, dist = knn.knn_full_thread(x, None, k)
    dt = time.time() - t
    st = knn.last_stats()
    ea, ed = orc.knn(x, None, k)
    bad_rows = int((a != ea).any(axis=1).sum())
    bad_d = int((dist.view(np.uint32) != ed.view(np.uint32)).sum())
    print(f"{gen} {n}x{d} k={k}: rows_tensor={st['rows_tensor']} rows_exact={st['rows_exact']} "
          f"launches={st['launches']} kernels={st['ms_kernels']:.2f}ms wall={dt*1e3:.1f}ms "
          f"bad_rows={bad_rows} bad_dist={bad_d}", flush=True)
    if bad_rows:
        r = int(np.nonzero((a != ea).any(axis=1))[0][0])
        print("  first bad row", r, a[r], ea[r], dist[r], ed[r])
