"""End-to-end timing of pqh_knn_search_rows with pageable host buffers (H2D / kernels / D2H)."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from pq_huffman_b200 import knn, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
x = synth.sift_like(n, 128)
knn.init([0])
for rep in range(4):
    t0 = time.perf_counter()
    a, d = knn.knn_full_thread(x, None, 17)
    wall = (time.perf_counter() - t0) * 1e3
    st = knn.last_stats()
    if rep:
        print("  wall %.1f ms | h2d %.2f kernels %.2f d2h %.2f" % (wall, st["ms_h2d"], st["ms_kernels"], st["ms_d2h"]), flush=True)
