"""One warm-up and two timed device edge sorts of 16M records (for ncu / quick timing)."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

from pq_huffman_b200 import knn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
rng = np.random.default_rng(0)
e = np.empty(n, dtype=knn.EDGE)
e["source"] = np.repeat(np.arange(n // 16, dtype=np.int32), 16)[:n]
e["target"] = rng.integers(0, n // 16, n)
e["dist"] = rng.integers(1000, 200000, n).astype(np.float32)
for rep in range(3):
    s = e.copy()
    t0 = time.perf_counter()
    ok = knn.sort_edges(s, knn.SORT_BY_DIST)
    print("sort_edges %d records: %.1f ms taken=%s" % (n, (time.perf_counter() - t0) * 1e3, ok), flush=True)
assert np.all(np.diff(s["dist"]) >= 0)
