"""One warm-up search + one measured search (device-resident) for ncu / quick timing."""
import argparse
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pq_huffman_b200 import knn, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--k", type=int, default=17)
ap.add_argument("--gen", default="sift_like")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--rows", type=int, default=0, help="search only the first ROWS query rows")
args = ap.parse_args()
x = getattr(synth, args.gen)(args.n, args.d)
knn.init([0])
dev = torch.device("cuda", 0)
d_x = torch.from_numpy(x).to(dev)
n, d, k = args.n, args.d, args.k
rows = args.rows or n
d_a = torch.empty((n, k), dtype=torch.int32, device=dev)
d_d = torch.empty((n, k), dtype=torch.float32, device=dev)
wb = knn.workspace_bytes(rows, n, d, k)
d_w = torch.empty(wb, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream(device=dev)
import time
for rep in range(args.reps):
    torch.cuda.synchronize(); _t0 = time.perf_counter()
    knn.search_device(0, n, n, d, k, d_x.data_ptr(), d_x.data_ptr(), True, 0, rows, d_a.data_ptr(),
                      d_d.data_ptr(), d_w.data_ptr(), wb, s.cuda_stream)
    torch.cuda.synchronize()
    st = knn.last_stats()
    st['wall_ms'] = (time.perf_counter() - _t0) * 1e3
    print({key: (round(v, 3) if isinstance(v, float) else v) for key, v in st.items()}, flush=True)
