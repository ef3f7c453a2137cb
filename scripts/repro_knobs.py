"""Run the knob test cases outside pytest (stderr visible): python scripts/repro_knobs.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from pq_huffman_b200 import knn, synth
from tests import oracle_ffi

orc = oracle_ffi.Oracle()
knn.init([0])
cases = {
    "sift_self": (synth.with_duplicates_and_ties(synth.sift_like(40000, 128, seed=77), n_dups=6, n_ties=6), None, 17),
    "gist_self": (synth.gist_like(33000, 96, seed=78), None, 9),
    "deep_bv": (synth.deep_like(34000, 96, seed=79), synth.deep_like(700, 96, seed=80), 33),
}
for name, (b, v, k) in cases.items():
    print("case", name, flush=True)
    a, d = knn.knn_full_thread(b, v, k)
    print("  stats", {kk: vv for kk, vv in knn.last_stats().items() if kk.startswith("rows") or kk == "inexact_operands"}, flush=True)
    ea, ed = orc.knn(b, v, k)
    bad = np.nonzero((a != ea).any(axis=1))[0]
    print("  rows differing:", bad.size, bad[:8], "dist bits equal:", np.array_equal(d.view(np.uint32), ed.view(np.uint32)), flush=True)
