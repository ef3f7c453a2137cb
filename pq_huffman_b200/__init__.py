"""pq_huffman_b200 -- B200-native exact kNN for pq_huffman's compute_nn_fast path.

The product is ``libpqh_knn.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/pqh_knn.h``).  This package is the thin Python host-side mirror used by
the tests and ``bench.py``: a ctypes binding of the C ABI plus the reference's
file formats and seeded synthetic data.  It never computes neighbours itself and
has no CPU fallback: importing :mod:`pq_huffman_b200.knn` fails loudly when the
CUDA library is missing.
"""

__all__ = ["knn", "vecs_io", "synth"]
