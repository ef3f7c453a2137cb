"""ctypes access to the CPU checkers under oracle/ (tests and bench only)."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_DIR = os.path.join(ORACLE_DIR, "lib")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


def ensure_built():
    """Build oracle/lib (always possible: gcc only) if it is missing."""
    if not os.path.exists(os.path.join(LIB_DIR, "libknn_oracle.so")) or not os.path.exists(
            os.path.join(LIB_DIR, "libknn_blas.so")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "lib"], stdout=subprocess.DEVNULL)


def find_openblas():
    import scipy
    pats = [os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs",
                         "libscipy_openblas-*.so")]
    for p in pats:
        hits = sorted(glob.glob(p))
        if hits:
            return hits[0]
    raise RuntimeError("scipy's bundled OpenBLAS (LP64) not found")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    """Scalar-FP32 parity oracle (oracle/knn_oracle.c)."""

    def __init__(self):
        ensure_built()
        self.lib = C.CDLL(os.path.join(LIB_DIR, "libknn_oracle.so"))
        vp = C.c_void_p
        self.lib.knn_full_thread.argtypes = [C.c_int] * 5 + [vp, vp, vp, vp, vp, C.c_int]
        self.lib.knn_full_thread.restype = None
        self.lib.pqh_oracle_knn_rows.argtypes = [C.c_int] * 4 + [vp, vp, C.c_int, C.c_int, vp, vp,
                                                                C.c_int]
        self.lib.pqh_oracle_knn_rows.restype = None
        self.lib.pqh_oracle_dist.argtypes = [vp, vp, C.c_int]
        self.lib.pqh_oracle_dist.restype = C.c_float

    def knn(self, base, queries=None, k=1, n_thread=None):
        b = _f32(base)
        v = b if queries is None or queries is base else _f32(queries)
        nq, d = v.shape
        assign = np.empty((nq, k), dtype=np.int32)
        dis = np.empty((nq, k), dtype=np.float32)
        nt = n_thread or os.cpu_count() or 1
        self.lib.knn_full_thread(2, nq, b.shape[0], d, k, b.ctypes.data, v.ctypes.data, None,
                                 assign.ctypes.data, dis.ctypes.data, nt)
        return assign, dis

    def knn_rows(self, base, queries, k, q_begin, q_end, n_thread=None):
        """Rows [q_begin,q_end); returns arrays of (q_end-q_begin) rows."""
        b = _f32(base)
        v = b if queries is None or queries is base else _f32(queries)
        nq, d = v.shape
        assign = np.empty((nq, k), dtype=np.int32)
        dis = np.empty((nq, k), dtype=np.float32)
        nt = n_thread or os.cpu_count() or 1
        self.lib.pqh_oracle_knn_rows(nq, b.shape[0], d, k, b.ctypes.data, v.ctypes.data, q_begin,
                                     q_end, assign.ctypes.data, dis.ctypes.data, nt)
        return assign[q_begin:q_end].copy(), dis[q_begin:q_end].copy()

    def dist(self, a, b):
        a = _f32(a)
        b = _f32(b)
        return float(self.lib.pqh_oracle_dist(a.ctypes.data, b.ctypes.data, a.size))


class BlasBaseline:
    """Yael-shaped sgemm baseline (oracle/knn_blas.c), the timed CPU arm."""

    def __init__(self):
        ensure_built()
        os.environ.setdefault("PQH_OPENBLAS", find_openblas())
        self.lib = C.CDLL(os.path.join(LIB_DIR, "libknn_blas.so"))
        vp = C.c_void_p
        self.lib.knn_full_thread.argtypes = [C.c_int] * 5 + [vp, vp, vp, vp, vp, C.c_int]
        self.lib.knn_full_thread.restype = None
        self.lib.pqh_blas_knn_rows.argtypes = [C.c_int] * 4 + [vp, vp, C.c_int, C.c_int, vp, vp,
                                                              C.c_int]
        self.lib.pqh_blas_knn_rows.restype = None

    def knn_rows(self, base, queries, k, q_begin, q_end, n_thread=None):
        b = _f32(base)
        v = b if queries is None or queries is base else _f32(queries)
        nq, d = v.shape
        assign = np.empty((nq, k), dtype=np.int32)
        dis = np.empty((nq, k), dtype=np.float32)
        nt = n_thread or os.cpu_count() or 1
        self.lib.pqh_blas_knn_rows(nq, b.shape[0], d, k, b.ctypes.data, v.ctypes.data, q_begin,
                                   q_end, assign.ctypes.data, dis.ctypes.data, nt)
        return assign[q_begin:q_end].copy(), dis[q_begin:q_end].copy()


def load_ref_cnf():
    """The reference's own compute_nn_fast objects as a library (oracle/_ref), or None."""
    path = os.path.join(REF_DIR, "libref_cnf.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.get_real_dist.argtypes = [vp, vp, C.c_int]
    lib.get_real_dist.restype = C.c_float
    lib.fast_nn_heap_push.argtypes = [vp, vp, C.c_int]
    lib.fast_nn_heap_push.restype = None
    lib.fast_nn_heap_sort.argtypes = [vp, C.c_int, vp, vp]
    lib.fast_nn_heap_sort.restype = None
    return lib


def brute_force_f64(base, queries, k, self_join):
    """float64 numpy brute force: returns the k nearest ids per row ordered by (dist, id)."""
    b = np.asarray(base, dtype=np.float64)
    v = np.asarray(queries, dtype=np.float64)
    d2 = ((v * v).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * v @ b.T)
    if self_join:
        d2[np.arange(v.shape[0]), np.arange(v.shape[0])] = -1.0
    ids = np.lexsort((np.broadcast_to(np.arange(b.shape[0]), d2.shape), d2), axis=1)[:, :k]
    return ids.astype(np.int32), np.take_along_axis(d2, ids, axis=1)
