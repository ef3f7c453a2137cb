"""Host-side mirror of the reference's kNN call for tests and bench.py.

``knn_full_thread`` here has the argument meaning of Yael's function as the
reference calls it (compute_nn_fast.c:478-481) and of ``ynumpy.knn`` in
compute_nn.py:22: arrays in, (ids, distances) out, squared L2, ascending, k per
row.  All arithmetic happens inside libpqh_knn.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib

ALGO_AUTO = 0
ALGO_EXACT = 1


class PqhError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise PqhError(f"libpqh_knn error {rc}: {_lib.load().pqh_last_error().decode()}")


def init(devices=None):
    """Create the device context (pqh_knn_init)."""
    lib = _lib.load()
    if devices is None:
        _check(lib.pqh_knn_init(None, 0))
    else:
        arr = (C.c_int * len(devices))(*devices)
        _check(lib.pqh_knn_init(arr, len(devices)))


def shutdown():
    _lib.load().pqh_knn_shutdown()


def set_algo(algo):
    _check(_lib.load().pqh_knn_set_algo(int(algo)))


def last_stats():
    st = _lib.Stats()
    _check(_lib.load().pqh_knn_last_stats(C.byref(st)))
    return {name: getattr(st, name) for name, _ in _lib.Stats._fields_}


def last_fallback_rows():
    """Query rows of the last host-pointer search that the first tensor-core pass could not
    certify (redone by the second pass or the exact kernel): a test hook."""
    lib = _lib.load()
    n = lib.pqh_knn_last_fallback_rows(None, 0)
    out = np.empty(max(n, 1), dtype=np.int32)
    n = min(n, lib.pqh_knn_last_fallback_rows(out.ctypes.data, out.size))
    return out[:n]


def _as_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array [n x d]")
    return a


def knn_full_thread(base, queries=None, k=1, n_thread=1):
    """Exact squared-L2 kNN through the drop-in symbol.

    ``queries is None`` (or the same array object as ``base``) is the self-join
    compute_nn_fast performs: row i's first neighbour is i itself.
    Returns (assign int32[nq,k], dis float32[nq,k]).
    """
    lib = _lib.load()
    b = _as_f32(base)
    v = b if queries is None or queries is base else _as_f32(queries)
    if v.shape[1] != b.shape[1]:
        raise ValueError("dimension mismatch")
    nq, d = v.shape
    nb = b.shape[0]
    assign = np.empty((nq, k), dtype=np.int32)
    dis = np.empty((nq, k), dtype=np.float32)
    lib.knn_full_thread(2, nq, nb, d, k, b.ctypes.data, v.ctypes.data, None,
                        assign.ctypes.data, dis.ctypes.data, n_thread)
    return assign, dis


def search_rows(base, queries, k, q_begin, q_end, self_join=None, out=None):
    """Rows [q_begin,q_end) of the search (pqh_knn_search_rows): one rank's share."""
    lib = _lib.load()
    b = _as_f32(base)
    v = b if queries is None or queries is base else _as_f32(queries)
    if self_join is None:
        self_join = v is b
    nq, d = v.shape
    n = q_end - q_begin
    if out is None:
        assign = np.empty((n, k), dtype=np.int32)
        dis = np.empty((n, k), dtype=np.float32)
    else:
        assign, dis = out
    _check(lib.pqh_knn_search_rows(nq, b.shape[0], d, k, b.ctypes.data, v.ctypes.data,
                                   int(bool(self_join)), q_begin, q_end,
                                   assign.ctypes.data, dis.ctypes.data))
    return assign, dis


def search_device(device, nq_total, nb, d, k, d_base, d_queries, self_join, q_begin, q_end,
                  d_assign, d_dis, d_work, work_bytes, stream=None):
    """Device-pointer search (pqh_knn_search_device); pointers are integers."""
    lib = _lib.load()
    _check(lib.pqh_knn_search_device(device, nq_total, nb, d, k, d_base, d_queries,
                                     int(bool(self_join)), q_begin, q_end, d_assign, d_dis,
                                     d_work, work_bytes, stream))


def workspace_bytes(nq, nb, d, k):
    return int(_lib.load().pqh_knn_workspace_bytes(nq, nb, d, k))


def dataset_load(rows, ld=None):
    """Keep ``rows`` resident on every device of the context (pqh_knn_dataset_load).

    ``rows`` is float32[n, d]; or, with ``ld``, a flat float32 buffer whose vector i starts at
    element i*ld (an .fvecs image: pass the buffer from its second float on, ld = d + 1), in
    which case ``rows`` must be a tuple (buffer, n, d).
    """
    lib = _lib.load()
    if ld is None:
        a = _as_f32(rows)
        _check(lib.pqh_knn_dataset_load(a.shape[0], a.shape[1], a.ctypes.data, a.shape[1]))
    else:
        buf, n, d = rows
        _check(lib.pqh_knn_dataset_load(n, d, buf.ctypes.data, ld))


def dataset_search_block(row_ids, k, slot=0):
    """Self-join among the rows ``row_ids`` of the resident data set: what the reference's
    run_block gets from knn_full_thread for that block (ids are positions within the block)."""
    lib = _lib.load()
    ids = np.ascontiguousarray(row_ids, dtype=np.int32)
    m = ids.shape[0]
    assign = np.empty((m, k), dtype=np.int32)
    dis = np.empty((m, k), dtype=np.float32)
    _check(lib.pqh_knn_dataset_search_block(slot, m, ids.ctypes.data, k, assign.ctypes.data,
                                            dis.ctypes.data))
    return assign, dis


def dataset_release():
    _lib.load().pqh_knn_dataset_release()


# ---- device-side merge of the partitioned mode (pqh_knn_heaps_*) ----
ITEM = np.dtype([("id", "<u4"), ("dist", "<f4")])   # nn_item_t, fast_nn_temp_file.h:11-14


def heaps_create(n, k, image=None):
    """n heaps of k items on the device; ``image`` = an existing temp-file image (ITEM[n*k])."""
    lib = _lib.load()
    if image is not None:
        image = np.ascontiguousarray(image, dtype=ITEM)
        assert image.size == n * k
        _check(lib.pqh_knn_heaps_create(n, k, image.ctypes.data))
    else:
        _check(lib.pqh_knn_heaps_create(n, k, None))


def dataset_search_block_merge(row_ids, k1, ticket, slot=0):
    """Search the block ``row_ids`` of the resident data set and merge it into the heaps on the
    device (run_block + run_merge_block, compute_nn_fast.c:469-508); merges happen in ticket order."""
    lib = _lib.load()
    ids = np.ascontiguousarray(row_ids, dtype=np.int32)
    _check(lib.pqh_knn_dataset_search_block_merge(slot, ids.shape[0], ids.ctypes.data, k1, ticket))


def heaps_export(n, k, want_image=True):
    """(image ITEM[n,k] or None, ids uint32[n,k], dists float32[n,k]) as the reference's temp file
    and result files hold them."""
    lib = _lib.load()
    image = np.empty((n, k), dtype=ITEM) if want_image else None
    ids = np.empty((n, k), dtype=np.uint32)
    dists = np.empty((n, k), dtype=np.float32)
    _check(lib.pqh_knn_heaps_export(image.ctypes.data if want_image else None, ids.ctypes.data,
                                    dists.ctypes.data))
    return image, ids, dists


def heaps_release():
    _lib.load().pqh_knn_heaps_release()


EDGE = np.dtype([("source", "<i4"), ("target", "<i4"), ("dist", "<f4")])   # mst_edge_t, mst.h:12-16
SORT_BY_DIST, SORT_BY_SOURCE = 0, 1
ENOTORDER = -6


def sort_edges(edges, mode=SORT_BY_DIST):
    """In-place stable sort of EDGE records on the device, the order mst.c:175 / :214 get from
    qsort.  Returns False (records untouched) when the reference's distance comparator is not an
    order on this data (PQH_ENOTORDER)."""
    lib = _lib.load()
    assert edges.dtype == EDGE and edges.flags.c_contiguous
    rc = lib.pqh_sort_edges(edges.shape[0], edges.ctypes.data, mode)
    if rc == ENOTORDER:
        return False
    _check(rc)
    return True


def nn_joint_histograms(pq_codes, nn_indices, k_star=256, nn_first=0, nn_count=None):
    """(stats float64[nn_count, m, k_star, k_star], num_known int64[nn_count]): what
    collect_stats_at_kth_nn (nn_huffman_estimator.c:69-107) computes for each neighbour rank."""
    lib = _lib.load()
    pq = np.ascontiguousarray(pq_codes, dtype=np.uint8)
    nn = np.ascontiguousarray(nn_indices).view(np.uint32)
    n, m = pq.shape
    num_nn = nn.shape[1]
    assert nn.shape[0] == n
    if nn_count is None:
        nn_count = num_nn - nn_first
    stats = np.empty((nn_count, m, k_star, k_star), dtype=np.float64)
    known = np.empty(nn_count, dtype=np.int64)
    _check(lib.pqh_nn_joint_histograms(n, m, k_star, num_nn, pq.ctypes.data, nn.ctypes.data,
                                       nn_first, nn_count, stats.ctypes.data, known.ctypes.data))
    return stats, known
