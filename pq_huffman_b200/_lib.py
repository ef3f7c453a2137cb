"""ctypes loader for libpqh_knn.so (C ABI: include/pqh_knn.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpqh_knn.so")

# every symbol include/pqh_knn.h declares; tests/test_abi.py checks the export list
EXPORTS = (
    "knn_full_thread",
    "pqh_knn_init",
    "pqh_knn_shutdown",
    "pqh_knn_device_count",
    "pqh_last_error",
    "pqh_knn_set_algo",
    "pqh_knn_search_rows",
    "pqh_knn_search_rows_ld",
    "pqh_knn_workspace_bytes",
    "pqh_knn_search_device",
    "pqh_knn_last_stats",
    "pqh_knn_last_fallback_rows",
    "pqh_knn_dataset_load",
    "pqh_knn_dataset_search_block",
    "pqh_knn_dataset_release",
    "pqh_knn_heaps_create",
    "pqh_knn_dataset_search_block_merge",
    "pqh_knn_heaps_export",
    "pqh_knn_heaps_release",
    "pqh_knn_host_alloc",
    "pqh_knn_host_free",
    "pqh_sort_edges",
    "pqh_qsort_edges",
    "pqh_nn_joint_histograms",
)


class Stats(C.Structure):
    """struct pqh_knn_stats (include/pqh_knn.h)."""

    _fields_ = [
        ("rows", C.c_longlong),
        ("rows_tensor", C.c_longlong),
        ("rows_exact", C.c_longlong),
        ("rows_second_pass", C.c_longlong),
        ("launches", C.c_longlong),
        ("ms_total", C.c_double),
        ("ms_h2d", C.c_double),
        ("ms_kernels", C.c_double),
        ("ms_d2h", C.c_double),
        ("ms_prep", C.c_double),
        ("ms_gemm", C.c_double),
        ("ms_rerank", C.c_double),
        ("ms_exact", C.c_double),
        ("candidates", C.c_double),
        ("eps_ratio_max", C.c_double),
        ("inexact_operands", C.c_longlong),
    ]


_lib = None


def load():
    """Load the CUDA library; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C pq_huffman_b200/csrc`). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    lib.knn_full_thread.argtypes = [C.c_int] * 5 + [vp, vp, vp, vp, vp, C.c_int]
    lib.knn_full_thread.restype = None
    lib.pqh_knn_init.argtypes = [ip, C.c_int]
    lib.pqh_knn_init.restype = C.c_int
    lib.pqh_knn_shutdown.argtypes = []
    lib.pqh_knn_shutdown.restype = None
    lib.pqh_knn_device_count.argtypes = []
    lib.pqh_knn_device_count.restype = C.c_int
    lib.pqh_last_error.argtypes = []
    lib.pqh_last_error.restype = C.c_char_p
    lib.pqh_knn_set_algo.argtypes = [C.c_int]
    lib.pqh_knn_set_algo.restype = C.c_int
    lib.pqh_knn_search_rows.argtypes = [C.c_int] * 4 + [vp, vp] + [C.c_int] * 3 + [vp, vp]
    lib.pqh_knn_search_rows.restype = C.c_int
    lib.pqh_knn_search_rows_ld.argtypes = ([C.c_int] * 4 + [vp, C.c_longlong, vp, C.c_longlong]
                                           + [C.c_int] * 3 + [vp, vp])
    lib.pqh_knn_search_rows_ld.restype = C.c_int
    lib.pqh_knn_workspace_bytes.argtypes = [C.c_int] * 4
    lib.pqh_knn_workspace_bytes.restype = C.c_longlong
    lib.pqh_knn_search_device.argtypes = (
        [C.c_int] * 5 + [vp, vp] + [C.c_int] * 3 + [vp, vp, vp, C.c_longlong, vp]
    )
    lib.pqh_knn_search_device.restype = C.c_int
    lib.pqh_knn_last_stats.argtypes = [C.POINTER(Stats)]
    lib.pqh_knn_last_stats.restype = C.c_int
    lib.pqh_knn_last_fallback_rows.argtypes = [vp, C.c_int]
    lib.pqh_knn_last_fallback_rows.restype = C.c_int
    lib.pqh_knn_dataset_load.argtypes = [C.c_int, C.c_int, vp, C.c_longlong]
    lib.pqh_knn_dataset_load.restype = C.c_int
    lib.pqh_knn_dataset_search_block.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, vp]
    lib.pqh_knn_dataset_search_block.restype = C.c_int
    lib.pqh_knn_heaps_create.argtypes = [C.c_longlong, C.c_int, vp]
    lib.pqh_knn_heaps_create.restype = C.c_int
    lib.pqh_knn_dataset_search_block_merge.argtypes = [C.c_int, C.c_int, vp, C.c_int, C.c_longlong]
    lib.pqh_knn_dataset_search_block_merge.restype = C.c_int
    lib.pqh_knn_heaps_export.argtypes = [vp, vp, vp]
    lib.pqh_knn_heaps_export.restype = C.c_int
    lib.pqh_knn_heaps_release.argtypes = []
    lib.pqh_knn_heaps_release.restype = None
    lib.pqh_knn_dataset_release.argtypes = []
    lib.pqh_knn_dataset_release.restype = None
    lib.pqh_knn_host_alloc.argtypes = [C.c_longlong]
    lib.pqh_knn_host_alloc.restype = C.c_void_p
    lib.pqh_knn_host_free.argtypes = [C.c_void_p]
    lib.pqh_knn_host_free.restype = None
    lib.pqh_sort_edges.argtypes = [C.c_longlong, vp, C.c_int]
    lib.pqh_sort_edges.restype = C.c_int
    lib.pqh_qsort_edges.argtypes = [vp, C.c_size_t, C.c_size_t, vp]
    lib.pqh_qsort_edges.restype = None
    lib.pqh_nn_joint_histograms.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int,
                                            C.c_int, vp, vp]
    lib.pqh_nn_joint_histograms.restype = C.c_int
    del fp
    _lib = lib
    return lib
