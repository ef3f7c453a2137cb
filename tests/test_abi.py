"""CPU-side checks of the drop-in boundary: the library loads and exports what
include/pqh_knn.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from pq_huffman_b200 import _lib, knn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pqh_knn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(knn_full_thread|pqh_[a-z_]+)\s*\(", text)))


def test_header_and_loader_agree():
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_functions():
        assert getattr(lib, name) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared_functions()) <= exported
    # nothing but the C ABI is exported as a C symbol from our namespace
    assert not [s for s in exported if s.startswith("pqh_") and s not in _lib.EXPORTS]


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_without_fallback():
    x = np.zeros((8, 4), dtype=np.float32)
    with pytest.raises(knn.PqhError) as ei:
        knn.search_rows(x, None, 2, 0, 8)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    # the void drop-in symbol follows the reference's convention: message + exit(1)
    code = ("import numpy as np; from pq_huffman_b200 import knn; "
            "x=np.zeros((8,4),dtype=np.float32); knn.knn_full_thread(x,None,2)")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert p.returncode == 1
    assert "libpqh_knn" in p.stderr


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    x = np.zeros((8, 4), dtype=np.float32)
    out_i = np.zeros((8, 2), dtype=np.int32)
    out_d = np.zeros((8, 2), dtype=np.float32)
    rc = lib.pqh_knn_search_rows(8, 8, 4, 2, x.ctypes.data, x.ctypes.data, 1, 5, 3,
                                 out_i.ctypes.data, out_d.ctypes.data)
    assert rc == -1 and b"row range" in lib.pqh_last_error()
    rc = lib.pqh_knn_search_rows(8, 8, 0, 2, x.ctypes.data, x.ctypes.data, 1, 0, 8,
                                 out_i.ctypes.data, out_d.ctypes.data)
    assert rc == -1
    assert lib.pqh_knn_set_algo(77) == -1
    assert lib.pqh_knn_workspace_bytes(1000, 1000, 16, 9) > 0


def test_qsort_shaped_entry_is_qsort_for_what_it_does_not_take():
    """pqh_qsort_edges (mst.c built with -Dqsort=pqh_qsort_edges) only takes over large arrays of
    12-byte edge records; anything else -- here 4-byte ints, and a short array of edge records --
    is sorted by libc, GPU or not."""
    lib = _lib.load()
    CMP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)

    def int_cmp(a, b):
        x, y = C.cast(a, C.POINTER(C.c_int))[0], C.cast(b, C.POINTER(C.c_int))[0]
        return (x > y) - (x < y)

    v = np.random.default_rng(0).integers(-1000, 1000, 500).astype(np.int32)
    w = v.copy()
    lib.pqh_qsort_edges(w.ctypes.data, w.size, 4, C.cast(CMP(int_cmp), C.c_void_p))
    assert np.array_equal(w, np.sort(v))

    e = np.zeros(300, dtype=knn.EDGE)
    e["dist"] = np.random.default_rng(1).integers(0, 50, 300).astype(np.float32)
    e["source"] = np.arange(300)

    def dist_cmp(a, b):
        x = C.cast(a + 8, C.POINTER(C.c_float))[0]
        y = C.cast(b + 8, C.POINTER(C.c_float))[0]
        return (x > y) - (x < y)

    s = e.copy()
    lib.pqh_qsort_edges(s.ctypes.data, s.size, 12, C.cast(CMP(dist_cmp), C.c_void_p))
    assert np.all(np.diff(s["dist"]) >= 0) and sorted(s["source"]) == list(range(300))


def test_workspace_is_bounded_in_the_number_of_rows():
    """pqh_knn_workspace_bytes needs no device: candidate lists are kept for one slice of query
    rows at a time ($PQH_TC_SLICE_MB, 2 GiB by default), so a Deep10M-sized search asks for
    operands + one slice, not for 41 GB of lists (round 1)."""
    w1 = knn.workspace_bytes(1_000_000, 1_000_000, 128, 17)
    w10 = knn.workspace_bytes(10_000_000, 10_000_000, 96, 33)
    assert 0 < w1 <= 4 * 2**30
    assert 0 < w10 <= 8 * 2**30
    # the lists do not grow with n beyond the slice budget: 10x the rows, far less than 10x the bytes
    assert w10 < 4 * w1


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_edge_sort_without_a_gpu_fails_loudly():
    """A sort that pqh_qsort_edges takes (large array of edge records, comparator answering like
    mst_dist_comparator) is the device's job: without a device the process ends with a message
    and exit(1), it does not quietly sort on the CPU."""
    code = r'''
import ctypes as C, numpy as np
from pq_huffman_b200 import _lib, knn
lib = _lib.load()
CMP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
def dist_cmp(a, b):
    x = C.cast(a + 8, C.POINTER(C.c_float))[0]; y = C.cast(b + 8, C.POINTER(C.c_float))[0]
    d = np.float32(x) - np.float32(y)
    return 0 if abs(float(d)) < 1e-9 else (-1 if d < 0 else 1)
e = np.zeros(70000, dtype=knn.EDGE); e["dist"] = np.arange(70000, dtype=np.float32)[::-1]
lib.pqh_qsort_edges(e.ctypes.data, e.size, 12, C.cast(CMP(dist_cmp), C.c_void_p))
print("sorted on the CPU")
'''
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert p.returncode == 1 and "sorted on the CPU" not in p.stdout
    assert "pqh_qsort_edges failed" in p.stderr
