"""Host logic of the native driver's partitioned mode (pq_huffman_b200/host/pqh_partition.c,
built as libpqh_host.so) against the reference's OWN compiled objects (oracle/_ref/libref_cnf.so):
the blocks-info cache file must come out byte-identical (it holds every box boundary), a cache
written by the reference must load, and block membership must agree row by row with
is_vector_in_block (fast_nn_blocks_info.c:174-227)."""
import ctypes as C
import os

import numpy as np
import pytest

from pq_huffman_b200 import synth, vecs_io
from tests import oracle_ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "pq_huffman_b200", "libpqh_host.so")


class Split(C.Structure):
    _fields_ = [("dimension", C.c_int), ("digit_weight", C.c_longlong),
                ("lo", C.POINTER(C.c_float)), ("hi", C.POINTER(C.c_float))]


class Partition(C.Structure):
    _fields_ = [("num_splits", C.c_int), ("overlap", C.c_double), ("blocks_per_dim", C.c_int),
                ("num_blocks", C.c_longlong), ("splits", C.POINTER(Split))]


@pytest.fixture(scope="module")
def hostlib():
    if not os.path.exists(HOSTLIB):
        pytest.skip("libpqh_host.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(HOSTLIB)
    vp = C.c_void_p
    lib.pqh_partition_build.argtypes = [C.POINTER(Partition), vp, C.c_longlong, C.c_longlong,
                                        C.c_int, C.c_int, C.c_double, C.c_int]
    lib.pqh_partition_save.argtypes = [C.POINTER(Partition), C.c_char_p]
    lib.pqh_partition_load.argtypes = [C.POINTER(Partition), C.c_char_p]
    lib.pqh_partition_columns.argtypes = [C.POINTER(Partition), vp, C.c_longlong, C.c_longlong]
    lib.pqh_partition_columns.restype = C.POINTER(C.c_float)
    lib.pqh_partition_block_rows.argtypes = [C.POINTER(Partition), C.POINTER(C.c_float),
                                             C.c_longlong, C.c_longlong, C.c_longlong,
                                             C.c_longlong, C.c_int, vp]
    lib.pqh_partition_block_rows.restype = C.c_longlong
    lib.pqh_partition_free.argtypes = [C.POINTER(Partition)]
    return lib


def _fvecs_image(x):
    n, d = x.shape
    img = np.empty((n, d + 1), dtype=np.float32)
    img[:, 0] = np.array([d], dtype=np.int32).view(np.float32)[0]
    img[:, 1:] = x
    return img


def _ref_blocks_info(refcnf, path, n, d, splits, per_dim, overlap):
    info = (C.c_char * 64)()                      # blocks_info_t, opaque here
    refcnf.blocks_info_init.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong, C.c_int, C.c_int,
                                        C.c_int, C.c_double]
    refcnf.blocks_info_init.restype = None
    refcnf.blocks_info_init(info, path.encode(), n, d, splits, per_dim, overlap)
    return info


CASES = [("sift", 3000, 24, 5, 3, 0.3), ("gist", 2500, 16, 3, 4, 0.1), ("sift", 500, 8, 2, 2, 0.0),
         ("ties", 1200, 12, 4, 3, 0.5), ("gist", 64, 6, 1, 7, 0.3)]


@pytest.mark.parametrize("gen,n,d,splits,per_dim,overlap", CASES)
def test_partition_matches_reference(tmp_path, hostlib, refcnf, gen, n, d, splits, per_dim, overlap):
    x = {"sift": synth.sift_like, "gist": synth.gist_like}.get(gen, synth.sift_like)(n, d, seed=n)
    if gen == "ties":
        x = synth.with_duplicates_and_ties(x)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    img = _fvecs_image(x)

    info = _ref_blocks_info(refcnf, inp, n, d, splits, per_dim, overlap)
    ref_cache = str(tmp_path / "ref.cache")
    refcnf.blocks_info_save_filename.argtypes = [C.c_void_p, C.c_char_p]
    refcnf.blocks_info_save_filename(info, ref_cache.encode())

    part = Partition()
    base = img.ctypes.data + 4                    # first float of row 0, rows d+1 floats apart
    assert hostlib.pqh_partition_build(C.byref(part), base, n, d + 1, splits, per_dim, overlap, 4) == 0
    assert part.num_blocks == per_dim ** splits
    my_cache = str(tmp_path / "my.cache")
    assert hostlib.pqh_partition_save(C.byref(part), my_cache.encode()) == 0
    assert open(my_cache, "rb").read() == open(ref_cache, "rb").read()

    loaded = Partition()
    assert hostlib.pqh_partition_load(C.byref(loaded), ref_cache.encode()) == 0
    assert loaded.num_splits == splits and loaded.num_blocks == part.num_blocks

    cols = hostlib.pqh_partition_columns(C.byref(loaded), base, n, d + 1)
    refcnf.is_vector_in_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
    refcnf.is_vector_in_block.restype = C.c_int
    rng = np.random.default_rng(1)
    blocks = sorted(set([0, int(part.num_blocks) - 1] +
                        [int(b) for b in rng.integers(0, part.num_blocks, size=12)]))
    out = np.empty(n, dtype=np.int32)
    covered = 0
    for b in blocks:
        m = hostlib.pqh_partition_block_rows(C.byref(loaded), cols, 0, n, b, b, 0, out.ctypes.data)
        ref_rows = [i for i in range(n)
                    if refcnf.is_vector_in_block(x[i].ctypes.data, d, info, b)]
        assert out[:m].tolist() == ref_rows, f"block {b}"
        covered += m
        # a row range gives the matching slice
        lo, hi = n // 3, 2 * n // 3
        m2 = hostlib.pqh_partition_block_rows(C.byref(loaded), cols, lo, hi, b, b, 0, out.ctypes.data)
        assert out[:m2].tolist() == [i for i in ref_rows if lo <= i < hi]
    assert covered > 0
    hostlib.pqh_partition_free(C.byref(part))
    hostlib.pqh_partition_free(C.byref(loaded))


def test_every_row_is_in_some_block(hostlib):
    n, d = 4000, 10
    x = synth.gist_like(n, d, seed=4)
    img = _fvecs_image(x)
    part = Partition()
    assert hostlib.pqh_partition_build(C.byref(part), img.ctypes.data + 4, n, d + 1, 3, 3, 0.3, 2) == 0
    cols = hostlib.pqh_partition_columns(C.byref(part), img.ctypes.data + 4, n, d + 1)
    seen = np.zeros(n, dtype=np.int64)
    out = np.empty(n, dtype=np.int32)
    for b in range(part.num_blocks):
        m = hostlib.pqh_partition_block_rows(C.byref(part), cols, 0, n, b, b, 0, out.ctypes.data)
        assert (np.diff(out[:m]) > 0).all()
        seen[out[:m]] += 1
    assert seen.min() >= 1 and seen.max() > 1          # a cover, and the boxes do overlap
    hostlib.pqh_partition_free(C.byref(part))
