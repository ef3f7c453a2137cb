"""The reference's on-disk vector formats (host side, numpy only).

* ``.fvecs`` / ``.bvecs`` / ``.ivecs``: per row an int32 dimension count followed
  by d values (compute_nn_fast.c:134-149 derives N from the file size;
  fast_nn_block_loader2.c:99-108 reads rows of d+1 floats).
* "light" files (``.ivecsl`` / ``.fvecsl`` / ``.bvecsl``): one 8-byte header
  ``uint32 n, uint32 d`` then n*d dense values (vecs_io.c:43-56,70-76,
  utils/io.py:74-98).  compute_nn_fast's outputs are light files
  (fast_nn_temp_file.c:253-301).
"""
import struct

import numpy as np

_VEC_DTYPES = {"f": np.float32, "i": np.int32, "b": np.uint8}


def write_xvecs(path, matrix, kind="f"):
    """Write [n x d] as .fvecs/.ivecs/.bvecs (int32 d prefix per row)."""
    dt = _VEC_DTYPES[kind]
    m = np.ascontiguousarray(matrix, dtype=dt)
    n, d = m.shape
    rec = np.dtype([("d", "<i4"), ("v", dt, (d,))])
    out = np.empty(n, dtype=rec)
    out["d"] = d
    out["v"] = m
    out.tofile(path)


def read_xvecs(path, kind="f"):
    dt = _VEC_DTYPES[kind]
    with open(path, "rb") as f:
        head = f.read(4)
        if not head:
            return np.empty((0, 0), dtype=dt)
        d = struct.unpack("<i", head)[0]
    rec = np.dtype([("d", "<i4"), ("v", dt, (d,))])
    raw = np.fromfile(path, dtype=rec)
    if not np.all(raw["d"] == d):
        raise ValueError(f"{path}: rows with differing dimension prefixes")
    return np.ascontiguousarray(raw["v"])


def write_light(path, matrix):
    """Write a light file: uint32 n, uint32 d, then the dense matrix."""
    m = np.ascontiguousarray(matrix)
    if m.ndim == 1:
        m = m.reshape(-1, 1)
    with open(path, "wb") as f:
        f.write(struct.pack("<II", m.shape[0], m.shape[1]))
        m.tofile(f)


def read_light(path, dtype):
    with open(path, "rb") as f:
        n, d = struct.unpack("<II", f.read(8))
        data = np.fromfile(f, dtype=dtype, count=n * d)
    if data.size != n * d:
        raise ValueError(f"{path}: truncated light file ({data.size} of {n * d} values)")
    return data.reshape(n, d)


def read_nn_result(template):
    """(<tpl>nn_indices.ivecsl as uint32, <tpl>nn_dist.fvecsl) -- compute_nn_fast.c:224-225."""
    return (read_light(template + "nn_indices.ivecsl", np.uint32),
            read_light(template + "nn_dist.fvecsl", np.float32))
