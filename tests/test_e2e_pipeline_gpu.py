"""BASELINE config 5 in miniature: GPU kNN -> mst_builder -> huffman_encoder, all through the
reference's OWN unmodified binaries (oracle/_ref, built from /root/reference/src by
oracle/Makefile and shipped prebuilt to the GPU box).

  compute_nn_fast_gpu     = reference objects + libpqh_knn.so  (the drop-in)
  compute_nn_fast_oracle  = reference objects + the CPU oracle

The two must write byte-identical nn_indices.ivecsl / nn_dist.fvecsl (single block and the
default partitioned mode), hence identical MST and identical Huffman bit counts downstream.
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from pq_huffman_b200 import synth, vecs_io
from tests import oracle_ffi

pytestmark = pytest.mark.gpu

REF = oracle_ffi.REF_DIR
NEEDED = ["compute_nn_fast_gpu", "compute_nn_fast_oracle", "mst_builder", "huffman_encoder"]


def _have_ref():
    return all(os.path.exists(os.path.join(REF, b)) for b in NEEDED)


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def _run(binary, args, cwd):
    p = subprocess.run([os.path.join(REF, binary)] + args, cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, f"{binary} failed: {p.stderr[-500:]}"
    return p.stdout


def _knn_cli(kind, tmp_path, inp, k, extra):
    out = tmp_path / kind
    out.mkdir()
    tpl = str(out) + "/"
    _run(f"compute_nn_fast_{kind}", [inp, tpl, str(k), "--temp-file", str(out / "tmp.vecs"),
                                     "--num-threads", "8"] + extra, str(out))
    return tpl


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("mode,extra", [
    ("single_block", ["--num-dims", "0"]),
    ("partitioned", ["--num-dims", "2", "--num-blocks-per-dim", "3",
                     "--block-overlap-fraction", "0.3"]),
])
def test_cli_outputs_and_downstream_bits_identical(tmp_path, mode, extra):
    n, d, k, m = 20000, 128, 16, 8
    x = synth.with_duplicates_and_ties(synth.sift_like(n, d, seed=31), n_dups=10, n_ties=10)
    inp = str(tmp_path / "data.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    pq_tpl = str(tmp_path) + "/pq_"
    vecs_io.write_light(pq_tpl + "pq_indices.bvecsl", synth.pq_codes_stand_in(x, m))

    tpl = {kind: _knn_cli(kind, tmp_path, inp, k, extra) for kind in ("oracle", "gpu")}
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl"):
        assert _sha(tpl["oracle"] + name) == _sha(tpl["gpu"] + name), f"{mode}: {name} differs"

    ids, dist = vecs_io.read_nn_result(tpl["gpu"])
    assert ids.shape == (n, k) and np.all(np.diff(dist, axis=1) >= 0)

    stats = {}
    for kind in ("oracle", "gpu"):
        _run("mst_builder", [tpl[kind], tpl[kind], str(k), "--pq-template", pq_tpl], str(tmp_path))
        _run("huffman_encoder", [pq_tpl, tpl[kind], str(m), "--no-sort", "--tree",
                                 tpl[kind] + "mst.tree"], str(tmp_path))
        stats[kind] = {f: _sha(tpl[kind] + f) for f in
                       ("mst.tree", "huffman_indices.bin", "huffman_codebooks.bin",
                        "huffman_children.bin")}
        stats[kind]["stats"] = open(tpl[kind] + "stats.json").read()
        stats[kind]["hstats"] = open(tpl[kind] + "huffman_stats.txt").read()
    assert stats["oracle"] == stats["gpu"], f"{mode}: downstream artefacts differ"
    assert "length_bit" in stats["gpu"]["hstats"] or "length_bit" in stats["gpu"]["stats"]
