"""The native host driver pq_huffman_b200/pqh_compute_nn_fast (bulk I/O, single-block exact mode)
must write the same bytes as the reference's own CLI: compared with oracle/_ref/
compute_nn_fast_oracle (reference objects + CPU oracle) on inputs with ties and duplicates,
including the kept temp file, and with the golden digests."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from pq_huffman_b200 import synth, vecs_io
from tests import oracle_ffi
from tests.golden import make_golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pq_huffman_b200", "pqh_compute_nn_fast")
REF_CLI = os.path.join(oracle_ffi.REF_DIR, "compute_nn_fast_oracle")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def _run(binary, inp, outdir, k, extra=()):
    os.makedirs(outdir, exist_ok=True)
    tpl = str(outdir) + "/"
    p = subprocess.run([binary, inp, tpl, str(k), "--num-dims", "0", "--temp-file",
                        tpl + "tmp.vecs", "--num-threads", "4", *extra],
                       capture_output=True, text=True, cwd=str(outdir))
    assert p.returncode == 0, p.stderr[-800:]
    return tpl


@pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(REF_CLI)),
                    reason="host CLI or oracle/_ref not built")
@pytest.mark.parametrize("gen,n,d,k", [("ties", 6000, 32, 16), ("gist", 3000, 960, 16),
                                       ("ties", 40, 8, 64), ("sift", 2, 16, 5)])
def test_host_cli_matches_reference_cli(tmp_path, gen, n, d, k):
    x = make_golden.make_input("ties" if gen == "ties" else gen + "_like", n, d, 900 + n)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    ours = _run(CLI, inp, tmp_path / "ours", k, ["--no-delete-temp-file"])
    ref = _run(REF_CLI, inp, tmp_path / "ref", k, ["--no-delete-temp-file"])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(ours + name) == _sha(ref + name), f"{name} differs"


@pytest.mark.skipif(not os.path.exists(CLI), reason="host CLI not built")
def test_host_cli_reproduces_golden_digest(tmp_path):
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "index.json")))["cli"][
        "cli_single_block"]
    x = make_golden.make_input("ties", c["n"], c["d"], c["seed"])
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    tpl = _run(CLI, inp, tmp_path / "o", c["num_nn"])
    assert _sha(tpl + "nn_indices.ivecsl") == c["nn_indices.ivecsl"]
    assert _sha(tpl + "nn_dist.fvecsl") == c["nn_dist.fvecsl"]
    assert not os.path.exists(tpl + "tmp.vecs")          # deleted by default, as the reference


@pytest.mark.skipif(not os.path.exists(CLI), reason="host CLI not built")
def test_host_cli_rejects_partitioned_mode(tmp_path):
    x = synth.sift_like(100, 8)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    p = subprocess.run([CLI, inp, str(tmp_path) + "/", "4", "--num-dims", "2"],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "single-block" in p.stderr
