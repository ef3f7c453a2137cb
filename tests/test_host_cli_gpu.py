"""The native host driver pq_huffman_b200/pqh_compute_nn_fast (bulk I/O; exact single-block mode
and the partitioned approximate mode over a data set resident in HBM) must write the same bytes
as the reference's own CLI: compared with oracle/_ref/compute_nn_fast_oracle (reference objects +
CPU oracle) on inputs with ties and duplicates, including the kept temp file, the blocks-info
cache and the block statistics, and with the golden digests."""
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


def _golden_case(binary, tmp_path, name):
    """Run `binary` on the committed golden input of tests/golden/index.json["cli"][name] and
    compare the digests of both output files (made by the reference CLI + oracle,
    tests/golden/make_golden.py)."""
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "index.json")))["cli"][name]
    x = make_golden.make_input("ties", c["n"], c["d"], c["seed"])
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    assert hashlib.sha256(x.tobytes()).hexdigest() == c["input_sha256"]
    os.makedirs(tmp_path / "o", exist_ok=True)
    tpl = str(tmp_path / "o") + "/"
    p = subprocess.run([binary, inp, tpl, str(c["num_nn"]), *c["args"], "--temp-file",
                        tpl + "tmp.vecs"], capture_output=True, text=True, cwd=tpl)
    assert p.returncode == 0, p.stderr[-800:]
    assert _sha(tpl + "nn_indices.ivecsl") == c["nn_indices.ivecsl"]
    assert _sha(tpl + "nn_dist.fvecsl") == c["nn_dist.fvecsl"]
    assert not os.path.exists(tpl + "tmp.vecs")          # deleted by default, as the reference


@pytest.mark.skipif(not os.path.exists(CLI), reason="host CLI not built")
@pytest.mark.parametrize("name", ["cli_single_block", "cli_partitioned"])
def test_host_cli_reproduces_golden_digest(tmp_path, name):
    _golden_case(CLI, tmp_path, name)


def _run_partitioned(binary, inp, outdir, k, extra=()):
    os.makedirs(outdir, exist_ok=True)
    tpl = str(outdir) + "/"
    p = subprocess.run([binary, inp, tpl, str(k), "--temp-file", tpl + "tmp.vecs",
                        "--num-threads", "4", *extra],
                       capture_output=True, text=True, cwd=str(outdir))
    assert p.returncode == 0, p.stderr[-800:]
    return tpl, p.stdout


PARTITIONED = [
    # generator, n, d, k, flags: the reference's default geometry (5 splits x 3 boxes = 243 blocks)
    ("ties", 5000, 16, 8, []),
    # blocks large enough for the tensor-core path (9 blocks of ~10-25k rows)
    ("sift_like", 60000, 32, 16, ["--num-dims", "2"]),
    ("gist_like", 30000, 24, 5, ["--num-dims", "3", "--num-blocks-per-dim", "2",
                                 "--block-overlap-fraction", "0.5"]),
    # more neighbours asked than the small blocks hold: heaps keep {UINT_MAX, inf} slots
    ("sift_like", 300, 8, 24, ["--num-dims", "4"]),
]


@pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(REF_CLI)),
                    reason="host CLI or oracle/_ref not built")
@pytest.mark.parametrize("gen,n,d,k,flags", PARTITIONED)
def test_host_cli_partitioned_mode_matches_reference_cli(tmp_path, gen, n, d, k, flags):
    x = make_golden.make_input(gen, n, d, 70 + n)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    extra = [*flags, "--no-delete-temp-file", "--with-blocks-stat"]
    ours, out = _run_partitioned(CLI, inp, tmp_path / "ours", k,
                                 [*extra, "--blocks-info-cache", str(tmp_path / "ours.cache")])
    ref, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref", k,
                              [*extra, "--blocks-info-cache", str(tmp_path / "ref.cache")])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs", "blocks_stat.txt"):
        assert _sha(ours + name) == _sha(ref + name), f"{name} differs"
    assert _sha(str(tmp_path / "ours.cache")) == _sha(str(tmp_path / "ref.cache"))
    assert "blocks" in out
    # second run: boundaries from the REFERENCE's cache file, same result
    again, out2 = _run_partitioned(CLI, inp, tmp_path / "again", k,
                                   [*flags, "--blocks-info-cache", str(tmp_path / "ref.cache")])
    assert "Blocks info: Loaded" in out2
    assert _sha(again + "nn_indices.ivecsl") == _sha(ref + "nn_indices.ivecsl")
    assert _sha(again + "nn_dist.fvecsl") == _sha(ref + "nn_dist.fvecsl")
    assert not os.path.exists(again + "tmp.vecs")


@pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(REF_CLI)),
                    reason="host CLI or oracle/_ref not built")
def test_host_cli_resumes_from_temp_file(tmp_path):
    """--blocks-from and --no-init-temp-file (compute_nn_fast.c:545-555, 626-629): a run over
    blocks [3,9) into fresh heaps, then a second run over blocks [6,9) that continues from the
    REFERENCE's kept temp file -- both legs byte-identical to the reference doing the same."""
    import shutil
    n, d, k = 20000, 16, 10
    x = make_golden.make_input("sift_like", n, d, 5)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    flags = ["--num-dims", "2", "--no-delete-temp-file"]
    ours, _ = _run_partitioned(CLI, inp, tmp_path / "ours", k, [*flags, "--blocks-from", "3"])
    ref, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref", k, [*flags, "--blocks-from", "3"])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(ours + name) == _sha(ref + name), f"leg 1: {name} differs"
    whole, _ = _run_partitioned(REF_CLI, inp, tmp_path / "whole", k, flags)
    assert _sha(whole + "nn_indices.ivecsl") != _sha(ref + "nn_indices.ivecsl")  # a partial result

    for who in ("ours2", "ref2"):
        os.makedirs(tmp_path / who)
        shutil.copy(ref + "tmp.vecs", str(tmp_path / who / "tmp.vecs"))
    ours2, _ = _run_partitioned(CLI, inp, tmp_path / "ours2", k,
                                [*flags, "--blocks-from", "6", "--no-init-temp-file"])
    ref2, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref2", k,
                               [*flags, "--blocks-from", "6", "--no-init-temp-file"])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(ours2 + name) == _sha(ref2 + name), f"leg 2: {name} differs"


@pytest.mark.skipif(not os.path.exists(CLI), reason="host CLI not built")
def test_host_merge_and_device_merge_write_the_same_bytes(tmp_path):
    """The partitioned mode merges the blocks on the device by default (pqh_knn_heaps_*);
    --host-merge keeps the heaps on the host, merged by the driver's own restatement of
    fast_nn_heap_push.  Results, and the kept temp file, must not differ."""
    x = make_golden.make_input("ties", 30000, 24, 3)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    flags = ["--num-dims", "3", "--no-delete-temp-file"]
    dev, out = _run_partitioned(CLI, inp, tmp_path / "dev", 14, flags)
    host, _ = _run_partitioned(CLI, inp, tmp_path / "host", 14, [*flags, "--host-merge"])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(dev + name) == _sha(host + name), f"{name} differs"
    assert "d2h 0.0 ms" in out or "d2h" in out


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(not (os.path.exists(CLI) and os.path.exists(REF_CLI)) or _n_devices() < 2,
                    reason="needs the host CLI, oracle/_ref and two GPUs")
def test_host_cli_partitioned_mode_on_two_gpus(tmp_path):
    """PQH_DEVICES=0,1: blocks alternate between the devices, merges stay in block order, so the
    bytes do not change."""
    x = make_golden.make_input("sift_like", 50000, 32, 11)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    flags = ["--num-dims", "3", "--no-delete-temp-file"]
    os.makedirs(tmp_path / "two")
    tpl = str(tmp_path / "two") + "/"
    env = dict(os.environ, PQH_DEVICES="0,1")
    env.pop("CUDA_VISIBLE_DEVICES", None)
    p = subprocess.run([CLI, inp, tpl, "12", "--temp-file", tpl + "tmp.vecs", *flags],
                       capture_output=True, text=True, cwd=tpl, env=env)
    assert p.returncode == 0, p.stderr[-800:]
    ref, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref", 12, flags)
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(tpl + name) == _sha(ref + name), f"{name} differs"
