"""Host logic of the native driver on the CPU: pq_huffman_b200/host/*.c linked to a CPU stand-in
for libpqh_knn.so (oracle/pqh_shim.c -> oracle/lib/pqh_compute_nn_fast_shim, test infrastructure)
must write the bytes the reference's own CLI writes -- partition, block order, heap merge,
temp-file resume, cache file, block statistics.  The same comparisons run with the real binary
over the GPU library in tests/test_host_cli_gpu.py."""
import os
import shutil

import pytest

from pq_huffman_b200 import vecs_io
from tests import oracle_ffi
from tests.golden import make_golden
from tests.test_host_cli_gpu import REF_CLI, _golden_case, _run, _run_partitioned, _sha

SHIM = os.path.join(oracle_ffi.LIB_DIR, "pqh_compute_nn_fast_shim")

pytestmark = pytest.mark.skipif(not (os.path.exists(SHIM) and os.path.exists(REF_CLI)),
                                reason="oracle/lib shim or oracle/_ref not built")

CASES = [
    ("ties", 3000, 16, 8, []),                                   # default: 243 blocks
    ("sift_like", 6000, 24, 16, ["--num-dims", "2"]),
    ("gist_like", 4000, 12, 5, ["--num-dims", "3", "--num-blocks-per-dim", "2",
                                "--block-overlap-fraction", "0.5"]),
    ("sift_like", 300, 8, 24, ["--num-dims", "4"]),              # blocks smaller than num_nn
    ("gist_like", 2000, 10, 4, ["--num-dims", "1", "--num-blocks-per-dim", "5",
                                "--block-overlap-fraction", "0"]),
]


@pytest.mark.parametrize("merge", [[], ["--host-merge"]], ids=["device-merge-api", "host-merge"])
@pytest.mark.parametrize("gen,n,d,k,flags", CASES)
def test_partitioned_mode_bytes(tmp_path, gen, n, d, k, flags, merge):
    """Default: the blocks go through pqh_knn_heaps_* / ..._search_block_merge (the stand-in
    implements them on the CPU); --host-merge: the driver's own heap code merges them."""
    x = make_golden.make_input(gen, n, d, 70 + n)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    extra = [*flags, "--no-delete-temp-file", "--with-blocks-stat"]
    ours, out = _run_partitioned(SHIM, inp, tmp_path / "ours", k,
                                 [*extra, *merge, "--blocks-info-cache", str(tmp_path / "ours.cache")])
    ref, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref", k,
                              [*extra, "--blocks-info-cache", str(tmp_path / "ref.cache")])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs", "blocks_stat.txt"):
        assert _sha(ours + name) == _sha(ref + name), f"{name} differs"
    assert _sha(str(tmp_path / "ours.cache")) == _sha(str(tmp_path / "ref.cache"))
    again, out2 = _run_partitioned(SHIM, inp, tmp_path / "again", k,
                                   [*flags, *merge, "--blocks-info-cache", str(tmp_path / "ref.cache")])
    assert "Blocks info: Loaded" in out2
    assert _sha(again + "nn_indices.ivecsl") == _sha(ref + "nn_indices.ivecsl")
    assert not os.path.exists(again + "tmp.vecs")


def test_single_block_mode_bytes(tmp_path):
    x = make_golden.make_input("ties", 1500, 16, 3)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    ours = _run(SHIM, inp, tmp_path / "ours", 12, ["--no-delete-temp-file", "--with-blocks-stat"])
    ref = _run(REF_CLI, inp, tmp_path / "ref", 12, ["--no-delete-temp-file", "--with-blocks-stat"])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs", "blocks_stat.txt"):
        assert _sha(ours + name) == _sha(ref + name), f"{name} differs"


@pytest.mark.parametrize("merge", [[], ["--host-merge"]], ids=["device-merge-api", "host-merge"])
def test_resume_from_temp_file_bytes(tmp_path, merge):
    n, d, k = 4000, 12, 10
    x = make_golden.make_input("sift_like", n, d, 5)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    flags = ["--num-dims", "2", "--no-delete-temp-file"]
    ours, _ = _run_partitioned(SHIM, inp, tmp_path / "ours", k,
                               [*flags, *merge, "--blocks-from", "3"])
    ref, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref", k, [*flags, "--blocks-from", "3"])
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(ours + name) == _sha(ref + name), f"leg 1: {name} differs"
    for who in ("ours2", "ref2"):
        os.makedirs(tmp_path / who)
        shutil.copy(ref + "tmp.vecs", str(tmp_path / who / "tmp.vecs"))
    more = [*flags, "--blocks-from", "6", "--no-init-temp-file"]
    ours2, _ = _run_partitioned(SHIM, inp, tmp_path / "ours2", k, [*more, *merge])
    ref2, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref2", k, more)
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl", "tmp.vecs"):
        assert _sha(ours2 + name) == _sha(ref2 + name), f"leg 2: {name} differs"


def test_blocks_to_keeps_the_reference_arithmetic(tmp_path):
    """--blocks-to J inside the range leaves NO blocks in the reference (compute_nn_fast.c:
    549-555: end is pulled back to --blocks-from); the driver writes the same empty result."""
    x = make_golden.make_input("sift_like", 800, 8, 9)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    flags = ["--num-dims", "2", "--blocks-to", "5"]
    ours, _ = _run_partitioned(SHIM, inp, tmp_path / "ours", 4, flags)
    ref, _ = _run_partitioned(REF_CLI, inp, tmp_path / "ref", 4, flags)
    for name in ("nn_indices.ivecsl", "nn_dist.fvecsl"):
        assert _sha(ours + name) == _sha(ref + name), f"{name} differs"


def test_blocks_per_pass_only_groups_blocks(tmp_path):
    """--num-dimensions-at-pass P makes the reference load B^P blocks per pass over the file,
    sharing a prefix test (fast_nn_block_loader2.c:111-118); the result cannot depend on it."""
    x = make_golden.make_input("gist_like", 2500, 10, 17)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    flags = ["--num-dims", "3"]
    plain, _ = _run_partitioned(SHIM, inp, tmp_path / "plain", 6, flags)
    for p in ("1", "2", "3"):
        grouped, _ = _run_partitioned(SHIM, inp, tmp_path / ("pass" + p), 6,
                                      [*flags, "--num-dimensions-at-pass", p])
        for name in ("nn_indices.ivecsl", "nn_dist.fvecsl"):
            assert _sha(grouped + name) == _sha(plain + name), f"P={p}: {name} differs"


def test_driver_rejects_bad_input(tmp_path):
    import subprocess
    x = make_golden.make_input("sift_like", 50, 8, 1)
    inp = str(tmp_path / "in.fvecs")
    vecs_io.write_xvecs(inp, x, "f")
    tpl = str(tmp_path) + "/"

    def run(*args):
        return subprocess.run([SHIM, *args], capture_output=True, text=True, cwd=str(tmp_path))

    assert run(inp, tpl).returncode == 1                                   # too few arguments
    assert run(inp, tpl, "4", "--no-such-flag").returncode == 1
    assert run(str(tmp_path / "missing.fvecs"), tpl, "4").returncode == 1
    assert run(inp, tpl, "0").returncode == 1                               # num_nn < 1
    assert run(inp, tpl, "4", "--num-dims", "9").returncode == 1            # more splits than dims
    p = run(inp, tpl, "4", "--num-dims", "1", "--no-init-temp-file", "--temp-file",
            str(tmp_path / "absent.vecs"))
    assert p.returncode == 1 and "cannot read" in p.stderr
    with open(str(tmp_path / "ragged.fvecs"), "wb") as f:                   # not whole rows
        f.write(open(inp, "rb").read()[:-3])
    assert run(str(tmp_path / "ragged.fvecs"), tpl, "4").returncode == 1


@pytest.mark.parametrize("name", ["cli_single_block", "cli_partitioned"])
def test_golden_digests(tmp_path, name):
    _golden_case(SHIM, tmp_path, name)
