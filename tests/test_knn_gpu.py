"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Bar: ids bit-exact, distances bit-exact (the library reproduces the scalar FP32
sequence of get_real_dist, compute_nn_fast.c:304-311, so no tolerance is needed).
"""
import numpy as np
import pytest

from pq_huffman_b200 import knn, synth

pytestmark = pytest.mark.gpu

ALGOS = [knn.ALGO_EXACT, knn.ALGO_AUTO]


def assert_rows_equal(got, exp):
    (ga, gd), (ea, ed) = got, exp
    assert ga.shape == ea.shape
    bad = np.nonzero((ga != ea).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} rows differ, first {bad[:5]}: {ga[bad[:1]]} vs {ea[bad[:1]]}"
    assert np.array_equal(gd.view(np.uint32), ed.view(np.uint32))


@pytest.fixture(autouse=True)
def _reset_algo():
    yield
    knn.set_algo(knn.ALGO_AUTO)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("gen,n,d,k", [
    ("sift", 5000, 128, 17),
    ("sift", 1000, 32, 5),
    ("gist", 3000, 960, 17),
    ("deep", 6000, 96, 33),
    ("uniform", 777, 5, 3),
    ("uniform", 300, 100, 64),
])
def test_self_join_matches_oracle(oracle, algo, gen, n, d, k):
    x = {"sift": synth.sift_like, "gist": synth.gist_like, "deep": synth.deep_like,
         "uniform": synth.uniform}[gen](n, d, seed=n + d)
    knn.set_algo(algo)
    assert_rows_equal(knn.knn_full_thread(x, None, k), oracle.knn(x, None, k))


@pytest.mark.parametrize("algo", ALGOS)
def test_distinct_queries(oracle, algo):
    b = synth.sift_like(4000, 64, seed=1)
    v = synth.sift_like(333, 64, seed=2)
    knn.set_algo(algo)
    assert_rows_equal(knn.knn_full_thread(b, v, 10), oracle.knn(b, v, 10))


@pytest.mark.parametrize("algo", ALGOS)
def test_ties_and_duplicates(oracle, algo):
    x = synth.with_duplicates_and_ties(synth.sift_like(3000, 16, seed=3), n_dups=20, n_ties=20)
    knn.set_algo(algo)
    got = knn.knn_full_thread(x, None, 9)
    assert np.array_equal(got[0][:, 0], np.arange(3000))
    assert_rows_equal(got, oracle.knn(x, None, 9))


@pytest.mark.parametrize("algo", ALGOS)
def test_all_rows_identical(oracle, algo):
    """Worst case for ties: every distance is 0; order must be by position."""
    x = np.tile(synth.sift_like(1, 24, seed=9), (200, 1))
    knn.set_algo(algo)
    assert_rows_equal(knn.knn_full_thread(x, None, 8), oracle.knn(x, None, 8))


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n,k", [(1, 1), (2, 2), (5, 8), (64, 64), (65, 2), (129, 130)])
def test_small_and_ragged(oracle, algo, n, k):
    x = synth.uniform(n, 9, seed=n)
    knn.set_algo(algo)
    assert_rows_equal(knn.knn_full_thread(x, None, k), oracle.knn(x, None, k))


def test_empty_inputs():
    x = np.zeros((0, 8), dtype=np.float32)
    a, d = knn.knn_full_thread(x, None, 3)
    assert a.shape == (0, 3)
    b = synth.uniform(10, 8)
    a, d = knn.knn_full_thread(b, x, 3)
    assert a.shape == (0, 3)
    a, d = knn.knn_full_thread(x, b, 3)          # empty base: nothing to report
    assert np.all(a == -1) and np.all(np.isinf(d))


@pytest.mark.parametrize("algo", ALGOS)
def test_row_slices_tile_the_full_result(oracle, algo):
    x = synth.sift_like(5000, 128, seed=4)
    knn.set_algo(algo)
    full = oracle.knn(x, None, 17)
    for lo, hi in [(0, 1), (0, 2500), (2500, 5000), (4999, 5000), (1234, 1301)]:
        a, d = knn.search_rows(x, None, 17, lo, hi)
        assert_rows_equal((a, d), (full[0][lo:hi], full[1][lo:hi]))


def test_certificate_failure_falls_back_to_exact(oracle):
    """Data the FP16 tensor-core pass cannot resolve: all points within 1e-3 of a large common
    offset, so FP16 rounding (ulp 2^-7 at 10.0) swamps every difference.  The certificate must
    refuse those rows and the exact kernel must redo them; the answer stays bit-exact."""
    rng = np.random.default_rng(5)
    x = (10.0 + 1e-3 * rng.normal(size=(4000, 64))).astype(np.float32)
    knn.set_algo(knn.ALGO_AUTO)
    got = knn.knn_full_thread(x, None, 9)
    st = knn.last_stats()
    assert st["rows_exact"] > 0, "expected the certificate to reject rows on this input"
    assert st["rows_tensor"] + st["rows_exact"] == 4000
    assert_rows_equal(got, oracle.knn(x, None, 9))


def test_mixed_certified_and_fallback_rows(oracle):
    """Half the set is well separated (certifies), half is the degenerate cloud above."""
    rng = np.random.default_rng(6)
    a = synth.sift_like(2500, 64, seed=12)
    b = (100.0 + 1e-2 * rng.normal(size=(2500, 64))).astype(np.float32)
    x = np.ascontiguousarray(np.concatenate([a, b])[rng.permutation(5000)])
    got = knn.knn_full_thread(x, None, 12)
    st = knn.last_stats()
    assert 0 < st["rows_exact"] < 5000 and st["rows_tensor"] > 0
    assert_rows_equal(got, oracle.knn(x, None, 12))


@pytest.mark.parametrize("gen,n,d,k", [("sift", 6000, 128, 17), ("gist", 4000, 960, 17),
                                       ("deep", 6000, 96, 33), ("uniform", 5000, 272, 9),
                                       ("gist", 5000, 100, 17)])
def test_certificate_error_bound_dominates_real_error(tc_env, gen, n, d, k):
    """The certificate assumes |tensor score - true score| <= eps (FP16 rounding + the tensor
    core's FP32 accumulation).  Measure the real error of every candidate score against an FP64
    dot product on the device and check that the bound holds with room to spare."""
    # this test is about the FP16 path, also for uint8 data
    tc_env(PQH_TC_CHECK_EPS=1, PQH_TC_I8=0)
    x = {"sift": synth.sift_like, "gist": synth.gist_like, "deep": synth.deep_like,
         "uniform": synth.uniform}[gen](n, d, seed=7 * n + d)
    knn.knn_full_thread(x, None, k)
    st = knn.last_stats()
    assert st["rows_tensor"] > 0
    assert st["eps_ratio_max"] <= 1.0, st
    if gen == "sift":
        assert st["inexact_operands"] == 0      # integer data: FP16 operands are exact
        assert st["eps_ratio_max"] == 0.0       # ... and so is the tensor-core accumulation
    else:
        assert st["inexact_operands"] == 1


def test_stats_report_launches(oracle):
    x = synth.sift_like(2000, 128, seed=6)
    knn.knn_full_thread(x, None, 17)
    st = knn.last_stats()
    assert st["rows"] == 2000 and st["launches"] >= 2
    assert st["rows_tensor"] + st["rows_exact"] >= 2000


def test_medium_property_100k():
    """configs[0]-sized (100k x 128, k=16+1): too slow for the oracle in a unit test, so
    check size-independent properties: self first, ascending, symmetric distances,
    and a sampled set of rows against the oracle's row-slice entry."""
    from tests import oracle_ffi
    x = synth.sift_like(100_000, 128)
    a, d = knn.knn_full_thread(x, None, 17)
    assert np.array_equal(a[:, 0], np.arange(100_000))
    assert np.all(np.diff(d, axis=1) >= 0)
    assert a.min() >= 0 and a.max() < 100_000
    rows = np.random.default_rng(0).choice(100_000, 64, replace=False)
    orc = oracle_ffi.Oracle()
    for r in rows:
        ea, ed = orc.knn_rows(x, None, 17, int(r), int(r) + 1)
        assert np.array_equal(a[r], ea[0]) and np.array_equal(d[r], ed[0])


@pytest.mark.parametrize("n,d,k", [(2048, 128, 17), (2049, 128, 17), (2175, 64, 9), (4097, 128, 2),
                                   (3000, 1, 4), (3000, 128, 34), (3000, 64, 65), (3000, 64, 66), (3000, 40, 100)])
def test_tensor_path_boundaries(oracle, n, d, k):
    """Sizes around the tensor-core path's limits: the smallest base it takes (2048), ragged last
    tiles with the row's own column inside them, k = 2, d = 1, the largest k it takes (65 ->
    kcap 64) and the first k it leaves to the exact kernel (66), and a large k."""
    x = synth.sift_like(n, d, seed=3 * n + k)
    got = knn.knn_full_thread(x, None, k)
    st = knn.last_stats()
    assert st["rows_tensor"] + st["rows_exact"] >= n
    if k <= 65:
        assert st["rows_tensor"] > 0
    else:
        assert st["rows_tensor"] == 0
    assert_rows_equal(got, oracle.knn(x, None, k))


KNOBS = [
    dict(PQH_TC_MMA=1),
    dict(PQH_TC_ORDER=0),
    dict(PQH_TC_ORDER=0, PQH_TC_MMA=1, PQH_TC_EARLY=0),
    dict(PQH_TC_EARLY=0),
    dict(PQH_TC_EARLY=0, PQH_TC_I8=0),
    dict(PQH_TC_ORDER=0, PQH_TC_I8=0),
    dict(PQH_TC_STREAM=1),
    dict(PQH_TC_STREAM=1, PQH_TC_ORDER=0),
    dict(PQH_TC_I8=0, PQH_TC_MMA=1),
    dict(PQH_TC_SLICE_MB=64),
    dict(PQH_TC_DYN=0),
    dict(PQH_TC_WIDE_D=4096),
    dict(PQH_TC_DYN=0, PQH_TC_MMA=1, PQH_TC_I8=0),
]


_KNOB_CASES = {}


def _knob_cases(oracle):
    """Inputs and oracle answers shared by every configuration (computed once)."""
    if not _KNOB_CASES:
        x = synth.with_duplicates_and_ties(synth.sift_like(40000, 128, seed=77), n_dups=6, n_ties=6)
        g = synth.gist_like(33000, 96, seed=78)
        b = synth.deep_like(34000, 96, seed=79)
        v = synth.deep_like(700, 96, seed=80)
        _KNOB_CASES.update(x=x, g=g, b=b, v=v, ex=oracle.knn(x, None, 17), eg=oracle.knn(g, None, 9),
                           ebv=oracle.knn(b, v, 33))
    return _KNOB_CASES


@pytest.mark.parametrize("knobs", KNOBS, ids=lambda k: ",".join(f"{a[7:]}={b}" for a, b in k.items()))
def test_kernel_configurations_stay_exact(tc_env, oracle, knobs):
    """Every selectable configuration of the tensor-core path -- one or two MMA issuers, stage
    release before or after the filtering, locality order on or off, work items on demand or round
    robin, the streaming kernel forced on a small d, small row slices -- must produce the oracle's
    answer: the knobs only decide how fast.  (Defaults: two issuers, early release, locality
    order, items on demand.)"""
    c = _knob_cases(oracle)
    tc_env(**knobs)
    got = knn.knn_full_thread(c["x"], None, 17)
    assert knn.last_stats()["rows_tensor"] > 0
    assert_rows_equal(got, c["ex"])
    assert_rows_equal(knn.knn_full_thread(c["g"], None, 9), c["eg"])
    assert_rows_equal(knn.knn_full_thread(c["b"], c["v"], 33), c["ebv"])
    a, dd = knn.search_rows(c["x"], None, 17, 5000, 32345)
    assert_rows_equal((a, dd), (c["ex"][0][5000:32345], c["ex"][1][5000:32345]))


def test_row_slices_bound_the_scratch(tc_env, oracle):
    """The first pass runs in row slices so that the candidate lists never exceed the budget
    ($PQH_TC_SLICE_MB); the workspace the library asks for must stay bounded as n grows."""
    tc_env(PQH_TC_SLICE_MB=64)
    x = synth.sift_like(90000, 64, seed=5)
    got = knn.knn_full_thread(x, None, 17)
    rows = np.random.default_rng(1).choice(90000, 48, replace=False)
    for r in rows:
        ea, ed = oracle.knn_rows(x, None, 17, int(r), int(r) + 1)
        assert np.array_equal(got[0][r], ea[0]) and np.array_equal(got[1][r], ed[0])
    tc_env(PQH_TC_SLICE_MB=None)
    # Deep10M-shaped: operands 1.9 GB + lists of one slice, not 41 GB of lists
    assert knn.workspace_bytes(10_000_000, 10_000_000, 96, 33) <= 8 * 2**30


def test_list_capacity_with_long_candidate_lists(oracle):
    """k = 33 distinct queries keep L = 96 candidates per row: a shrink leaves up to 144 entries and
    a tile may add 128 more, so the lists need the 512-entry capacity (a 256-entry list overflowed
    here, intermittently, in the first tiles of a sweep)."""
    b = synth.deep_like(40000, 96, seed=91)
    v = synth.deep_like(1500, 96, seed=92)
    exp = oracle.knn(b, v, 33)
    for _ in range(3):
        assert_rows_equal(knn.knn_full_thread(b, v, 33), exp)


def test_mass_duplicates_through_tensor_path(oracle):
    """2500 copies of one vector: far more exact ties than the candidate lists hold, so the
    certificate cannot accept those rows (a non-candidate is exactly as close as the k-th
    neighbour) and the exact kernel has to order them by position."""
    rng = np.random.default_rng(3)
    one = synth.sift_like(1, 48, seed=2)
    x = np.concatenate([np.tile(one, (2500, 1)), synth.sift_like(700, 48, seed=5)])
    x = np.ascontiguousarray(x[rng.permutation(len(x))])
    got = knn.knn_full_thread(x, None, 9)
    st = knn.last_stats()
    assert st["rows_exact"] >= 2500
    assert_rows_equal(got, oracle.knn(x, None, 9))


def test_distinct_queries_ragged_through_tensor_path(oracle):
    b = synth.gist_like(5001, 72, seed=8)
    v = synth.gist_like(777, 72, seed=9)
    got = knn.knn_full_thread(b, v, 12)
    assert knn.last_stats()["rows_tensor"] > 0
    assert_rows_equal(got, oracle.knn(b, v, 12))
    a, d = knn.search_rows(b, v, 12, 100, 613)
    full = oracle.knn(b, v, 12)
    assert_rows_equal((a, d), (full[0][100:613], full[1][100:613]))


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs")
def test_in_library_multi_gpu_row_split(oracle):
    """PQH_DEVICES-style context over two GPUs: knn_full_thread splits the query rows, every
    device gets the base over its own link, results land in the caller's arrays."""
    x = synth.sift_like(30000, 128, seed=44)
    try:
        knn.init([0, 1])
        got = knn.knn_full_thread(x, None, 17)
        st = knn.last_stats()
        assert st["rows"] == 30000
    finally:
        knn.init([0])
    assert_rows_equal(got, oracle.knn(x, None, 17))


def test_full_size_sift1m_properties():
    """BASELINE configs[1] at full size (1M x 128, k = 16 + self).  The oracle cannot do 10^12
    pairs in a test, so check size-independent properties on all rows and the oracle itself on
    a sample of rows (its row-slice entry point does 1M distances per row)."""
    from tests import oracle_ffi
    n, d, k = 1_000_000, 128, 17
    x = synth.sift_like(n, d)
    a, dist = knn.knn_full_thread(x, None, k)
    st = knn.last_stats()
    assert st["rows"] == n and st["rows_tensor"] + st["rows_exact"] >= n
    assert np.array_equal(a[:, 0], np.arange(n, dtype=np.int32))     # self first
    assert np.all(dist[:, 0] == 0.0)
    assert np.all(np.diff(dist, axis=1) >= 0)                        # ascending
    assert a.min() >= 0 and a.max() < n
    srt = np.sort(a, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])                         # k distinct ids per row
    orc = oracle_ffi.Oracle()
    rng = np.random.default_rng(1)
    for r in rng.choice(n, 48, replace=False):
        ea, ed = orc.knn_rows(x, None, k, int(r), int(r) + 1)
        assert np.array_equal(a[r], ea[0]), r
        assert np.array_equal(dist[r].view(np.uint32), ed[0].view(np.uint32)), r
    # reported distances are the scalar definition for the reported pairs
    for r in rng.choice(n, 200, replace=False):
        j = int(a[r, 5])
        assert np.float32(orc.dist(x[r], x[j])).view(np.uint32) == dist[r, 5].view(np.uint32)


@pytest.mark.parametrize("algo", ALGOS)
def test_resident_dataset_block_searches(oracle, algo):
    """pqh_knn_dataset_search_block == knn_full_thread on the gathered block (run_block,
    compute_nn_fast.c:469-483): large blocks (tensor-core path), small ones (exact kernel),
    unsorted row ids, single-row and empty blocks, repeated calls on one resident data set."""
    n, d = 30000, 32
    x = synth.with_duplicates_and_ties(synth.sift_like(n, d, seed=12), n_dups=30, n_ties=30)
    knn.set_algo(algo)
    knn.dataset_load(x)
    rng = np.random.default_rng(3)
    try:
        for m, k, ordered in [(9000, 17, True), (500, 9, True), (4000, 17, False), (2, 2, True),
                              (1, 1, True), (n, 5, True)]:
            ids = rng.choice(n, size=m, replace=False).astype(np.int32)
            if ordered:
                ids.sort()
            got = knn.dataset_search_block(ids, k)
            assert_rows_equal(got, oracle.knn(x[ids], None, k))
        a, dd = knn.dataset_search_block(np.empty(0, dtype=np.int32), 3)
        assert a.shape == (0, 3)
        st = knn.last_stats()
        assert st["rows"] == 9000 + 500 + 4000 + 2 + 1 + n          # stats accumulate
        with pytest.raises(knn.PqhError):
            knn.dataset_search_block(np.array([0, n], dtype=np.int32), 2)
    finally:
        knn.dataset_release()
    with pytest.raises(knn.PqhError):
        knn.dataset_search_block(np.array([0, 1], dtype=np.int32), 2)


def test_resident_dataset_from_fvecs_image(oracle):
    n, d = 6000, 24
    x = synth.gist_like(n, d, seed=8)
    img = np.empty((n, d + 1), dtype=np.float32)
    img[:, 0] = np.array([d], dtype=np.int32).view(np.float32)[0]
    img[:, 1:] = x
    flat = img.reshape(-1)
    knn.dataset_load((flat[1:], n, d), ld=d + 1)
    try:
        ids = np.arange(100, 5100, dtype=np.int32)
        assert_rows_equal(knn.dataset_search_block(ids, 6), oracle.knn(x[ids], None, 6))
    finally:
        knn.dataset_release()


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs")
def test_resident_dataset_blocks_on_two_gpus_concurrently(oracle):
    """Whole blocks round-robin over the devices (small-block alternative): two host
    threads search different blocks of the same resident data set on slots 0 and 1 at once."""
    import threading
    n, d, k = 40000, 64, 11
    x = synth.sift_like(n, d, seed=21)
    rng = np.random.default_rng(9)
    blocks = [np.sort(rng.choice(n, size=m, replace=False)).astype(np.int32)
              for m in (12000, 9000, 700, 15000, 3000, 11000)]
    out = [None] * len(blocks)
    try:
        knn.init([0, 1])
        knn.dataset_load(x)

        def work(slot):
            for i in range(slot, len(blocks), 2):
                out[i] = knn.dataset_search_block(blocks[i], k, slot=slot)

        th = [threading.Thread(target=work, args=(s,)) for s in (0, 1)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert knn.last_stats()["rows"] == sum(len(b) for b in blocks)
    finally:
        knn.dataset_release()
        knn.init([0])
    for ids, got in zip(blocks, out):
        assert_rows_equal(got, oracle.knn(x[ids], None, k))


def test_integer_tensor_path(tc_env, oracle):
    """uint8-valued data (SIFT-shaped) takes the exact INT8 tensor-core path (stats:
    inexact_operands == 2), PQH_TC_I8=0 keeps it on the FP16 path; data the 8-bit operands or the
    fold columns cannot hold falls back to FP16 by itself.  Same bits either way."""
    x = synth.with_duplicates_and_ties(synth.sift_like(9000, 128, seed=31), n_dups=10, n_ties=10)
    exp = oracle.knn(x, None, 17)
    assert_rows_equal(knn.knn_full_thread(x, None, 17), exp)
    assert knn.last_stats()["inexact_operands"] == 2
    tc_env(PQH_TC_I8=0)
    assert_rows_equal(knn.knn_full_thread(x, None, 17), exp)
    assert knn.last_stats()["inexact_operands"] == 0
    tc_env(PQH_TC_I8=None)
    # distinct queries, ragged sizes, odd dimension (padding to 32-byte K steps)
    b = synth.sift_like(5000, 100, seed=32)
    v = synth.sift_like(777, 100, seed=33)
    assert_rows_equal(knn.knn_full_thread(b, v, 9), oracle.knn(b, v, 9))
    assert knn.last_stats()["inexact_operands"] == 2
    # |y|^2 beyond what the fold columns hold: bright rows
    rng = np.random.default_rng(5)
    bright = rng.integers(200, 256, size=(4000, 128)).astype(np.float32)
    assert_rows_equal(knn.knn_full_thread(bright, None, 9), oracle.knn(bright, None, 9))
    assert knn.last_stats()["inexact_operands"] == 0
    # one non-integer value anywhere: the whole call stays on FP16
    y = synth.sift_like(4000, 64, seed=34)
    y[1234, 7] += 0.5
    assert_rows_equal(knn.knn_full_thread(y, None, 9), oracle.knn(y, None, 9))
    assert knn.last_stats()["inexact_operands"] in (0, 1)
    z = synth.sift_like(4000, 64, seed=35)
    z[17, 3] = -1.0
    assert_rows_equal(knn.knn_full_thread(z, None, 9), oracle.knn(z, None, 9))
    assert knn.last_stats()["inexact_operands"] in (0, 1)


def test_device_side_merge_follows_the_reference_heap_rules(oracle):
    """pqh_knn_heaps_* + pqh_knn_dataset_search_block_merge against a pure-Python restatement
    of run_merge_block / fast_nn_heap_push / fast_nn_heap_sort (tests/ref_semantics.py, pinned
    against the reference's compiled functions in test_oracle_pins.py): overlapping blocks with
    duplicates and ties, a block smaller than k, a single-row block, a resumed image."""
    from tests import ref_semantics as rs
    n, d, k = 3000, 16, 6
    x = synth.with_duplicates_and_ties(synth.sift_like(n, d, seed=21), n_dups=40, n_ties=40)
    rng = np.random.default_rng(2)
    blocks = [np.sort(rng.choice(n, m, replace=False)).astype(np.int32)
              for m in (2600, 900, 4, 1, 2500, 700)]
    heaps = [[[rs.UINT_MAX, np.float32(np.inf)] for _ in range(k)] for _ in range(n)]

    def merge_host(ids):
        m = len(ids)
        k1 = min(k, m - 1) + 1
        if m < 2:
            return k1
        ea, ed = oracle.knn(x[ids], None, k1)
        for r in range(m):
            for c in range(1, k1):
                rs.heap_push(heaps[ids[r]], [int(ids[ea[r, c]]), np.float32(ed[r, c])], k)
        return k1

    knn.dataset_load(x)
    try:
        knn.heaps_create(n, k)
        for t, ids in enumerate(blocks[:3]):
            knn.dataset_search_block_merge(ids, merge_host(ids), t)
        image, _, _ = knn.heaps_export(n, k)
        exp = np.array([[tuple(it) for it in h] for h in heaps], dtype=knn.ITEM)
        assert np.array_equal(image["id"], exp["id"])
        assert np.array_equal(image["dist"].view(np.uint32), exp["dist"].view(np.uint32))
        # resume from that image with the remaining blocks
        knn.heaps_create(n, k, image)
        for t, ids in enumerate(blocks[3:]):
            knn.dataset_search_block_merge(ids, merge_host(ids), t)
        image, ids_out, dists_out = knn.heaps_export(n, k)
    finally:
        knn.heaps_release()
        knn.dataset_release()
    exp = np.array([[tuple(it) for it in h] for h in heaps], dtype=knn.ITEM)
    assert np.array_equal(image["id"], exp["id"])
    assert np.array_equal(image["dist"].view(np.uint32), exp["dist"].view(np.uint32))
    for r in rng.choice(n, 300, replace=False):
        ei, ed = rs.heap_sort([list(it) for it in heaps[r]], k)
        assert list(ids_out[r]) == [int(v) for v in ei]
        assert np.array_equal(dists_out[r].view(np.uint32),
                              np.array(ed, dtype=np.float32).view(np.uint32))


def _random_case(seed):
    """One seeded shape around the limits of the locality order: a base just below / above the
    size from which cells are used (32768), row slices of it as queries (a tile then spans many
    cells: the sweep window is sized per tile), distinct queries, small and large k, d on either
    side of the widths the kernels switch at (64: warp-wide re-rank fetches; 144: streaming)."""
    rng = np.random.default_rng(seed)
    n = int(rng.choice([32767, 32768, 33000, 40000, 50001, 65536]))
    d = int(rng.choice([8, 24, 63, 64, 96, 128, 144, 148, 200]))
    k = int(rng.choice([1, 2, 9, 17, 33, 65]))
    gen = [synth.sift_like, synth.deep_like, synth.uniform][int(rng.integers(0, 3))]
    mode = ["self", "slice", "distinct"][int(rng.integers(0, 3))]
    return n, d, k, gen, mode, rng


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_around_the_locality_order(oracle, seed):
    n, d, k, gen, mode, rng = _random_case(seed)
    x = gen(n, d, seed=seed)
    if seed % 4 == 0:
        x = synth.with_duplicates_and_ties(x, n_dups=50, n_ties=50)
    rows = n
    if mode == "self":
        assert_rows_equal(knn.knn_full_thread(x, None, k), oracle.knn(x, None, k))
    elif mode == "slice":
        lo = int(rng.integers(0, n - 3000))
        rows = int(rng.integers(1, 3000))
        exp = oracle.knn_rows(x, None, k, lo, lo + rows)
        assert_rows_equal(knn.search_rows(x, None, k, lo, lo + rows), exp)
    else:
        rows = int(rng.integers(1, 2500))
        v = gen(rows, d, seed=seed + 100)
        assert_rows_equal(knn.knn_full_thread(x, v, k), oracle.knn(x, v, k))
    kcap = k - (mode != "distinct")          # neighbours asked for beyond the row itself
    if 1 <= kcap <= 64 and rows >= 256:
        assert knn.last_stats()["rows_tensor"] > 0
