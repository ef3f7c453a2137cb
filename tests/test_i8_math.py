"""The arithmetic behind the INT8 tensor-core path (pq_huffman_b200/csrc/knn_tc_i8.cuh), restated
with numpy integers: the score identity the certificate relies on, and the s8 digit decomposition
of the fold value that prep_rows_i8_kernel writes (knn_tc.cu).  CPU only -- the CUDA side is
checked bit for bit against the oracle in tests/test_knn_gpu.py::test_integer_tensor_path."""
import numpy as np

FOLD_FULL = 30            # columns whose A-side constant is 255 (the remaining two: 1)
FOLD_UNIT = 255 * 128


def fold_digits(norm2):
    """s8 digits b[32] with sum(c * b) == -ceil(norm2 / 2) for c = [255]*30 + [1, 1]."""
    g = (int(norm2) + 1) // 2
    q, r = divmod(g, FOLD_UNIT)
    if q >= FOLD_FULL:
        return None                                    # beyond capacity: the call stays on FP16
    b = np.zeros(32, dtype=np.int64)
    b[:q] = -128
    b[q] = -(r // 255)
    r2 = r % 255
    first = min(r2, 127)
    b[30], b[31] = -first, -(r2 - first)
    return b


def test_fold_digits_are_exact_and_fit_s8():
    c = np.array([255] * FOLD_FULL + [1, 1], dtype=np.int64)
    rng = np.random.default_rng(0)
    cases = [0, 1, 2, 254, 255, 256, FOLD_UNIT - 1, FOLD_UNIT, 2 * FOLD_UNIT + 1,
             2 * FOLD_FULL * FOLD_UNIT - 3, 2 * FOLD_FULL * FOLD_UNIT - 2]
    cases += [int(v) for v in rng.integers(0, 2 * FOLD_FULL * FOLD_UNIT - 1, size=2000)]
    for n2 in cases:
        b = fold_digits(n2)
        assert b is not None, n2
        assert b.min() >= -128 and b.max() <= 127
        assert int((c * b).sum()) == -((n2 + 1) // 2), n2
    assert fold_digits(2 * FOLD_FULL * FOLD_UNIT - 1) is None      # first norm that does not fit
    assert fold_digits(128 * 255 * 255) is None                     # brightest possible 128-d row


def test_score_identity_and_certificate_bound():
    rng = np.random.default_rng(1)
    n, d, k, L = 400, 128, 8, 12
    y = rng.integers(0, 256, size=(n, d)).astype(np.int64)
    y[:, d // 2:] //= 8                                             # keep |y|^2 inside the capacity
    x = y[:40]
    g = (np.sum(y * y, axis=1) + 1) // 2
    t = x @ (y - 128).T - g[None, :]                                # what the tensor cores produce
    dist = ((x[:, None, :] - y[None, :, :]) ** 2).sum(axis=2)
    x1 = x.sum(axis=1)
    xn2 = (x * x).sum(axis=1)
    delta = xn2[:, None] - 256 * x1[:, None] - 2 * t - dist
    assert set(np.unique(delta)) <= {0, 1}
    # within a row the score order is the distance order up to that 0/1
    for r in range(x.shape[0]):
        order = np.argsort(-t[r], kind="stable")
        tau = t[r, order[L - 1]]                                    # threshold after a shrink to L
        rest = order[L:]
        lb = xn2[r] - 256 * x1[r] - 2 * tau - 1
        assert (dist[r, rest] >= lb).all()                          # the certificate's premise
        kth = np.sort(dist[r, order[:L]])[k - 1]
        if lb > kth:                                                # accepted rows are complete
            assert kth <= dist[r, rest].min()
            assert set(np.argsort(dist[r], kind="stable")[:k]) <= set(order[:L]) or \
                   np.sort(dist[r])[k - 1] == kth
