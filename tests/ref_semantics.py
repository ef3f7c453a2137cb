"""Pure-Python restatement of the reference's merge/export rules (small cases only).

Follows fast_nn_temp_file.c:11-63 (max-heap push with duplicate-id rejection and
strict `<`, heap sort), compute_nn_fast.c:485-508 (run_merge_block: skip column
0, push columns 1..num_nn_real) and fast_nn_temp_file.c:83-129,253-301 (temp
file initialised to {UINT_MAX, +inf}; rows heap-sorted ascending on export).
Pinned against the reference's compiled functions in tests/test_oracle_pins.py.
"""
import numpy as np

UINT_MAX = 0xFFFFFFFF


def heap_push_impl(heap, item, num_nn):
    """fast_nn_heap_push_impl, fast_nn_temp_file.c:23-46. heap: list of [index, dist]."""
    pos = 0
    while True:
        left = 2 * (pos + 1) - 1
        right = 2 * (pos + 1)
        swap = pos
        if right < num_nn and heap[right][1] > heap[left][1] and heap[right][1] > item[1]:
            swap = right
        elif left < num_nn and heap[left][1] > item[1]:
            swap = left
        else:
            break
        heap[pos] = heap[swap]
        pos = swap
    heap[pos] = list(item)


def heap_push(heap, item, num_nn):
    """fast_nn_heap_push, fast_nn_temp_file.c:48-54."""
    if item[1] >= heap[0][1] or any(h[0] == item[0] for h in heap[:num_nn]):
        return
    heap_push_impl(heap, item, num_nn)


def heap_sort(heap, num_nn):
    """fast_nn_heap_sort, fast_nn_temp_file.c:56-63. Returns (indices, dists) ascending."""
    idx = [0] * num_nn
    dst = [0.0] * num_nn
    n = num_nn
    while n > 0:
        n -= 1
        idx[n] = heap[0][0]
        dst[n] = heap[0][1]
        heap_push_impl(heap, heap[n], n)
    return idx, dst


def single_block_result(assign, dis, num_nn):
    """What compute_nn_fast --num-dims 0 writes for one block = whole dataset:
    returns (uint32[N x num_nn], float32[N x num_nn])."""
    n, k1 = assign.shape
    num_nn_real = k1 - 1
    out_i = np.empty((n, num_nn), dtype=np.uint32)
    out_d = np.empty((n, num_nn), dtype=np.float32)
    for r in range(n):
        heap = [[UINT_MAX, np.float32(np.inf)] for _ in range(num_nn)]
        for c in range(1, num_nn_real + 1):
            heap_push(heap, [int(assign[r, c]), np.float32(dis[r, c])], num_nn)
        idx, dst = heap_sort(heap, num_nn)
        out_i[r] = idx
        out_d[r] = dst
    return out_i, out_d


def joint_histograms(pq_codes, nn_indices, k_star, nn_index):
    """collect_stats_at_kth_nn (nn_huffman_estimator.c:69-107): for every vector whose neighbour of
    rank nn_index is known (index != (vector_id_t)-1) and every sub-quantiser i,
    stats[i][code of the vector][code of the neighbour] += 1.  Returns (stats float64[m, k*, k*],
    number of vectors with a known neighbour)."""
    pq = np.asarray(pq_codes, dtype=np.uint8)
    nn = np.asarray(nn_indices).view(np.uint32)
    n, m = pq.shape
    col = nn[:, nn_index]
    known = col != np.uint32(0xFFFFFFFF)
    src = pq[known].astype(np.int64)
    dst = pq[col[known].astype(np.int64)].astype(np.int64)
    stats = np.zeros((m, k_star, k_star), dtype=np.float64)
    for i in range(m):
        np.add.at(stats[i], (src[:, i], dst[:, i]), 1.0)
    return stats, int(known.sum())
