/*
 * pqh_knn.h -- C ABI of libpqh_knn.so, the B200 (sm_100a) implementation of the
 * exact k-nearest-neighbour hot path of trukanduk/pq_huffman.
 *
 * Every entry point takes plain pointers and sizes; no CUDA, torch or C++ types
 * cross this boundary.  File:line citations are relative to the reference's
 * src/ directory.
 *
 * Drop-in symbol
 * --------------
 * knn_full_thread() below has exactly the prototype of Yael v438's
 * <yael/nn.h> knn_full_thread, the single unresolved symbol of the reference's
 * compute_nn_fast (call site compute_nn_fast.c:478-481, link line
 * Makefile:5,28-30).  Linking the reference's unmodified objects against
 * libpqh_knn.so instead of -lyael moves the search onto the GPU; see
 * INTEGRATION.md.
 *
 * Error behaviour follows the reference's convention for this path
 * (compute_nn_fast.c:154-156, fast_nn_block_loader2.c:88-91): the void entry
 * points print a message to stderr and exit(1) on an unrecoverable error (no
 * usable CUDA device, CUDA/NCCL failure, out of device memory).  There is NO
 * CPU fallback.  The pqh_* entry points that return int report 0 on success
 * and a negative PQH_E* code otherwise, and pqh_last_error() returns the text.
 */
#ifndef PQH_KNN_H_
#define PQH_KNN_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQH_OK 0
#define PQH_EINVAL (-1)  /* bad argument */
#define PQH_ENODEV (-2)  /* no usable sm_100 CUDA device */
#define PQH_ECUDA (-3)   /* CUDA runtime / driver failure */
#define PQH_ENOMEM (-4)  /* host or device allocation failed */
#define PQH_ECOMM (-5)   /* multi-GPU exchange failed */
#define PQH_ENOTORDER (-6) /* pqh_sort_edges: the comparator is not an order on this data */

/*
 * Replaces: Yael knn_full_thread, called at compute_nn_fast.c:478-481 (dead
 * twin at :416-419) and, through ynumpy.knn, at compute_nn.py:22.
 *
 *   distance_type  must be 2 (squared L2), the only value the reference passes
 *   nq, nb, d      number of queries, base vectors, dimensions
 *   k              neighbours per query (the reference passes min(num_nn,n-1)+1)
 *   b              base,    float32[nb x d] row-major, host memory
 *   v              queries, float32[nq x d] row-major, host memory; v == b
 *                  (same pointer, nq == nb) selects the self-join the reference
 *                  uses: query i is base row i
 *   b_weights      must be NULL (the reference passes NULL)
 *   assign         out, int32[nq x k]: base positions, ascending distance
 *   dis            out, float32[nq x k]: squared L2, bit-identical to the scalar
 *                  FP32 definition get_real_dist (compute_nn_fast.c:304-311)
 *   n_thread       accepted for compatibility; the GPU path ignores it
 *
 * Order within a row: ascending (distance, base position).  For a self-join,
 * column 0 is the query itself with distance 0 (the caller drops column 0 as
 * "self", compute_nn_fast.c:498, compute_nn.py:23-24).  If k > nb the unfilled
 * tail of a row is assign = -1, dis = +inf.
 *
 * Buffers are caller-owned; nothing is retained after return.
 */
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *v, const float *b_weights, int *assign, float *dis,
                     int n_thread);

/*
 * Context control.  The first search creates the context lazily on the devices
 * named by $PQH_DEVICES (comma separated ordinals, default: device 0 only);
 * these calls make it explicit.  Replaces nothing in the reference (it has no
 * device state); they exist so that compute_nn_fast's hundreds of per-block
 * calls (run(), compute_nn_fast.c:556-610) reuse streams, pinned staging rings,
 * NCCL communicators and device buffers instead of re-creating them.
 *
 * A context over several devices creates one NCCL communicator per device
 * (ncclCommInitAll; libnccl.so.2 is resolved at run time).  A host-pointer search
 * then splits the query rows evenly over the devices; every device uploads 1/G of
 * the base over its own PCIe link, one in-place ncclAllGather over NVLink
 * assembles the base on all of them, and each device writes its rows straight into
 * the caller's arrays.  If NCCL cannot be loaded or initialised, pqh_knn_init
 * fails with PQH_ECOMM ($PQH_NO_NCCL=1 opts out: every device then receives the
 * whole base over its own link).
 *
 * The $PQH_TC_* tuning knobs are read when a context is created.
 */
int pqh_knn_init(const int *devices, int n_devices); /* NULL/0: $PQH_DEVICES or {0} */
void pqh_knn_shutdown(void);
int pqh_knn_device_count(void);                      /* devices in the live context, 0 if none */
const char *pqh_last_error(void);

/*
 * Search algorithm selection (debug / test hook; the default is AUTO).
 *   PQH_ALGO_AUTO    tensor-core pass + exact re-rank + certificate, rows that
 *                    fail the certificate are redone by the exact kernel
 *   PQH_ALGO_EXACT   FP32 SIMT kernel only (no tensor cores)
 */
#define PQH_ALGO_AUTO 0
#define PQH_ALGO_EXACT 1
int pqh_knn_set_algo(int algo);

/*
 * Row-sliced search: queries [q_begin, q_end) of v against all nb rows of b.
 * Row q of the result is written at assign + (q - q_begin) * k (and likewise
 * dis).  This is the unit of work one rank of a multi-process run owns
 * (partition by query-block rows); knn_full_thread is
 * pqh_knn_search_rows over [0, nq).  self_join != 0 means query q is base row
 * q (what v == b means for knn_full_thread).
 */
int pqh_knn_search_rows(int nq_total, int nb, int d, int k, const float *b, const float *v,
                        int self_join, int q_begin, int q_end, int *assign, float *dis);

/*
 * Same search with explicit leading dimensions (in floats) of the two host matrices: row i of
 * b starts at b + i * ldb.  Lets a caller hand over vectors still interleaved with other data --
 * e.g. an .fvecs file image, whose rows are [int32 d | d floats] (ldb = d + 1, b pointing one
 * float into the buffer; the layout fast_nn_block_loader2.c:99-108 strips row by row) -- the
 * library strips the padding in the H2D copy itself.
 */
int pqh_knn_search_rows_ld(int nq_total, int nb, int d, int k, const float *b, long long ldb,
                           const float *v, long long ldv, int self_join, int q_begin, int q_end,
                           int *assign, float *dis);

/*
 * Device-resident variant for callers that already hold the vectors in HBM
 * (bench.py's device-timed arm, a future GPU loader): all pointers are device
 * pointers on `device`; the kernels are launched on `stream` (a cudaStream_t
 * passed as void*; NULL = the context's own stream).  The call is NOT fully
 * asynchronous: it synchronises `stream` up to three times (to learn which operand
 * flavour the data allows, how many rows failed the certificate, and -- rarely --
 * how many remained after the second pass); with a caller's stream it returns
 * with the last kernels still in flight, with NULL it returns when the results
 * are complete.  d_work must point to pqh_knn_workspace_bytes(...) bytes of device
 * scratch; the figure is bounded in n (candidate lists are kept for one slice of
 * query rows at a time, $PQH_TC_SLICE_MB, 2 GiB by default).
 */
long long pqh_knn_workspace_bytes(int nq, int nb, int d, int k);
int pqh_knn_search_device(int device, int nq_total, int nb, int d, int k, const float *d_b,
                          const float *d_v, int self_join, int q_begin, int q_end,
                          int *d_assign, float *d_dis, void *d_work, long long work_bytes,
                          void *stream);

/*
 * Resident data set + block searches: the partitioned mode of compute_nn_fast
 * (run(), compute_nn_fast.c:510-618) runs hundreds of self-joins over OVERLAPPING
 * subsets ("blocks", fast_nn_block.h:4-11) of one file, re-reading the file and
 * re-filling a block_t for each (fast_nn_block_loader2.c:86-146).  Here the file
 * is uploaded once to every device of the context and a block is named by the
 * row ids it holds; the rows are gathered on the device.
 *
 *   pqh_knn_dataset_load          rows: float32[n x d] on the host, row i at rows + i*ld
 *                                 (ld = d + 1 for an .fvecs image); replaces any previous
 *                                 data set and zeroes the stats, which then ACCUMULATE
 *                                 over the block searches
 *   pqh_knn_dataset_search_block  exactly knn_full_thread(2, m, m, d, k, blk, blk, NULL,
 *                                 assign, dis, nt) for blk = the m rows row_ids[0..m) in
 *                                 that order (run_block, compute_nn_fast.c:469-483):
 *                                 assign holds positions WITHIN the block, column 0 is the
 *                                 row itself.  `slot` selects the device (index into the
 *                                 context's device list, 0 for one GPU); calls on different
 *                                 slots may run concurrently from different host threads.
 *   pqh_knn_dataset_release       frees the device copies
 *   pqh_knn_host_alloc / _free    page-locked host memory for the result buffers of repeated
 *                                 calls (plain malloc works too, at pageable-copy speed)
 */
int pqh_knn_dataset_load(int n, int d, const float *rows, long long ld);
int pqh_knn_dataset_search_block(int slot, int m, const int *row_ids, int k, int *assign,
                                 float *dis);
void pqh_knn_dataset_release(void);
void *pqh_knn_host_alloc(long long bytes);
void pqh_knn_host_free(void *p);

/*
 * Device-side merge of the partitioned mode.  The reference merges every searched block into
 * per-vector k-slot max-heaps kept in its temp file (run_merge_block, compute_nn_fast.c:485-508;
 * fast_nn_heap_push, fast_nn_temp_file.c:11-54: an item is dropped unless strictly nearer than
 * the current worst, or if its id is already present) and heap-sorts them at the end
 * (temp_file_to_result / fast_nn_heap_sort, fast_nn_temp_file.c:56-63,253-301).  Here the heaps
 * live in HBM (device slot 0) and a block's result never travels to the host:
 *
 *   pqh_knn_heaps_create               n heaps of k items {uint32 index, float dist}
 *                                      (nn_item_t, fast_nn_temp_file.h:11-14); image = the bytes
 *                                      of an existing temp file (--no-init-temp-file) or NULL for
 *                                      fresh heaps {UINT_MAX, +inf} (init_temp_file, :83-129)
 *   pqh_knn_dataset_search_block_merge pqh_knn_dataset_search_block + run_merge_block on the
 *                                      device.  `ticket` numbers the blocks 0, 1, 2, ... in the
 *                                      order the reference's serial loop would merge them: the
 *                                      searches may run concurrently on different slots, the
 *                                      merges are applied in ticket order (a call waits for its
 *                                      turn; every ticket must be used exactly once).  Blocks of
 *                                      fewer than two rows only take their turn.
 *   pqh_knn_heaps_export               image (n*k items, may be NULL): the heap array as the temp
 *                                      file holds it; ids/dists (n*k each, may both be NULL): the
 *                                      heap-sorted rows as nn_indices.ivecsl / nn_dist.fvecsl
 *                                      hold them (unfilled slots 0xFFFFFFFF / +inf)
 *   pqh_knn_heaps_release              frees them
 */
int pqh_knn_heaps_create(long long n, int k, const void *image);
int pqh_knn_dataset_search_block_merge(int slot, int m, const int *row_ids, int k1,
                                       long long ticket);
int pqh_knn_heaps_export(void *image, unsigned *ids, float *dists);
void pqh_knn_heaps_release(void);

/*
 * The graph's first consumer: mst_builder turns the two output files into
 * N*k records {int32 source, int32 target, float dist} (mst_edge_t, mst.h:12-16) and sorts them
 * with libc qsort -- by distance before Kruskal's pass (mst.c:175 with mst_dist_comparator,
 * mst.c:15-26: distances closer than 1e-9 count as equal) and, for the forest's edges, by source
 * (mst.c:214, mst_source_comparator, mst.c:28-32).  glibc's qsort is a stable merge sort, so
 * where the comparator is an order the result is "ascending, equal keys in input order"; that is
 * what the device computes (least-significant-digit radix sort with order-keeping scatters,
 * edge_sort.cu).
 *
 *   pqh_sort_edges    sorts n such records in place (host memory).  PQH_SORT_BY_DIST is only
 *                     taken when the reference comparator IS an order on the data -- no NaN, and
 *                     no two different distances closer than 1e-9; otherwise nothing is changed
 *                     and PQH_ENOTORDER is returned (the caller keeps its own sort).
 *   pqh_qsort_edges   qsort's prototype, for mst.c compiled with -Dqsort=pqh_qsort_edges (no
 *                     source change): arrays of at least PQH_EDGE_SORT_MIN
 *                     12-byte records whose comparator answers a few probe records the way one
 *                     of mst.c's two does go to pqh_sort_edges; every other call -- other record
 *                     sizes, the per-row sorts of mst.c:134, PQH_ENOTORDER -- IS qsort.  A device
 *                     failure ends the process with exit(1) like knn_full_thread; there is no
 *                     silent CPU path for a sort the device was asked to do.
 */
#define PQH_SORT_BY_DIST 0
#define PQH_SORT_BY_SOURCE 1
#define PQH_EDGE_SORT_MIN 65536
int pqh_sort_edges(long long n, void *edges, int mode);
void pqh_qsort_edges(void *base, size_t n, size_t size,
                     int (*compar)(const void *, const void *));

/*
 * The estimator's joint code histograms (nn_huffman_estimator.c:69-107,
 * collect_stats_at_kth_nn), for the neighbour ranks [nn_first, nn_first + nn_count) at once:
 * stats[(j * m + i) * k_star * k_star + a * k_star + b] = number of vectors v whose neighbour of
 * rank nn_first + j is known (index != (unsigned)-1) and that have code a in sub-quantiser i
 * while that neighbour has code b; num_known[j] = how many vectors had a known neighbour of
 * that rank (the function's return value).  pq_codes is uint8[n x m] (pq_indices.bvecsl),
 * nn_indices uint32[n x num_nn] (nn_indices.ivecsl), k_star <= 256.  The counts are integers, so
 * the doubles are exactly the reference's whatever the order of accumulation.  The function is
 * static in the reference: the binding is the four-line change INTEGRATION.md shows.
 */
int pqh_nn_joint_histograms(long long n, int m, int k_star, int num_nn,
                            const unsigned char *pq_codes, const unsigned *nn_indices,
                            int nn_first, int nn_count, double *stats, long long *num_known);

/*
 * Counters of the last search on this thread's context, for tests and
 * bench.py: how the rows were resolved and how many kernels were launched.
 */
typedef struct pqh_knn_stats {
    long long rows;            /* query rows searched */
    long long rows_tensor;     /* rows resolved by the tensor-core pass + certificate */
    long long rows_exact;      /* rows (re)done by the exact FP32 kernel */
    long long rows_second_pass;/* rows whose first certificate failed and that were redone by a
                                  second tensor-core pass with longer candidate lists */
    long long launches;        /* kernels launched */
    double ms_total;           /* wall time of the call, host clock */
    double ms_h2d, ms_kernels, ms_d2h; /* device-event times of the phases */
    double ms_prep;            /* operand conversion + norms (device events) */
    double ms_gemm;            /* the tcgen05 distance-GEMM + filter kernel alone */
    double ms_rerank;          /* exact FP32 re-rank + certificate */
    double ms_exact;           /* exact FP32 kernel (fallback rows or PQH_ALGO_EXACT) */
    double candidates;         /* candidates re-ranked, summed over rows */
    double eps_ratio_max;      /* $PQH_TC_CHECK_EPS set: max |tensor score - fp64 score| / bound */
    long long inexact_operands;/* 1 if some FP16 operand was rounded (not integer-like data);
                                  2 if the data are integers in [0,255] and the exact INT8
                                  tensor-core path was taken ($PQH_TC_I8=0 disables it) */
} pqh_knn_stats_t;
int pqh_knn_last_stats(pqh_knn_stats_t *out);

/*
 * Test hook: the rows of the last host-pointer search (absolute query ids) that the first
 * tensor-core pass could not certify and that were therefore redone by the second pass or the
 * exact kernel.  Copies at most `capacity` ids to `rows` and returns how many there were (at
 * most 2^20 are remembered).  The full-size parity tests check every one of them against the
 * oracle.
 */
int pqh_knn_last_fallback_rows(int *rows, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* PQH_KNN_H_ */
