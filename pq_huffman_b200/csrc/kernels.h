// kernels.h -- host-callable launchers of the kNN kernels (internal).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stddef.h>

#include <vector>

namespace pqh {

// ---- knn_exact.cu ----------------------------------------------------------
size_t exact_smem_bytes();
int exact_pick_nsplit(int n_rows, int nb, int num_sms);
size_t exact_workspace_bytes(int n_rows, int kcap, int nsplit);
size_t exact_workspace_bound(int n_rows_max, int kcap, int num_sms);
cudaError_t launch_knn_exact(const float *d_base, const float *d_queries, const int *d_row_ids,
                             int q_begin, int q_row0, int out_q_begin, int n_rows, int nb, int d,
                             int k,
                             int self_join, int *d_assign, float *d_dis, void *d_work,
                             size_t work_bytes, int num_sms, cudaStream_t stream, int *launches);

// ---- knn_tc.cu -------------------------------------------------------------
struct TcTimes {
    float ms_prep = 0.f, ms_gemm = 0.f, ms_rerank = 0.f;
    double candidates = 0.0;
    int inexact = 0;
    long long rows_pass2 = 0;   // rows redone by the second tensor-core pass
    float eps_ratio_max = 0.f;  // PQH_TC_CHECK_EPS: max |S'_tensor - S'_fp64| / eps
    int slices = 0;             // row slices the first pass was cut into
    // absolute ids of the rows the first pass did not certify (redone by the second pass or
    // the exact kernel); at most 2^20 are kept -- a test hook (pqh_knn_last_fallback_rows)
    std::vector<int> pass1_failed;
};
// The PQH_TC_* environment knobs are read once, when the device context is created
// (pqh_knn_init / first search); tc_reload_env() is that read.
void tc_reload_env();
bool tc_supported(int n_rows, int nb, int d, int k, int self_join);
size_t tc_workspace_bytes(int n_rows, int nb, int d, int k, int self_join, int num_sms);
cudaError_t launch_knn_tc(const float *d_base, const float *d_queries, int q_row0, int q_begin,
                          int out_q_begin, int n_rows, int nb, int d, int k, int self_join,
                          int *d_assign, float *d_dis, void *d_work, size_t work_bytes,
                          int num_sms, cudaStream_t stream, int *launches, int **d_fb_rows_out,
                          unsigned *h_fb_count, TcTimes *times);

// ---- knn_order.cu (locality order of the sweep) ----
struct OrderOut {
    const int *base_perm;  // operand row p of the base = data row base_perm[p]
    const int *base_inv;   // inverse
    const int *q_perm;     // query row r of the slice = absolute query id q_perm[r]
    const int *q_self;     // self-join: position of query r's own vector in the permuted base
    const int *qt_start;   // first base tile of the sweep of each query tile
};
struct OrderBufs {
    int C;  // cells (= pivots); 0: no locality order
    int *piv_ids, *cell_b, *base_perm, *base_inv, *cell_q, *q_perm, *q_self;
    int *piv_raw, *piv_group, *group_top;  // pivot numbering: evenly spaced rows, their groups
    int *hist_b, *start_b, *cur_b, *hist_q, *start_q, *cur_q, *qt_start;
};
int order_num_cells(int nb);
size_t order_workspace_bytes(int slice_rows, int nb, int slice_qtiles);
void order_carve(void *d_work, int slice_rows, int nb, int slice_qtiles, OrderBufs *w);
cudaError_t launch_order_pivots(const OrderBufs &w, const float *d_base, int nb, int d,
                                cudaStream_t stream, int *launches);
cudaError_t launch_order_pick(const void *cand, const int *cnt, int n_rows, int gcap, int nlists,
                              int tail_row0, int ns_tail, int *cell, int *hist,
                              cudaStream_t stream, int *launches);
cudaError_t launch_order_sort_base(const OrderBufs &w, int nb, cudaStream_t stream, int *launches);
cudaError_t launch_order_clear_queries(const OrderBufs &w, cudaStream_t stream);
cudaError_t launch_order_sort_queries(const OrderBufs &w, int q_begin, int n_rows, int self_join,
                                      int n_qtiles, int own_tiles, cudaStream_t stream,
                                      int *launches, OrderOut *out);
cudaError_t launch_order_self_of(const int *d_ids, int n, const int *base_inv, int *d_q_self,
                                 cudaStream_t stream, int *launches);

// ---- edge_sort.cu (stable sort of the graph's edge records for the MST builder) ----
size_t edge_sort_workspace_bytes(size_t n);
cudaError_t launch_edge_sort(const void *d_in, void *d_out, size_t n, int mode, void *d_work,
                             uint32_t *d_flags, cudaStream_t stream, int *launches);

// ---- nn_stats.cu (joint code histograms of the graph for the estimator) ----
cudaError_t launch_nn_histograms(const void *d_pq, const void *d_nn, long long n, int m, int k_star,
                                 int num_nn, int nn_first, int nn_count, uint32_t *d_cnt,
                                 unsigned long long *d_known, uint32_t *d_flags, double *d_out,
                                 cudaStream_t stream, int *launches);

// ---- knn_tcs.cu (streaming kernel for large d) ----
cudaError_t launch_tcs(const void *maps, const void *args, int gcap, int d16, int n_items,
                       int num_sms, cudaStream_t stream);

}  // namespace pqh
