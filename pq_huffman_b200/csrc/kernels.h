// kernels.h -- host-callable launchers of the kNN kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

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
};
bool tc_supported(int n_rows, int nb, int d, int k, int self_join);
size_t tc_workspace_bytes(int n_rows, int nb, int d, int k, int self_join, int num_sms);
cudaError_t launch_knn_tc(const float *d_base, const float *d_queries, int q_row0, int q_begin,
                          int out_q_begin, int n_rows, int nb, int d, int k, int self_join,
                          int *d_assign, float *d_dis, void *d_work, size_t work_bytes,
                          int num_sms, cudaStream_t stream, int *launches, int **d_fb_rows_out,
                          unsigned *h_fb_count, cudaEvent_t *ev, TcTimes *times);

// ---- knn_tc3.cu (CTA-pair kernel; plan/maps/args are the structs of knn_tc_common.cuh) ----
int tc3_pick_nstage(const void *plan);
cudaError_t launch_tc3(const void *maps, const void *args, int gcap, int n_items, int num_sms,
                       cudaStream_t stream);

// ---- knn_tc5.cu (query operand in tensor memory) ----
int tc5_tile_rows();
int tc5_pick_nstage(const void *plan, int d16);
cudaError_t launch_tc5(const void *maps, const void *args, const void *xh_q, const void *fold_a,
                       int gcap, int d16, int n_items, int num_sms, cudaStream_t stream);

// ---- knn_tcs.cu (streaming kernel for large d) ----
cudaError_t launch_tcs(const void *maps, const void *args, int gcap, int d16, int n_items,
                       int num_sms, cudaStream_t stream);

}  // namespace pqh
