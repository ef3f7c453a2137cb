/*
 * pqh_shim.c -- TEST INFRASTRUCTURE, never shipped: the subset of include/pqh_knn.h that the
 * native driver (pq_huffman_b200/host/pqh_compute_nn_fast.c) calls, implemented on the CPU with
 * the scalar oracle (knn_oracle.c), so that the driver's HOST logic -- partition, block order,
 * heap merge, file formats, temp-file resume -- can be held against the reference's binary by
 * the CPU test suite (tests/test_host_cli_cpu.py).  Linked only into
 * oracle/lib/pqh_compute_nn_fast_shim; the product binary links libpqh_knn.so and has no CPU path.
 */
#include <stdlib.h>
#include <string.h>

#include "../include/pqh_knn.h"

/* knn_oracle.c */
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *v, const float *b_weights, int *assign, float *dis, int n_thread);

static float *g_rows;
static int g_n, g_d;

const char *pqh_last_error(void) { return "cpu shim"; }
int pqh_knn_init(const int *devices, int n_devices) {
    (void)devices;
    (void)n_devices;
    return PQH_OK;
}
void pqh_knn_shutdown(void) { pqh_knn_dataset_release(); }
int pqh_knn_device_count(void) { return 2; } /* exercise the multi-device ring of the driver */
int pqh_knn_last_stats(pqh_knn_stats_t *out) {
    memset(out, 0, sizeof(*out));
    return PQH_OK;
}
void *pqh_knn_host_alloc(long long bytes) { return malloc((size_t)bytes); }
void pqh_knn_host_free(void *p) { free(p); }

static float *densify(const float *rows, int n, int d, long long ld) {
    float *x = (float *)malloc(sizeof(float) * (size_t)(n ? n : 1) * d);
    for (int i = 0; i < n; ++i) memcpy(x + (size_t)i * d, rows + (size_t)i * ld, sizeof(float) * d);
    return x;
}

int pqh_knn_dataset_load(int n, int d, const float *rows, long long ld) {
    free(g_rows);
    g_rows = densify(rows, n, d, ld);
    g_n = n;
    g_d = d;
    return PQH_OK;
}
void pqh_knn_dataset_release(void) {
    free(g_rows);
    g_rows = NULL;
}
int pqh_knn_dataset_search_block(int slot, int m, const int *row_ids, int k, int *assign,
                                 float *dis) {
    (void)slot;
    float *blk = (float *)malloc(sizeof(float) * (size_t)(m ? m : 1) * g_d);
    for (int i = 0; i < m; ++i)
        memcpy(blk + (size_t)i * g_d, g_rows + (size_t)row_ids[i] * g_d, sizeof(float) * g_d);
    knn_full_thread(2, m, m, g_d, k, blk, blk, NULL, assign, dis, 2);
    free(blk);
    return PQH_OK;
}
int pqh_knn_search_rows_ld(int nq_total, int nb, int d, int k, const float *b, long long ldb,
                           const float *v, long long ldv, int self_join, int q_begin, int q_end,
                           int *assign, float *dis) {
    (void)v;
    (void)ldv;
    if (!self_join || q_begin != 0 || q_end != nq_total || nq_total != nb) return PQH_EINVAL;
    float *x = densify(b, nb, d, ldb);
    knn_full_thread(2, nb, nb, d, k, x, x, NULL, assign, dis, 4);
    free(x);
    return PQH_OK;
}
