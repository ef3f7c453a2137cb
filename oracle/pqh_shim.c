/*
 * pqh_shim.c -- TEST INFRASTRUCTURE, never shipped: the subset of include/pqh_knn.h that the
 * native driver (pq_huffman_b200/host/pqh_compute_nn_fast.c) calls, implemented on the CPU with
 * the scalar oracle (knn_oracle.c), so that the driver's HOST logic -- partition, block order,
 * heap merge, file formats, temp-file resume -- can be held against the reference's binary by
 * the CPU test suite (tests/test_host_cli_cpu.py).  Linked only into
 * oracle/lib/pqh_compute_nn_fast_shim; the product binary links libpqh_knn.so and has no CPU path.
 */
#include <math.h>
#include <pthread.h>
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

/* ---- device-side merge API, on the CPU: the reference's heap rules restated once more
 * (fast_nn_temp_file.c:11-63, compute_nn_fast.c:485-508) so that the driver's ticket-ordered,
 * result-stays-on-the-device flow can be held against the reference's binary without a GPU ---- */
typedef struct {
    unsigned id;
    float dist;
} item_t;
static item_t *g_heaps;
static long long g_hn, g_ticket;
static int g_hk;
static pthread_mutex_t g_hmu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_hcv = PTHREAD_COND_INITIALIZER;

static void sink(item_t *heap, int n, item_t x) {
    int hole = 0;
    for (;;) {
        int left = 2 * hole + 1, right = left + 1, pick = -1;
        if (right < n && heap[right].dist > heap[left].dist && heap[right].dist > x.dist) pick = right;
        else if (left < n && heap[left].dist > x.dist) pick = left;
        if (pick < 0) break;
        heap[hole] = heap[pick];
        hole = pick;
    }
    heap[hole] = x;
}

int pqh_knn_heaps_create(long long n, int k, const void *image) {
    free(g_heaps);
    g_heaps = (item_t *)malloc(sizeof(item_t) * (size_t)(n * k + 1));
    g_hn = n;
    g_hk = k;
    g_ticket = 0;
    if (image) {
        memcpy(g_heaps, image, sizeof(item_t) * (size_t)(n * k));
    } else {
        for (long long i = 0; i < n * k; ++i) {
            g_heaps[i].id = 0xFFFFFFFFu;
            g_heaps[i].dist = INFINITY;
        }
    }
    return PQH_OK;
}

int pqh_knn_dataset_search_block_merge(int slot, int m, const int *row_ids, int k1,
                                       long long ticket) {
    int *assign = NULL;
    float *dis = NULL;
    if (m > 1 && k1 > 1) {
        assign = (int *)malloc(sizeof(int) * (size_t)m * k1);
        dis = (float *)malloc(sizeof(float) * (size_t)m * k1);
        pqh_knn_dataset_search_block(slot, m, row_ids, k1, assign, dis);
    }
    pthread_mutex_lock(&g_hmu);
    while (g_ticket != ticket) pthread_cond_wait(&g_hcv, &g_hmu);
    if (assign) {
        for (int r = 0; r < m; ++r) {
            item_t *heap = g_heaps + (size_t)row_ids[r] * g_hk;
            for (int c = 1; c < k1; ++c) {
                const int a = assign[(size_t)r * k1 + c];
                if (a < 0) continue;
                item_t x = {(unsigned)row_ids[a], dis[(size_t)r * k1 + c]};
                int present = 0;
                if (x.dist >= heap[0].dist) continue;
                for (int i = 0; i < g_hk; ++i) present |= heap[i].id == x.id;
                if (!present) sink(heap, g_hk, x);
            }
        }
    }
    ++g_ticket;
    pthread_cond_broadcast(&g_hcv);
    pthread_mutex_unlock(&g_hmu);
    free(assign);
    free(dis);
    return PQH_OK;
}

int pqh_knn_heaps_export(void *image, unsigned *ids, float *dists) {
    if (image) memcpy(image, g_heaps, sizeof(item_t) * (size_t)(g_hn * g_hk));
    if (ids && dists) {
        item_t *w = (item_t *)malloc(sizeof(item_t) * (size_t)(g_hk + 1));
        for (long long r = 0; r < g_hn; ++r) {
            memcpy(w, g_heaps + (size_t)r * g_hk, sizeof(item_t) * (size_t)g_hk);
            for (int n = g_hk - 1; n >= 0; --n) {
                ids[(size_t)r * g_hk + n] = w[0].id;
                dists[(size_t)r * g_hk + n] = w[0].dist;
                sink(w, n, w[n]);
            }
        }
        free(w);
    }
    return PQH_OK;
}

void pqh_knn_heaps_release(void) {
    free(g_heaps);
    g_heaps = NULL;
}
