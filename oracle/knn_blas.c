/*
 * knn_blas.c -- Yael-shaped CPU baseline for the kNN hot path (timed arm only).
 *
 * TEST/BENCH INFRASTRUCTURE ONLY (see knn_oracle.c header).  This is what
 * `bench.py --impl reference` and bench.py's cpu_baseline leg time on the host
 * cores.  It is NOT the parity oracle: its distances come from the
 * |x|^2 + |y|^2 - 2 x.y expansion and differ from get_real_dist in the last
 * bits.
 *
 * Why a stand-in: the reference's CPU implementation of this path is Yael
 * v438's knn_full_thread (call site compute_nn_fast.c:478-481), which is not
 * vendored and not installed.  Yael's published algorithm, restated here:
 * queries are split across n_thread threads; each thread walks its queries in
 * blocks against base blocks, forms -2 x.y with BLAS sgemm, adds the
 * precomputed squared norms, and keeps one k-max-heap per query whose root is
 * replaced on a strictly smaller distance; heaps are sorted ascending at the
 * end.  BLAS here is the OpenBLAS bundled with scipy (symbol scipy_sgemm_),
 * opened with dlopen from the path in $PQH_OPENBLAS, single-threaded per call
 * so that the n_thread split is the only parallelism (as in Yael).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*sgemm_fn)(const char *, const char *, const int *, const int *, const int *,
                         const float *, const float *, const int *, const float *, const int *,
                         const float *, float *, const int *);
typedef void (*set_threads_fn)(int);

static sgemm_fn g_sgemm = NULL;

static void load_blas(void) {
    if (g_sgemm) return;
    const char *path = getenv("PQH_OPENBLAS");
    if (!path) {
        fprintf(stderr, "knn_blas: set PQH_OPENBLAS to scipy's libscipy_openblas*.so\n");
        exit(1);
    }
    void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        fprintf(stderr, "knn_blas: dlopen(%s): %s\n", path, dlerror());
        exit(1);
    }
    g_sgemm = (sgemm_fn)dlsym(h, "scipy_sgemm_");
    set_threads_fn st = (set_threads_fn)dlsym(h, "scipy_openblas_set_num_threads");
    if (!g_sgemm || !st) {
        fprintf(stderr, "knn_blas: scipy_sgemm_ / scipy_openblas_set_num_threads not found\n");
        exit(1);
    }
    st(1);
}

typedef struct {
    float dist;
    int id;
} item_t;

static void sift_down(item_t *heap, int n, item_t x) {
    int pos = 0;
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, c;
        if (l >= n) break;
        c = (r < n && heap[r].dist > heap[l].dist) ? r : l;
        if (!(heap[c].dist > x.dist)) break;
        heap[pos] = heap[c];
        pos = c;
    }
    heap[pos] = x;
}

typedef struct {
    int nb, d, k, self_join;
    const float *b, *v, *bnorm;
    int *assign;
    float *dis;
    int q_begin, q_end;
} job_t;

enum { QBLK = 256, BBLK = 2048 };

static void *worker(void *arg) {
    job_t *job = (job_t *)arg;
    const int d = job->d, k = job->k, nb = job->nb;
    float *c = (float *)malloc(sizeof(float) * QBLK * BBLK);
    item_t *heaps = (item_t *)malloc(sizeof(item_t) * (size_t)k * QBLK);
    float *qnorm = (float *)malloc(sizeof(float) * QBLK);
    const float alpha = -2.0f, beta = 0.0f;
    for (int q0 = job->q_begin; q0 < job->q_end; q0 += QBLK) {
        int nqb = job->q_end - q0 < QBLK ? job->q_end - q0 : QBLK;
        for (int i = 0; i < nqb; ++i) {
            const float *row = job->v + (size_t)(q0 + i) * d;
            float s = 0.0f;
            for (int t = 0; t < d; ++t) s += row[t] * row[t];
            qnorm[i] = s;
            for (int t = 0; t < k; ++t) {
                heaps[(size_t)i * k + t].dist = INFINITY;
                heaps[(size_t)i * k + t].id = -1;
            }
        }
        for (int b0 = 0; b0 < nb; b0 += BBLK) {
            int nbb = nb - b0 < BBLK ? nb - b0 : BBLK;
            /* c[q][j] = -2 * sum_i b[b0+j][i] * v[q0+q][i]  (column-major view: B^T * V) */
            g_sgemm("T", "N", &nbb, &nqb, &d, &alpha, job->b + (size_t)b0 * d, &d,
                    job->v + (size_t)q0 * d, &d, &beta, c, &nbb);
            for (int i = 0; i < nqb; ++i) {
                item_t *heap = heaps + (size_t)i * k;
                const float *crow = c + (size_t)i * nbb;
                const float qn = qnorm[i];
                const float *bn = job->bnorm + b0;
                float worst = heap[0].dist;
                for (int j = 0; j < nbb; ++j) {
                    float dist = qn + bn[j] + crow[j];
                    if (dist < worst) {
                        item_t it;
                        it.id = b0 + j;
                        if (job->self_join && it.id == q0 + i) dist = -INFINITY;
                        it.dist = dist;
                        sift_down(heap, k, it);
                        worst = heap[0].dist;
                    }
                }
            }
        }
        for (int i = 0; i < nqb; ++i) {
            item_t *heap = heaps + (size_t)i * k;
            int *out_i = job->assign + (size_t)(q0 + i) * k;
            float *out_d = job->dis + (size_t)(q0 + i) * k;
            int n = k;
            while (n > 0) {
                item_t top = heap[0];
                item_t last = heap[n - 1];
                --n;
                if (n > 0) sift_down(heap, n, last);
                out_i[n] = top.id;
                out_d[n] = top.dist == -INFINITY ? 0.0f : (top.dist < 0.0f ? 0.0f : top.dist);
            }
        }
    }
    free(c);
    free(heaps);
    free(qnorm);
    return NULL;
}

static void run(int nb, int d, int k, const float *b, const float *v, int q_begin, int q_end,
                int *assign, float *dis, int n_thread) {
    load_blas();
    if (q_begin >= q_end || k <= 0) return;
    float *bnorm = (float *)malloc(sizeof(float) * (size_t)nb);
    for (int j = 0; j < nb; ++j) {
        const float *row = b + (size_t)j * d;
        float s = 0.0f;
        for (int t = 0; t < d; ++t) s += row[t] * row[t];
        bnorm[j] = s;
    }
    int n = q_end - q_begin;
    if (n_thread < 1) n_thread = 1;
    if (n_thread > n) n_thread = n;
    pthread_t *threads = (pthread_t *)malloc(sizeof(pthread_t) * n_thread);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * n_thread);
    int per = (n + n_thread - 1) / n_thread;
    int started = 0;
    for (int t = 0; t < n_thread; ++t) {
        job_t *j = jobs + t;
        j->nb = nb; j->d = d; j->k = k; j->self_join = (b == v);
        j->b = b; j->v = v; j->bnorm = bnorm; j->assign = assign; j->dis = dis;
        j->q_begin = q_begin + t * per;
        j->q_end = q_begin + (t + 1) * per < q_end ? q_begin + (t + 1) * per : q_end;
        if (j->q_begin >= j->q_end) break;
        pthread_create(threads + t, NULL, worker, j);
        ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(threads[t], NULL);
    free(threads);
    free(jobs);
    free(bnorm);
}

/* Same prototype as Yael's <yael/nn.h> knn_full_thread. */
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *v, const float *b_weights, int *assign, float *dis,
                     int n_thread) {
    if (distance_type != 2 || b_weights != NULL) {
        fprintf(stderr, "knn_blas: only distance_type=2 (squared L2), b_weights=NULL\n");
        exit(1);
    }
    run(nb, d, k, b, v, 0, nq, assign, dis, n_thread);
}

/* bounded sample for bench.py: queries [q_begin,q_end) against all of b */
void pqh_blas_knn_rows(int nq_total, int nb, int d, int k, const float *b, const float *v,
                       int q_begin, int q_end, int *assign, float *dis, int n_thread) {
    if (q_end > nq_total) q_end = nq_total;
    run(nb, d, k, b, v, q_begin, q_end, assign, dis, n_thread);
}
