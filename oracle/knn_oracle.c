/*
 * knn_oracle.c -- CPU restatement of the kNN hot path of trukanduk/pq_huffman.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pq_huffman_b200/,
 * include/) may call, link or import this file.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py use it, and there only as the checker / the timed CPU arm.
 *
 * PARITY UNPINNED with respect to Yael: the arithmetic of the reference path is
 * one external call, knn_full_thread() from Yael v438 (#include <yael/nn.h> at
 * compute_nn_fast.c:11, -lyael at Makefile:5, version pinned by
 * /root/reference/.profile:9,11).  Yael is neither vendored in the reference
 * nor installed here, and the reference holds no test, fixture or golden vector
 * for the kNN numbers.  What CAN be pinned is pinned in
 * tests/test_oracle_pins.py against the reference's own compiled sources
 * (oracle/_ref/libref_cnf.so): get_real_dist (the distance definition) and
 * fast_nn_heap_push / fast_nn_heap_sort (the merge order).
 *
 * What this file follows:
 *   - distance definition: get_real_dist, compute_nn_fast.c:304-311 --
 *     sequential FP32 `tmp = a[i]-b[i]; result += tmp*tmp`, i = 0..d-1, with
 *     separate multiply and add roundings (the reference is built with
 *     gcc -std=c99, i.e. -ffp-contract=off; this file must be built the same);
 *   - call shape: knn_full_thread(2, nq, nb, d, k, b, v, NULL, assign, dis, nt)
 *     at compute_nn_fast.c:478-481; outputs are row-major [nq x k], ascending;
 *   - position-0 rule: the caller discards column 0 of every row as "self"
 *     (compute_nn_fast.c:498, compute_nn.py:22-24), so for a self-join
 *     (v == b) query i reports base row i first with distance 0;
 *   - ties: Yael's order among equal distances is unspecified; this oracle
 *     fixes it as ascending (distance, base position).
 *
 * Threading: queries are split over n_thread pthreads (Yael does the same with
 * its n_thread argument).  The inner loop processes LANES base rows at once
 * from a transposed tile so gcc can vectorise ACROSS pairs; each pair still
 * sees exactly the scalar sequence above, so results are bit-identical to a
 * plain scalar loop (checked in tests/test_oracle.py).
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LANES 16
#define BASE_TILE 1024 /* base rows per transposed tile (multiple of LANES) */

typedef struct {
    float dist;
    int id;
} item_t;

/* (dist, id) lexicographic "a is worse (larger) than b" */
static inline int item_gt(item_t a, item_t b) {
    return a.dist > b.dist || (a.dist == b.dist && a.id > b.id);
}

/* max-heap on (dist,id): heap[0] is the worst of the k kept so far */
static void heap_sift_down(item_t *heap, int n, int pos, item_t x) {
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, c;
        if (l >= n) break;
        c = (r < n && item_gt(heap[r], heap[l])) ? r : l;
        if (!item_gt(heap[c], x)) break;
        heap[pos] = heap[c];
        pos = c;
    }
    heap[pos] = x;
}

static void heap_sift_up(item_t *heap, int pos, item_t x) {
    while (pos > 0) {
        int p = (pos - 1) / 2;
        if (!item_gt(x, heap[p])) break;
        heap[pos] = heap[p];
        pos = p;
    }
    heap[pos] = x;
}

/* exported for tests: the scalar definition itself (compute_nn_fast.c:304-311) */
float pqh_oracle_dist(const float *a, const float *b, int nd) {
    float result = 0.0f;
    for (int i = 0; i < nd; ++i) {
        float tmp = a[i] - b[i];
        result += tmp * tmp;
    }
    return result;
}

typedef struct {
    int nq, nb, d, k;
    const float *b, *v;
    int *assign;
    float *dis;
    int self_join;
    int q_begin, q_end;
} job_t;

static void *worker(void *arg) {
    job_t *job = (job_t *)arg;
    const int d = job->d, k = job->k, nb = job->nb;
    /* a self-join row is [self, k-1 others]: the heap only holds the others */
    const int kcap = job->self_join ? k - 1 : k;
    float *tile = (float *)aligned_alloc(64, sizeof(float) * (size_t)d * BASE_TILE);
    item_t *heaps = NULL;
    int *heap_sizes = NULL;
    const int QB = 64; /* queries per pass over a base tile */
    heaps = (item_t *)malloc(sizeof(item_t) * (size_t)k * QB);
    heap_sizes = (int *)malloc(sizeof(int) * QB);
    if (!tile || !heaps || !heap_sizes) {
        fprintf(stderr, "knn_oracle: out of memory\n");
        exit(1);
    }

    for (int q0 = job->q_begin; q0 < job->q_end; q0 += QB) {
        int nqb = job->q_end - q0 < QB ? job->q_end - q0 : QB;
        for (int i = 0; i < nqb; ++i) heap_sizes[i] = 0;

        for (int b0 = 0; b0 < nb; b0 += BASE_TILE) {
            int nbt = nb - b0 < BASE_TILE ? nb - b0 : BASE_TILE;
            int nbt_pad = (nbt + LANES - 1) / LANES * LANES;
            /* transpose base tile: tile[i][j] = b[(b0+j)*d + i]; pad lanes with 0 */
            for (int j = 0; j < nbt; ++j) {
                const float *row = job->b + (size_t)(b0 + j) * d;
                for (int i = 0; i < d; ++i) tile[(size_t)i * BASE_TILE + j] = row[i];
            }
            for (int j = nbt; j < nbt_pad; ++j)
                for (int i = 0; i < d; ++i) tile[(size_t)i * BASE_TILE + j] = 0.0f;

            for (int qi = 0; qi < nqb; ++qi) {
                const int q = q0 + qi;
                const float *qrow = job->v + (size_t)q * d;
                item_t *heap = heaps + (size_t)qi * k;
                int hs = heap_sizes[qi];
                for (int j0 = 0; j0 < nbt_pad; j0 += LANES) {
                    float acc[LANES];
                    for (int l = 0; l < LANES; ++l) acc[l] = 0.0f;
                    for (int i = 0; i < d; ++i) {
                        const float qv = qrow[i];
                        const float *t = tile + (size_t)i * BASE_TILE + j0;
                        for (int l = 0; l < LANES; ++l) {
                            float tmp = qv - t[l];
                            acc[l] += tmp * tmp;
                        }
                    }
                    int lim = nbt - j0 < LANES ? nbt - j0 : LANES;
                    for (int l = 0; l < lim; ++l) {
                        item_t it;
                        it.id = b0 + j0 + l;
                        it.dist = acc[l];
                        if (job->self_join && it.id == q) continue; /* placed at column 0 below */
                        if (hs < kcap) {
                            heap_sift_up(heap, hs++, it);
                        } else if (kcap > 0 && item_gt(heap[0], it)) {
                            heap_sift_down(heap, hs, 0, it);
                        }
                    }
                }
                heap_sizes[qi] = hs;
            }
        }

        /* heap-sort each query's heap ascending and write the row */
        for (int qi = 0; qi < nqb; ++qi) {
            const int q = q0 + qi;
            item_t *heap = heaps + (size_t)qi * k;
            int hs = heap_sizes[qi];
            int *out_i = job->assign + (size_t)q * k;
            float *out_d = job->dis + (size_t)q * k;
            int first = 0;
            if (job->self_join) {
                out_i[0] = q;
                out_d[0] = 0.0f;
                first = 1;
            }
            int n = hs;
            while (n > 0) {
                item_t top = heap[0];
                item_t last = heap[n - 1];
                --n;
                if (n > 0) heap_sift_down(heap, n, 0, last);
                out_i[first + n] = top.id;
                out_d[first + n] = top.dist;
            }
            for (int c = first + hs; c < k; ++c) { /* k > nb: unfilled slots */
                out_i[c] = -1;
                out_d[c] = INFINITY;
            }
        }
    }
    free(tile);
    free(heaps);
    free(heap_sizes);
    return NULL;
}

/* Same prototype as Yael's <yael/nn.h> knn_full_thread. */
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *v, const float *b_weights, int *assign, float *dis,
                     int n_thread) {
    if (distance_type != 2 || b_weights != NULL) {
        fprintf(stderr, "knn_oracle: only distance_type=2 (squared L2), b_weights=NULL\n");
        exit(1);
    }
    if (nq <= 0 || k <= 0) return;
    if (n_thread < 1) n_thread = 1;
    if (n_thread > nq) n_thread = nq;
    pthread_t *threads = (pthread_t *)malloc(sizeof(pthread_t) * n_thread);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * n_thread);
    int per = (nq + n_thread - 1) / n_thread;
    per = (per + 63) / 64 * 64;
    int started = 0;
    for (int t = 0; t < n_thread; ++t) {
        job_t *j = jobs + t;
        j->nq = nq; j->nb = nb; j->d = d; j->k = k;
        j->b = b; j->v = v; j->assign = assign; j->dis = dis;
        j->self_join = (b == v);
        j->q_begin = t * per;
        j->q_end = (t + 1) * per < nq ? (t + 1) * per : nq;
        if (j->q_begin >= j->q_end) break;
        pthread_create(threads + t, NULL, worker, j);
        ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(threads[t], NULL);
    free(threads);
    free(jobs);
}

/* Sliced entry used by bench.py's bounded CPU sample and by the world_size>1
 * tests: queries [q_begin, q_end) of v against all of b; outputs are written
 * at row q (same layout as the full call). */
void pqh_oracle_knn_rows(int nq_total, int nb, int d, int k, const float *b, const float *v,
                         int q_begin, int q_end, int *assign, float *dis, int n_thread) {
    if (q_end > nq_total) q_end = nq_total;
    if (q_begin >= q_end || k <= 0) return;
    int n = q_end - q_begin;
    if (n_thread < 1) n_thread = 1;
    if (n_thread > n) n_thread = n;
    pthread_t *threads = (pthread_t *)malloc(sizeof(pthread_t) * n_thread);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * n_thread);
    int per = (n + n_thread - 1) / n_thread;
    int started = 0;
    for (int t = 0; t < n_thread; ++t) {
        job_t *j = jobs + t;
        j->nq = nq_total; j->nb = nb; j->d = d; j->k = k;
        j->b = b; j->v = v; j->assign = assign; j->dis = dis;
        j->self_join = (b == v);
        j->q_begin = q_begin + t * per;
        j->q_end = q_begin + (t + 1) * per < q_end ? q_begin + (t + 1) * per : q_end;
        if (j->q_begin >= j->q_end) break;
        pthread_create(threads + t, NULL, worker, j);
        ++started;
    }
    for (int t = 0; t < started; ++t) pthread_join(threads[t], NULL);
    free(threads);
    free(jobs);
}
