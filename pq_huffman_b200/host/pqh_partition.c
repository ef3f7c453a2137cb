/*
 * pqh_partition.c -- see pqh_partition.h.  Restates, on an in-memory image of the input, what the
 * reference derives by re-reading the file once per split coordinate
 * (fast_nn_blocks_info.c:21-39): the boxes are quantile ranges of each coordinate, widened by
 * overlap/2 of the data set on both sides, the outermost ones opened by 1.0.
 */
#define _GNU_SOURCE
#include "pqh_partition.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* Ordering the reference sorts a coordinate with (fast_nn_blocks_info.c:41-50): values closer
 * than 1e-9 compare equal.  The difference is taken in float, as there.  Using libc qsort with
 * this ordering keeps the quantile VALUES identical to the reference's on the same libc even
 * where that tolerance makes the order implementation-defined. */
static int order_with_slack(const void *a, const void *b) {
    const float diff = *(const float *)a - *(const float *)b;
    if (fabs(diff) < 1e-9) return 0;
    return diff < 0 ? -1 : 1;
}

static long long clampll(long long x, long long lo, long long hi) {
    return x < lo ? lo : (x > hi ? hi : x);
}

typedef struct {
    pqh_split_t *split;
    const float *rows;
    long long n, ld;
    int blocks_per_dim;
    double overlap;
    int rc;
} split_job_t;

static void *build_split(void *arg) {
    split_job_t *j = (split_job_t *)arg;
    pqh_split_t *sp = j->split;
    const long long n = j->n;
    const int nb = j->blocks_per_dim;
    float *col = (float *)malloc(sizeof(float) * (size_t)n);
    sp->lo = (float *)malloc(sizeof(float) * 2 * (size_t)nb);
    if (!col || !sp->lo) {
        free(col);
        j->rc = -1;
        return NULL;
    }
    sp->hi = sp->lo + nb;
    for (long long i = 0; i < n; ++i) col[i] = j->rows[i * j->ld + sp->dimension];
    qsort(col, (size_t)n, sizeof(float), order_with_slack);
    const long long widen = (long long)(n * j->overlap / 2);
    for (int i = 0; i < nb; ++i) {
        const long long first = clampll(n * i / nb - widen, 0, n);
        const long long last = clampll(n * (i + 1) / nb + widen, 0, n);
        sp->lo[i] = col[first < n ? first : n - 1];
        sp->hi[i] = col[last < n - 1 ? last : n - 1];
    }
    /* the outermost boxes are open-ended (double arithmetic, rounded on the store) */
    sp->lo[0] = col[0] - 1.0;
    sp->hi[nb - 1] = col[n - 1] + 1.0;
    free(col);
    j->rc = 0;
    return NULL;
}

int pqh_partition_build(pqh_partition_t *p, const float *rows, long long n, long long ld,
                        int num_splits, int blocks_per_dim, double overlap, int threads) {
    memset(p, 0, sizeof(*p));
    if (num_splits < 0 || blocks_per_dim < 1 || n < 1 || !rows || num_splits > ld) return -1;
    p->num_splits = num_splits;
    p->overlap = overlap;
    p->blocks_per_dim = blocks_per_dim;
    p->num_blocks = 1;
    p->splits = (pqh_split_t *)calloc((size_t)(num_splits ? num_splits : 1), sizeof(pqh_split_t));
    split_job_t *jobs = (split_job_t *)calloc((size_t)(num_splits ? num_splits : 1), sizeof(*jobs));
    pthread_t *th = (pthread_t *)calloc((size_t)(num_splits ? num_splits : 1), sizeof(*th));
    if (!p->splits || !jobs || !th) return -1;
    for (int s = 0; s < num_splits; ++s) {
        pqh_split_t *sp = p->splits + s;
        sp->dimension = num_splits - 1 - s;
        sp->digit_weight = 1;
        for (int i = 0; i < sp->dimension; ++i) sp->digit_weight *= blocks_per_dim;
        p->num_blocks *= blocks_per_dim;
        jobs[s].split = sp;
        jobs[s].rows = rows;
        jobs[s].n = n;
        jobs[s].ld = ld;
        jobs[s].blocks_per_dim = blocks_per_dim;
        jobs[s].overlap = overlap;
    }
    int rc = 0;
    if (threads < 1) threads = 1;
    for (int s0 = 0; s0 < num_splits; s0 += threads) {
        const int s1 = s0 + threads < num_splits ? s0 + threads : num_splits;
        for (int s = s0; s < s1; ++s) pthread_create(th + s, NULL, build_split, jobs + s);
        for (int s = s0; s < s1; ++s) {
            pthread_join(th[s], NULL);
            if (jobs[s].rc) rc = -1;
        }
    }
    free(jobs);
    free(th);
    return rc;
}

void pqh_partition_free(pqh_partition_t *p) {
    if (!p || !p->splits) return;
    for (int s = 0; s < p->num_splits; ++s) free(p->splits[s].lo);
    free(p->splits);
    memset(p, 0, sizeof(*p));
}

/* Cache file: int num_splits | double overlap | int blocks_per_dim | long long num_blocks, then
 * per split: int dimension | long long digit_weight | float lo[nb] | float hi[nb]. */
int pqh_partition_save(const pqh_partition_t *p, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) {
        fprintf(stderr, "Cannot open file: %s\n", path);
        return -1;
    }
    fwrite(&p->num_splits, sizeof(int), 1, f);
    fwrite(&p->overlap, sizeof(double), 1, f);
    fwrite(&p->blocks_per_dim, sizeof(int), 1, f);
    fwrite(&p->num_blocks, sizeof(long long), 1, f);
    for (int s = 0; s < p->num_splits; ++s) {
        fwrite(&p->splits[s].dimension, sizeof(int), 1, f);
        fwrite(&p->splits[s].digit_weight, sizeof(long long), 1, f);
        fwrite(p->splits[s].lo, sizeof(float), 2 * (size_t)p->blocks_per_dim, f);
    }
    fclose(f);
    return 0;
}

int pqh_partition_load(pqh_partition_t *p, const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "No dimensions info file: %s\n", path);
        return -1;
    }
    memset(p, 0, sizeof(*p));
    int ok = fread(&p->num_splits, sizeof(int), 1, f) == 1 &&
             fread(&p->overlap, sizeof(double), 1, f) == 1 &&
             fread(&p->blocks_per_dim, sizeof(int), 1, f) == 1 &&
             fread(&p->num_blocks, sizeof(long long), 1, f) == 1 && p->num_splits >= 0 &&
             p->num_splits < 64 && p->blocks_per_dim >= 1;
    if (ok) {
        p->splits = (pqh_split_t *)calloc((size_t)(p->num_splits ? p->num_splits : 1),
                                          sizeof(pqh_split_t));
        ok = p->splits != NULL;
    }
    for (int s = 0; ok && s < p->num_splits; ++s) {
        pqh_split_t *sp = p->splits + s;
        const size_t nb = (size_t)p->blocks_per_dim;
        sp->lo = (float *)malloc(sizeof(float) * 2 * nb);
        sp->hi = sp->lo + nb;
        ok = sp->lo && fread(&sp->dimension, sizeof(int), 1, f) == 1 &&
             fread(&sp->digit_weight, sizeof(long long), 1, f) == 1 &&
             fread(sp->lo, sizeof(float), 2 * nb, f) == 2 * nb;
    }
    fclose(f);
    if (!ok) {
        fprintf(stderr, "%s: not a blocks-info cache file\n", path);
        pqh_partition_free(p);
        return -2;
    }
    return 0;
}

float *pqh_partition_columns(const pqh_partition_t *p, const float *rows, long long n,
                             long long ld) {
    const int ns = p->num_splits;
    float *cols = (float *)malloc(sizeof(float) * (size_t)n * (size_t)(ns ? ns : 1));
    if (!cols) return NULL;
    for (long long i = 0; i < n; ++i)
        for (int s = 0; s < ns; ++s) cols[i * ns + s] = rows[i * ld + p->splits[s].dimension];
    return cols;
}

int pqh_partition_contains(const pqh_partition_t *p, const float *c, long long block_id,
                           int s_begin, int s_end) {
    for (int s = s_begin; s < s_end; ++s) {
        const pqh_split_t *sp = p->splits + s;
        const long long box = block_id / sp->digit_weight % p->blocks_per_dim;
        if (c[s] < sp->lo[box] || c[s] > sp->hi[box]) return 0;
    }
    return 1;
}

long long pqh_partition_block_rows(const pqh_partition_t *p, const float *cols, long long row_begin,
                                   long long row_end, long long block_id, long long prefix_block,
                                   int prefix_len, int *out) {
    const int ns = p->num_splits;
    /* the boxes of this block, looked up once */
    float lo[64], hi[64], plo[64], phi[64];
    if (ns > 64) return -1;
    for (int s = 0; s < ns; ++s) {
        const pqh_split_t *sp = p->splits + s;
        const long long box = block_id / sp->digit_weight % p->blocks_per_dim;
        lo[s] = sp->lo[box];
        hi[s] = sp->hi[box];
        const long long pbox = prefix_block / sp->digit_weight % p->blocks_per_dim;
        plo[s] = sp->lo[pbox];
        phi[s] = sp->hi[pbox];
    }
    long long cnt = 0;
    for (long long i = row_begin; i < row_end; ++i) {
        const float *c = cols + i * ns;
        int in = 1;
        for (int s = 0; s < prefix_len && in; ++s) in = !(c[s] < plo[s] || c[s] > phi[s]);
        for (int s = 0; s < ns && in; ++s) in = !(c[s] < lo[s] || c[s] > hi[s]);
        if (in) out[cnt++] = (int)i;
    }
    return cnt;
}
