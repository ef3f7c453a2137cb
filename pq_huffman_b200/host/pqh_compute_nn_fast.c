/*
 * pqh_compute_nn_fast -- host driver with the command line of the reference's compute_nn_fast
 * (compute_nn_fast.c:120-132 for the flags, :623-658 for the flow), with the per-row stdio of
 * the reference replaced by bulk I/O (the unmodified CLI spends 4.5 s in
 * fseek/fread/fwrite per row at 1M x 128 even with an instant kNN) and, in the partitioned mode,
 * its one-pass-over-the-file-per-block loader (fast_nn_block_loader2.c:86-146) replaced by a
 * data set resident in HBM whose blocks are named by row ids (pqh_knn_dataset_*).
 *
 * It keeps the reference's FILE FORMATS and RESULT BYTES:
 *   in   <input>.fvecs        rows of [int32 d | d float32]            (compute_nn_fast.c:134-149)
 *   out  <tpl>nn_indices.ivecsl, <tpl>nn_dist.fvecsl
 *        [uint32 n | uint32 k | dense rows]                             (vecs_io.c:70-76)
 *   tmp  --temp-file, kept with --no-delete-temp-file: n x k records {uint32 id, float dist}
 *        holding each row's max-heap                                    (fast_nn_temp_file.h:11-14)
 * and reproduces the reference's merge: every row's k-slot heap starts as {UINT_MAX, +inf}
 * (or as the temp file left it, --no-init-temp-file); blocks are visited in ascending id, the
 * rows of a block in file order, and each row receives result columns 1..num_nn_real in order
 * (column 0 is the vector itself, compute_nn_fast.c:498) under the rules of fast_nn_heap_push
 * (fast_nn_temp_file.c:48-54: a candidate is dropped unless strictly closer than the heap's
 * worst, or if its id is already present); heaps are exported by repeated extract-max
 * (fast_nn_temp_file.c:56-63).  tests/test_host_cli_gpu.py checks the output files byte for byte
 * against the reference's own binary in both modes.
 *
 * Modes: --num-dims 0 is the exact search (one block = the whole data set, handed to
 * pqh_knn_search_rows_ld as the .fvecs image itself, leading dimension d+1); --num-dims S > 0
 * (the reference's default is 5) is its approximate mode over blocks_per_dim^S overlapping boxes
 * (pqh_partition.h).  There the blocks are searched by one host thread per device of
 * $PQH_DEVICES and merged ON THE DEVICE, in block order, into heaps resident in HBM
 * (pqh_knn_heaps_*: a block's result never travels to the host; --host-merge keeps the heaps
 * here and merges block b on the host cores while the GPU searches block b+1 -- the checker
 * for the device kernels).  --blocks-from / --blocks-to select a range
 * with the reference's arithmetic (compute_nn_fast.c:545-555, which leaves an EMPTY range for any
 * --blocks-to inside the range); --num-dimensions-at-pass only groups blocks and cannot change
 * the result.  The search itself is libpqh_knn.so -- no CPU fallback.
 */
#define _GNU_SOURCE
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/pqh_knn.h"
#include "pqh_partition.h"

typedef struct {
    uint32_t id;
    float dist;
} slot_t;

typedef struct {
    const char *input, *tpl, *temp_file;
    const char *blocks_info_cache;
    int num_nn, num_dims, delete_temp, init_temp, with_blocks_stat, num_threads, host_merge;
    int blocks_per_dim, dims_on_pass;
    double overlap;
    long long blocks_from, blocks_to;
} options_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void usage(const char *argv0) {
    fprintf(stderr,
            "Usage: %s <input .fvecs> <output files template> <num_nn> [optional args]\n"
            "  [--num-dims S] [--num-blocks-per-dim B] [--block-overlap-fraction F]\n"
            "  [--blocks-info-cache FILE] [--temp-file FILE] [--no-init-temp-file]\n"
            "  [--no-delete-temp-file] [--blocks-from I] [--blocks-to J] [--with-blocks-stat]\n"
            "  [--num-threads T] [--num-dimensions-at-pass P]   (the reference's flags)\n"
            "  [--host-merge]   merge the blocks on the host cores instead of on the device\n",
            argv0);
}

static char *join(const char *a, const char *b) {
    size_t la = strlen(a), lb = strlen(b);
    char *r = (char *)malloc(la + lb + 1);
    memcpy(r, a, la);
    memcpy(r + la, b, lb + 1);
    return r;
}

static options_t parse(int argc, char **argv) {
    options_t o;
    if (argc < 4) {
        fprintf(stderr, "Too few required args\n");
        usage(argv[0]);
        exit(1);
    }
    o.input = argv[1];
    o.tpl = argv[2];
    o.num_nn = atoi(argv[3]);
    o.temp_file = "nn_temp.vecs";
    o.num_dims = 5; /* the reference's default is the partitioned mode */
    o.delete_temp = 1;
    o.init_temp = 1;
    o.host_merge = 0;
    o.with_blocks_stat = 0;
    o.num_threads = 0; /* 0: flag absent -> every core this process may run on */
    o.blocks_info_cache = NULL;
    o.blocks_per_dim = 3;
    o.overlap = 0.3;
    o.dims_on_pass = 0;
    o.blocks_from = 0;
    o.blocks_to = -1;
    int ok = 1;
    for (int i = 4; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "--with-blocks-stat")) o.with_blocks_stat = 1;
        else if (!strcmp(a, "--no-delete-temp-file")) o.delete_temp = 0;
        else if (!strcmp(a, "--delete-temp-file")) o.delete_temp = 1;
        else if (!strcmp(a, "--no-init-temp-file")) o.init_temp = 0;
        else if (!strcmp(a, "--init-temp-file")) o.init_temp = 1;
        else if (!strcmp(a, "--temp-file") && i + 1 < argc) o.temp_file = argv[++i];
        else if (!strcmp(a, "--num-dims") && i + 1 < argc) o.num_dims = atoi(argv[++i]);
        else if (!strcmp(a, "--num-threads") && i + 1 < argc) o.num_threads = atoi(argv[++i]);
        else if (!strcmp(a, "--host-merge")) o.host_merge = 1;
        else if (!strcmp(a, "--blocks-info-cache") && i + 1 < argc) o.blocks_info_cache = argv[++i];
        else if (!strcmp(a, "--block-overlap-fraction") && i + 1 < argc) o.overlap = atof(argv[++i]);
        else if (!strcmp(a, "--num-blocks-per-dim") && i + 1 < argc) o.blocks_per_dim = atoi(argv[++i]);
        else if (!strcmp(a, "--blocks-from") && i + 1 < argc) o.blocks_from = atoll(argv[++i]);
        else if (!strcmp(a, "--blocks-to") && i + 1 < argc) o.blocks_to = atoll(argv[++i]) - 1;
        else if (!strcmp(a, "--num-dimensions-at-pass") && i + 1 < argc) o.dims_on_pass = atoi(argv[++i]);
        else {
            fprintf(stderr, "Unknown arg %s\n", a);
            ok = 0;
        }
    }
    if (!ok) exit(1);
    if (o.num_dims < 0 || o.num_dims > 60 || o.blocks_per_dim < 1) {
        fprintf(stderr, "bad --num-dims / --num-blocks-per-dim\n");
        exit(1);
    }
    if (o.num_nn < 1) {
        fprintf(stderr, "num_nn must be >= 1\n");
        exit(1);
    }
    if (o.num_threads < 0) o.num_threads = 0;
    return o;
}

/* ---- the reference's k-slot max-heap, restated (fast_nn_temp_file.c:23-63) ---- */

/* Put `x` into the hole at the root of heap[0..n) and let it sink: at each level the larger
 * child moves up when it is farther than x; the right child is only preferred when strictly
 * farther than the left one. */
static void sink_from_root(slot_t *heap, int n, slot_t x) {
    int hole = 0;
    for (;;) {
        int left = 2 * hole + 1, right = left + 1, pick = -1;
        if (right < n && heap[right].dist > heap[left].dist && heap[right].dist > x.dist)
            pick = right;
        else if (left < n && heap[left].dist > x.dist)
            pick = left;
        if (pick < 0) break;
        heap[hole] = heap[pick];
        hole = pick;
    }
    heap[hole] = x;
}

static void offer(slot_t *heap, int k, slot_t x) {
    if (x.dist >= heap[0].dist) return; /* ties lose */
    for (int i = 0; i < k; ++i)
        if (heap[i].id == x.id) return; /* already present */
    sink_from_root(heap, k, x);
}

/* extract-max k times: slot n-1 receives the current worst, the former last element re-enters
 * at the root of the shortened heap */
static void drain_ascending(slot_t *heap, int k, uint32_t *ids, float *dists) {
    for (int n = k - 1; n >= 0; --n) {
        ids[n] = heap[0].id;
        dists[n] = heap[0].dist;
        sink_from_root(heap, n, heap[n]);
    }
}

/* ---- fork-join helper: fn(t, nt, arg) on nt threads, the caller being thread 0 ---- */
typedef void (*par_fn)(int t, int nt, void *arg);
typedef struct {
    par_fn fn;
    int t, nt;
    void *arg;
} par_job_t;

static void *par_entry(void *p) {
    par_job_t *j = (par_job_t *)p;
    j->fn(j->t, j->nt, j->arg);
    return NULL;
}

static void parallel(int nt, par_fn fn, void *arg) {
    if (nt <= 1) {
        fn(0, 1, arg);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    par_job_t *jobs = (par_job_t *)malloc(sizeof(par_job_t) * (size_t)nt);
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn;
        jobs[t].t = t;
        jobs[t].nt = nt;
        jobs[t].arg = arg;
        if (t) pthread_create(th + t, NULL, par_entry, jobs + t);
    }
    fn(0, nt, arg);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

/* ---- the same with threads that stay: for the stage that forks once per block.  (Creating
 * and joining 15 threads took longer than finding a block's rows: 1.4-1.9 ms per block against
 * the ~0.3 ms the search itself needs -- as long as one block's search on the GPU.)  A pool
 * belongs to ONE calling thread. ---- */
typedef struct pool_s pool_t;
typedef struct {
    pool_t *pool;
    int t;
} pool_worker_t;
struct pool_s {
    int nt; /* threads including the caller */
    pthread_t *th;
    pool_worker_t *workers;
    pthread_mutex_t mu;
    pthread_cond_t cv_start, cv_done;
    par_fn fn;
    void *arg;
    unsigned long generation;
    int pending, stop;
};

static void *pool_entry(void *p) {
    pool_worker_t *w = (pool_worker_t *)p;
    pool_t *pl = w->pool;
    unsigned long seen = 0;
    for (;;) {
        pthread_mutex_lock(&pl->mu);
        while (pl->generation == seen && !pl->stop) pthread_cond_wait(&pl->cv_start, &pl->mu);
        if (pl->stop) {
            pthread_mutex_unlock(&pl->mu);
            return NULL;
        }
        seen = pl->generation;
        par_fn fn = pl->fn;
        void *arg = pl->arg;
        pthread_mutex_unlock(&pl->mu);
        fn(w->t, pl->nt, arg);
        pthread_mutex_lock(&pl->mu);
        if (--pl->pending == 0) pthread_cond_signal(&pl->cv_done);
        pthread_mutex_unlock(&pl->mu);
    }
}

static void pool_init(pool_t *pl, int nt) {
    memset(pl, 0, sizeof(*pl));
    pl->nt = nt < 1 ? 1 : nt;
    pthread_mutex_init(&pl->mu, NULL);
    pthread_cond_init(&pl->cv_start, NULL);
    pthread_cond_init(&pl->cv_done, NULL);
    pl->th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)pl->nt);
    pl->workers = (pool_worker_t *)malloc(sizeof(pool_worker_t) * (size_t)pl->nt);
    for (int t = 1; t < pl->nt; ++t) {
        pl->workers[t].pool = pl;
        pl->workers[t].t = t;
        pthread_create(pl->th + t, NULL, pool_entry, pl->workers + t);
    }
}

static void pool_run(pool_t *pl, par_fn fn, void *arg) {
    if (pl->nt > 1) {
        pthread_mutex_lock(&pl->mu);
        pl->fn = fn;
        pl->arg = arg;
        pl->pending = pl->nt - 1;
        ++pl->generation;
        pthread_cond_broadcast(&pl->cv_start);
        pthread_mutex_unlock(&pl->mu);
    }
    fn(0, pl->nt, arg);
    if (pl->nt > 1) {
        pthread_mutex_lock(&pl->mu);
        while (pl->pending > 0) pthread_cond_wait(&pl->cv_done, &pl->mu);
        pthread_mutex_unlock(&pl->mu);
    }
}

static void pool_destroy(pool_t *pl) {
    pthread_mutex_lock(&pl->mu);
    pl->stop = 1;
    pthread_cond_broadcast(&pl->cv_start);
    pthread_mutex_unlock(&pl->mu);
    for (int t = 1; t < pl->nt; ++t) pthread_join(pl->th[t], NULL);
    free(pl->th);
    free(pl->workers);
    pthread_cond_destroy(&pl->cv_start);
    pthread_cond_destroy(&pl->cv_done);
    pthread_mutex_destroy(&pl->mu);
}

/* ---- heaps of all rows ---- */
typedef struct {
    slot_t *heaps;
    long long n;
    int k;
    uint32_t *out_ids;
    float *out_dists;
} heaps_arg_t;

static void clear_heaps_part(int t, int nt, void *arg) {
    heaps_arg_t *a = (heaps_arg_t *)arg;
    const long long total = a->n * a->k;
    for (long long i = total * t / nt; i < total * (t + 1) / nt; ++i) {
        a->heaps[i].id = UINT_MAX;
        a->heaps[i].dist = INFINITY;
    }
}

static void export_heaps_part(int t, int nt, void *arg) {
    heaps_arg_t *a = (heaps_arg_t *)arg;
    const int k = a->k;
    slot_t *work = (slot_t *)malloc(sizeof(slot_t) * (size_t)k);
    for (long long r = a->n * t / nt; r < a->n * (t + 1) / nt; ++r) {
        memcpy(work, a->heaps + (size_t)r * k, sizeof(slot_t) * (size_t)k); /* keep the heap image */
        drain_ascending(work, k, a->out_ids + (size_t)r * k, a->out_dists + (size_t)r * k);
    }
    free(work);
}

/* One searched block into the heaps (run_merge_block, compute_nn_fast.c:485-508).  Rows of a
 * block are distinct, so row ranges can be merged concurrently. */
typedef struct {
    slot_t *heaps;
    int k, k1;
    long long m;
    const int *row_ids; /* block position -> row of the file; NULL: the identity */
    const int *assign;
    const float *dis;
} merge_arg_t;

static void merge_block_part(int t, int nt, void *arg) {
    merge_arg_t *a = (merge_arg_t *)arg;
    const int k = a->k, k1 = a->k1;
    for (long long r = a->m * t / nt; r < a->m * (t + 1) / nt; ++r) {
        slot_t *heap = a->heaps + (size_t)(a->row_ids ? a->row_ids[r] : r) * k;
        const int *ai = a->assign + (size_t)r * k1;
        const float *di = a->dis + (size_t)r * k1;
        for (int c = 1; c < k1; ++c) {
            slot_t x;
            if (ai[c] < 0) continue; /* unfilled tail of a row (k > block size): no neighbour */
            x.id = (uint32_t)(a->row_ids ? a->row_ids[ai[c]] : ai[c]);
            x.dist = di[c];
            offer(heap, k, x);
        }
    }
}

static void merge_block(int nt, slot_t *heaps, int k, int k1, long long m, const int *row_ids,
                        const int *assign, const float *dis) {
    merge_arg_t a = {heaps, k, k1, m, row_ids, assign, dis};
    if (m < 4096) nt = 1;
    parallel(nt, merge_block_part, &a);
}

/* ---- rows of one block, in file order, found by all cores ---- */
typedef struct {
    const pqh_partition_t *part;
    const float *cols;
    long long n, block, prefix_block;
    int prefix_len;
    int *ids;            /* thread t writes at ids + its first row: never overlaps */
    long long *counts;
} rows_arg_t;

static void block_rows_part(int t, int nt, void *arg) {
    rows_arg_t *a = (rows_arg_t *)arg;
    const long long r0 = a->n * t / nt, r1 = a->n * (t + 1) / nt;
    a->counts[t] = pqh_partition_block_rows(a->part, a->cols, r0, r1, a->block, a->prefix_block,
                                            a->prefix_len, a->ids + r0);
}

static long long block_rows(pool_t *pool, const pqh_partition_t *part, const float *cols,
                            long long n, long long block, long long prefix_block, int prefix_len,
                            int *ids, long long *counts) {
    rows_arg_t a = {part, cols, n, block, prefix_block, prefix_len, ids, counts};
    const int nt = pool->nt;
    pool_run(pool, block_rows_part, &a);
    long long m = counts[0];
    for (int t = 1; t < nt; ++t) {
        memmove(ids + m, ids + n * t / nt, sizeof(int) * (size_t)counts[t]);
        m += counts[t];
    }
    return m;
}

/* ---- the GPU side of the partitioned mode: one host thread per device ---- */
enum { SLOT_FREE = 0, SLOT_READY = 1, SLOT_SEARCHED = 2 };
typedef struct {
    int *ids;
    long long m;
    int k1;
    int *assign; /* page-locked, grown on demand by the device thread */
    float *dis;
    long long cap;
    int state;
} ring_slot_t;

typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t cv;
    ring_slot_t *ring;
    int ring_len, num_devices;
    long long begin, end;
    /* what the row-finding stage needs */
    const pqh_partition_t *part;
    const float *cols;
    long long n, blocks_on_pass;
    int prefix_len, k, threads;
    FILE *blocks_stat;
    double t_rows;
    int device_merge; /* blocks are merged into heaps resident on the device, in block order */
} pipeline_t;

typedef struct {
    pipeline_t *pl;
    int device_slot;
} device_thread_t;

static void *device_thread(void *arg) {
    device_thread_t *dt = (device_thread_t *)arg;
    pipeline_t *pl = dt->pl;
    for (long long b = pl->begin + dt->device_slot; b < pl->end; b += pl->num_devices) {
        ring_slot_t *s = pl->ring + (b - pl->begin) % pl->ring_len;
        pthread_mutex_lock(&pl->mu);
        while (s->state != SLOT_READY) pthread_cond_wait(&pl->cv, &pl->mu);
        pthread_mutex_unlock(&pl->mu);
        if (pl->device_merge) {
            /* search + run_merge_block on the device; the ticket keeps the merges in block order
             * across devices (blocks of one row only take their turn) */
            int rc = pqh_knn_dataset_search_block_merge(dt->device_slot, (int)s->m, s->ids, s->k1,
                                                        b - pl->begin);
            if (rc != PQH_OK) {
                fprintf(stderr, "block %lld: pqh_knn_dataset_search_block_merge failed: %s\n", b,
                        pqh_last_error());
                exit(1);
            }
        } else {
            const long long need = s->m * s->k1;
            if (need > s->cap) {
                pqh_knn_host_free(s->assign);
                pqh_knn_host_free(s->dis);
                s->cap = need + need / 4 + 1024;
                s->assign = (int *)pqh_knn_host_alloc(s->cap * (long long)sizeof(int));
                s->dis = (float *)pqh_knn_host_alloc(s->cap * (long long)sizeof(float));
                if (!s->assign || !s->dis) {
                    fprintf(stderr, "cannot allocate %lld result slots: %s\n", s->cap,
                            pqh_last_error());
                    exit(1);
                }
            }
            if (s->m > 1) { /* a block of one row has no neighbours to offer */
                int rc = pqh_knn_dataset_search_block(dt->device_slot, (int)s->m, s->ids, s->k1,
                                                      s->assign, s->dis);
                if (rc != PQH_OK) {
                    fprintf(stderr, "block %lld: pqh_knn_dataset_search_block failed: %s\n", b,
                            pqh_last_error());
                    exit(1);
                }
            }
        }
        pthread_mutex_lock(&pl->mu);
        s->state = SLOT_SEARCHED;
        pthread_cond_broadcast(&pl->cv);
        pthread_mutex_unlock(&pl->mu);
    }
    return NULL;
}

/* First stage: the rows of each block, in block order, as ring slots come free (a slot is
 * released by the merge of the block ring_len places earlier). */
static void *rows_thread(void *arg) {
    pipeline_t *pl = (pipeline_t *)arg;
    long long *counts = (long long *)malloc(sizeof(long long) * (size_t)pl->threads);
    pool_t pool;
    pool_init(&pool, pl->threads);
    for (long long b = pl->begin; b < pl->end; ++b) {
        ring_slot_t *s = pl->ring + (b - pl->begin) % pl->ring_len;
        pthread_mutex_lock(&pl->mu);
        while (s->state != SLOT_FREE) pthread_cond_wait(&pl->cv, &pl->mu);
        pthread_mutex_unlock(&pl->mu);
        const double tr = now_s();
        const long long pass_first =
            pl->begin + (b - pl->begin) / pl->blocks_on_pass * pl->blocks_on_pass;
        s->m = block_rows(&pool, pl->part, pl->cols, pl->n, b, pass_first, pl->prefix_len, s->ids,
                          counts);
        s->k1 = (int)(s->m - 1 < pl->k ? s->m - 1 : pl->k) + 1;
        if (pl->blocks_stat) fprintf(pl->blocks_stat, "%d ", (int)s->m);
        pl->t_rows += now_s() - tr;
        pthread_mutex_lock(&pl->mu);
        s->state = SLOT_READY;
        pthread_cond_broadcast(&pl->cv);
        pthread_mutex_unlock(&pl->mu);
    }
    pool_destroy(&pool);
    free(counts);
    return NULL;
}

static void write_light(const char *path, const void *data, size_t elem, long long n, int k) {
    FILE *f = fopen(path, "wb");
    if (!f) {
        fprintf(stderr, "Cannot open %s for writing\n", path);
        exit(1);
    }
    uint32_t head[2] = {(uint32_t)n, (uint32_t)k};
    if (fwrite(head, sizeof(uint32_t), 2, f) != 2 ||
        fwrite(data, elem, (size_t)n * k, f) != (size_t)n * k) {
        fprintf(stderr, "Short write to %s\n", path);
        exit(1);
    }
    fclose(f);
}

static double g_init_seconds = 0;

static void *init_device(void *arg) {
    (void)arg;
    const double t = now_s();
    int rc = pqh_knn_init(NULL, 0);
    g_init_seconds = now_s() - t;
    if (rc != PQH_OK) {
        fprintf(stderr, "pqh_knn_init failed: %s\n", pqh_last_error());
        exit(1);
    }
    return NULL;
}

int main(int argc, char **argv) {
    setvbuf(stdout, NULL, _IONBF, 0);
    options_t o = parse(argc, argv);
    const double t0 = now_s();
    /* This process only ever uses the devices named in $PQH_DEVICES (default: device 0); hiding
     * the others from CUDA (unless the user already chose) cuts runtime initialisation on an
     * 8-GPU box from ~1 s to a fraction of that.  Ordinals are then relative to the visible set. */
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        const char *devs = getenv("PQH_DEVICES");
        setenv("CUDA_VISIBLE_DEVICES", (devs && *devs) ? devs : "0", 0);
        if (devs && *devs) {
            /* remap PQH_DEVICES to 0..n-1 of the visible set */
            int n = 1;
            for (const char *c = devs; *c; ++c) n += (*c == ',');
            char buf[256];
            size_t off = 0;
            for (int i = 0; i < n && off + 8 < sizeof(buf); ++i)
                off += (size_t)snprintf(buf + off, sizeof(buf) - off, i ? ",%d" : "%d", i);
            setenv("PQH_DEVICES", buf, 1);
        }
    }
    /* CUDA context creation takes longer than reading the file: do both at once */
    pthread_t init_thread;
    pthread_create(&init_thread, NULL, init_device, NULL);

    /* ---- input: the whole .fvecs image in one read ---- */
    FILE *f = fopen(o.input, "rb");
    if (!f) {
        fprintf(stderr, "Cannot open %s file\n", o.input);
        return 1;
    }
    int32_t d = 0;
    if (fread(&d, sizeof(d), 1, f) != 1 || d <= 0) {
        fprintf(stderr, "%s: cannot read the dimension prefix\n", o.input);
        return 1;
    }
    struct stat st;
    fstat(fileno(f), &st);
    const long long row_bytes = 4LL * (d + 1);
    const long long n = st.st_size / row_bytes;
    printf("num vectors %lld * row_size %lld (num_dim %d) == filesize %lld\n", n, row_bytes, d,
           (long long)st.st_size);
    if (n * row_bytes != st.st_size || n > INT_MAX) {
        fprintf(stderr, "%s: size is not a whole number of %lld-byte rows\n", o.input, row_bytes);
        return 1;
    }
    float *image = (float *)malloc((size_t)st.st_size);
    if (!image) {
        fprintf(stderr, "out of memory for the input image\n");
        return 1;
    }
    rewind(f);
    if (fread(image, 1, (size_t)st.st_size, f) != (size_t)st.st_size) {
        fprintf(stderr, "%s: short read\n", o.input);
        return 1;
    }
    fclose(f);
    const double t_read = now_s();

    const int k = o.num_nn;
    /* host threads (row finding, merges): an explicit --num-threads is honoured, as the
     * reference hands it to knn_full_thread as the thread budget (compute_nn_fast.c:481);
     * without the flag every core of the affinity mask is used (cpusets count, not the box) */
    int nt = o.num_threads;
    if (nt <= 0) {
        cpu_set_t set;
        long ncpu = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set)
                                                                : sysconf(_SC_NPROCESSORS_ONLN);
        nt = (int)(ncpu > 64 ? 64 : (ncpu < 1 ? 1 : ncpu));
    }
    if (nt > n) nt = n > 0 ? (int)n : 1;

    /* ---- the heaps: fresh, or as an earlier run left them in the temp file ---- */
    slot_t *heaps = (slot_t *)malloc(sizeof(slot_t) * (size_t)n * k + 1);
    uint32_t *out_ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n * k + 1);
    float *out_dists = (float *)malloc(sizeof(float) * (size_t)n * k + 1);
    if (!heaps || !out_ids || !out_dists) {
        fprintf(stderr, "out of memory for %lld x %d results\n", n, k);
        return 1;
    }
    heaps_arg_t ha = {heaps, n, k, out_ids, out_dists};
    /* partitioned mode: the heaps live on the device and the blocks are merged there
     * (pqh_knn_heaps_*); --host-merge keeps them here, merged by the host cores */
    const int device_merge = !(n == 0 || o.num_dims == 0) && !o.host_merge;
    if (o.init_temp) {
        if (!device_merge) parallel(nt, clear_heaps_part, &ha);
    } else {
        FILE *tf = fopen(o.temp_file, "rb");
        if (!tf || fread(heaps, sizeof(slot_t), (size_t)n * k, tf) != (size_t)n * k) {
            fprintf(stderr, "--no-init-temp-file: cannot read %lld x %d records from %s\n", n, k,
                    o.temp_file);
            return 1;
        }
        fclose(tf);
    }

    /* ---- the blocks ---- */
    pqh_partition_t part;
    float *cols = NULL;
    if (n == 0 || o.num_dims == 0) {
        memset(&part, 0, sizeof(part));
        part.num_blocks = 1;
        part.blocks_per_dim = o.blocks_per_dim;
    } else {
        if (o.num_dims > d) {
            fprintf(stderr, "--num-dims %d exceeds the %d dimensions of the input\n", o.num_dims, d);
            return 1;
        }
        if (o.blocks_info_cache && pqh_partition_load(&part, o.blocks_info_cache) == 0) {
            printf("Blocks info: Loaded\n");
            if (part.num_splits > d) {
                fprintf(stderr, "%s does not describe this input\n", o.blocks_info_cache);
                return 1;
            }
        } else {
            if (pqh_partition_build(&part, image + 1, n, d + 1, o.num_dims, o.blocks_per_dim,
                                    o.overlap, nt) != 0) {
                fprintf(stderr, "cannot build the blocks info\n");
                return 1;
            }
            if (o.blocks_info_cache) pqh_partition_save(&part, o.blocks_info_cache);
            printf("Blocks info: Done\n");
        }
        cols = pqh_partition_columns(&part, image + 1, n, d + 1);
    }
    /* the block range, with the reference's arithmetic (compute_nn_fast.c:545-555) */
    long long begin = o.blocks_from, end = o.blocks_to;
    if (begin > part.num_blocks) begin = part.num_blocks;
    if (begin < 0) begin = 0;
    if (end < 0 || end > part.num_blocks) end = part.num_blocks;
    else if (end > o.blocks_from) end = o.blocks_from;
    long long blocks_on_pass = 1;
    for (int i = 0; i < o.dims_on_pass; ++i) blocks_on_pass *= part.blocks_per_dim;
    const int prefix_len = blocks_on_pass > 1 ? part.num_splits - o.dims_on_pass : 0;
    const double t_part = now_s();

    FILE *bs = NULL;
    if (o.with_blocks_stat) {
        char *p = join(o.tpl, "blocks_stat.txt");
        bs = fopen(p, "w");
        free(p);
    }
    pthread_join(init_thread, NULL);
    const double t_init = now_s();

    double t_wait_gpu = 0, t_merge_only = 0, t_rows = 0;
    long long rows_in_blocks = 0;
    if (n > 0 && part.num_splits == 0) {
        /* one block = the whole file (run_block, compute_nn_fast.c:469-483): k1 =
         * min(num_nn, n-1) + 1 columns, column 0 is the row itself */
        if (begin < end) {
            const int k1 = (int)(n - 1 < k ? n - 1 : k) + 1;
            int *assign = (int *)malloc(sizeof(int) * (size_t)n * k1);
            float *dis = (float *)malloc(sizeof(float) * (size_t)n * k1);
            if (bs) fprintf(bs, "%lld ", n);
            int rc = pqh_knn_search_rows_ld((int)n, (int)n, d, k1, image + 1, d + 1, image + 1,
                                            d + 1, 1, 0, (int)n, assign, dis);
            if (rc != PQH_OK) {
                fprintf(stderr, "pqh_knn_search_rows_ld failed: %s\n", pqh_last_error());
                return 1;
            }
            t_wait_gpu = now_s() - t_init;
            rows_in_blocks = n;
            const double tm = now_s();
            merge_block(nt, heaps, k, k1, n, NULL, assign, dis);
            t_merge_only = now_s() - tm;
            free(assign);
            free(dis);
        }
    } else if (n > 0 && begin < end) {
        if (pqh_knn_dataset_load((int)n, d, image + 1, d + 1) != PQH_OK) {
            fprintf(stderr, "pqh_knn_dataset_load failed: %s\n", pqh_last_error());
            return 1;
        }
        if (device_merge && pqh_knn_heaps_create(n, k, o.init_temp ? NULL : heaps) != PQH_OK) {
            fprintf(stderr, "pqh_knn_heaps_create failed: %s\n", pqh_last_error());
            return 1;
        }
        pipeline_t pl;
        pl.device_merge = device_merge;
        pthread_mutex_init(&pl.mu, NULL);
        pthread_cond_init(&pl.cv, NULL);
        pl.num_devices = pqh_knn_device_count();
        if (pl.num_devices < 1) pl.num_devices = 1;
        pl.ring_len = 2 * pl.num_devices + 2;
        pl.begin = begin;
        pl.end = end;
        pl.part = &part;
        pl.cols = cols;
        pl.n = n;
        pl.blocks_on_pass = blocks_on_pass;
        pl.prefix_len = prefix_len;
        pl.k = k;
        pl.threads = nt > 2 ? nt / 2 : 1; /* the other half of the cores merges meanwhile */
        pl.blocks_stat = bs;
        pl.t_rows = 0;
        pl.ring = (ring_slot_t *)calloc((size_t)pl.ring_len, sizeof(ring_slot_t));
        for (int i = 0; i < pl.ring_len; ++i) pl.ring[i].ids = (int *)malloc(sizeof(int) * (size_t)n);
        pthread_t rows_th;
        pthread_t *dth = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)pl.num_devices);
        device_thread_t *dts = (device_thread_t *)malloc(sizeof(device_thread_t) * (size_t)pl.num_devices);
        pthread_create(&rows_th, NULL, rows_thread, &pl);
        for (int g = 0; g < pl.num_devices; ++g) {
            dts[g].pl = &pl;
            dts[g].device_slot = g;
            pthread_create(dth + g, NULL, device_thread, dts + g);
        }
        /* three stages run concurrently: rows of block b+.. (rows_thread), search of the blocks
         * in between (one thread per device), merge of block b (here, in block order) */
        const int merge_threads = nt > 2 ? nt - pl.threads : 1;
        for (long long b = begin; b < end; ++b) {
            if ((b - begin) % 10 == 0)
                printf("Start block %6lld / %6lld (%6lld / %6lld)\n", b - begin, end - begin, b,
                       part.num_blocks);
            ring_slot_t *s = pl.ring + (b - begin) % pl.ring_len;
            const double tw = now_s();
            pthread_mutex_lock(&pl.mu);
            while (s->state != SLOT_SEARCHED) pthread_cond_wait(&pl.cv, &pl.mu);
            pthread_mutex_unlock(&pl.mu);
            const double tm = now_s();
            t_wait_gpu += tm - tw;
            if (s->m > 1 && !device_merge)
                merge_block(merge_threads, heaps, k, s->k1, s->m, s->ids, s->assign, s->dis);
            t_merge_only += now_s() - tm;
            rows_in_blocks += s->m;
            pthread_mutex_lock(&pl.mu);
            s->state = SLOT_FREE;
            pthread_cond_broadcast(&pl.cv);
            pthread_mutex_unlock(&pl.mu);
        }
        pthread_join(rows_th, NULL);
        t_rows = pl.t_rows;
        for (int g = 0; g < pl.num_devices; ++g) pthread_join(dth[g], NULL);
        for (int i = 0; i < pl.ring_len; ++i) {
            free(pl.ring[i].ids);
            pqh_knn_host_free(pl.ring[i].assign);
            pqh_knn_host_free(pl.ring[i].dis);
        }
        free(pl.ring);
        free(dth);
        free(dts);
    }
    if (bs) fclose(bs);
    const double t_knn = now_s();
    free(image);
    free(cols);

    /* ---- export, as the reference does from its temp file (temp_file_to_result) ---- */
    if (device_merge && n > 0 && begin < end) {
        /* heap-sorted rows (and the heap image, if the temp file is kept) from the device */
        if (pqh_knn_heaps_export(o.delete_temp ? NULL : heaps, out_ids, out_dists) != PQH_OK) {
            fprintf(stderr, "pqh_knn_heaps_export failed: %s\n", pqh_last_error());
            return 1;
        }
        pqh_knn_heaps_release();
    } else {
        if (device_merge && o.init_temp) parallel(nt, clear_heaps_part, &ha); /* no block ran */
        parallel(nt, export_heaps_part, &ha);
    }
    char *pi = join(o.tpl, "nn_indices.ivecsl"), *pd = join(o.tpl, "nn_dist.fvecsl");
    write_light(pi, out_ids, sizeof(uint32_t), n, k);
    write_light(pd, out_dists, sizeof(float), n, k);
    if (o.delete_temp) {
        remove(o.temp_file);
    } else {
        FILE *tf = fopen(o.temp_file, "wb");
        if (!tf || fwrite(heaps, sizeof(slot_t), (size_t)n * k, tf) != (size_t)n * k) {
            fprintf(stderr, "cannot write the temp file %s\n", o.temp_file);
            return 1;
        }
        fclose(tf);
    }
    const double t_end = now_s();
    pqh_knn_stats_t s;
    memset(&s, 0, sizeof(s));
    pqh_knn_last_stats(&s);
    printf("done: %lld blocks, %lld rows in blocks; read %.3f s, blocks info %.3f s, CUDA context "
           "%.3f s (+%.3f s not hidden), search+merge %.3f s (block rows %.3f s, waiting for the GPU %.3f s, merge "
           "%.3f s; device: h2d %.1f ms, kernels %.1f ms, d2h %.1f ms; %lld rows by tensor cores, "
           "%lld by the exact kernel), export+write %.3f s, total %.3f s\n",
           end > begin ? end - begin : 0LL, rows_in_blocks, t_read - t0, t_part - t_read,
           g_init_seconds, t_init - t_part, t_knn - t_init, t_rows, t_wait_gpu, t_merge_only, s.ms_h2d, s.ms_kernels,
           s.ms_d2h, s.rows_tensor, s.rows_exact, t_end - t_knn, t_end - t0);
    pqh_knn_shutdown();
    printf("context released in %.3f s\n", now_s() - t_end);
    return 0;
}
