/*
 * pqh_compute_nn_fast -- host driver for the exact single-block mode of the reference's
 * compute_nn_fast CLI (compute_nn_fast.c:120-132 for the flags, :623-658 for the flow), with
 * the per-row stdio of the reference replaced by bulk I/O (the unmodified CLI
 * spends 4.5 s in fseek/fread/fwrite per row at 1M x 128 even with an instant kNN).
 *
 * It keeps the reference's FILE FORMATS and RESULT BYTES:
 *   in   <input>.fvecs        rows of [int32 d | d float32]            (compute_nn_fast.c:134-149)
 *   out  <tpl>nn_indices.ivecsl, <tpl>nn_dist.fvecsl
 *        [uint32 n | uint32 k | dense rows]                             (vecs_io.c:70-76)
 *   tmp  --temp-file, kept with --no-delete-temp-file: n x k records {uint32 id, float dist}
 *        holding each row's max-heap                                    (fast_nn_temp_file.h:11-14)
 * and reproduces the reference's merge for one block covering the whole data set: every row's
 * k-slot heap starts as {UINT_MAX, +inf}, receives result columns 1..num_nn_real in order
 * (column 0 is the vector itself, compute_nn_fast.c:498) under the rules of fast_nn_heap_push
 * (fast_nn_temp_file.c:48-54: a candidate is dropped unless strictly closer than the heap's
 * worst, or if its id is already present) and is exported by repeated extract-max
 * (fast_nn_temp_file.c:56-63).  tests/test_host_cli_gpu.py checks the output files byte for byte
 * against the reference's own binary.
 *
 * Only the exact mode (--num-dims 0: one block = the whole data set) is taken here; for the
 * approximate partitioned mode link the reference's own objects against libpqh_knn.so
 * (INTEGRATION.md).  The search itself is libpqh_knn.so (no CPU fallback): the .fvecs image is
 * handed over as is, with leading dimension d+1, and the library strips the per-row prefix in
 * its H2D copy.
 */
#define _GNU_SOURCE
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/pqh_knn.h"

typedef struct {
    uint32_t id;
    float dist;
} slot_t;

typedef struct {
    const char *input, *tpl, *temp_file;
    int num_nn, num_dims, delete_temp, with_blocks_stat, num_threads;
} options_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void usage(const char *argv0) {
    fprintf(stderr,
            "Usage: %s <input .fvecs> <output files template> <num_nn> [optional args]\n"
            "  accepts the flags of the reference's compute_nn_fast; only the exact mode\n"
            "  (--num-dims 0, one block) is implemented here\n",
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
    o.with_blocks_stat = 0;
    o.num_threads = 1;
    int ok = 1;
    for (int i = 4; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "--with-blocks-stat")) o.with_blocks_stat = 1;
        else if (!strcmp(a, "--no-delete-temp-file")) o.delete_temp = 0;
        else if (!strcmp(a, "--delete-temp-file")) o.delete_temp = 1;
        else if (!strcmp(a, "--no-init-temp-file") || !strcmp(a, "--init-temp-file")) { /* one block: always fresh */ }
        else if (!strcmp(a, "--temp-file") && i + 1 < argc) o.temp_file = argv[++i];
        else if (!strcmp(a, "--num-dims") && i + 1 < argc) o.num_dims = atoi(argv[++i]);
        else if (!strcmp(a, "--num-threads") && i + 1 < argc) o.num_threads = atoi(argv[++i]);
        else if ((!strcmp(a, "--blocks-info-cache") || !strcmp(a, "--block-overlap-fraction") ||
                  !strcmp(a, "--num-blocks-per-dim") || !strcmp(a, "--blocks-from") ||
                  !strcmp(a, "--blocks-to") || !strcmp(a, "--num-dimensions-at-pass")) &&
                 i + 1 < argc)
            ++i; /* partition knobs: meaningless for a single block */
        else {
            fprintf(stderr, "Unknown arg %s\n", a);
            ok = 0;
        }
    }
    if (!ok) exit(1);
    if (o.num_dims != 0) {
        fprintf(stderr,
                "%s: only the exact single-block mode (--num-dims 0) is implemented here; for the\n"
                "partitioned mode link the reference's compute_nn_fast against libpqh_knn.so\n",
                argv[0]);
        exit(1);
    }
    if (o.num_nn < 1) {
        fprintf(stderr, "num_nn must be >= 1\n");
        exit(1);
    }
    if (o.num_threads < 1) o.num_threads = 1;
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

typedef struct {
    long long row_begin, row_end;
    int k, k1;
    const int *assign;
    const float *dis;
    slot_t *heaps;   /* n x k, kept for the temp file (NULL: scratch per thread) */
    uint32_t *out_ids;
    float *out_dists;
} merge_job_t;

static void *merge_rows(void *arg) {
    merge_job_t *j = (merge_job_t *)arg;
    const int k = j->k, k1 = j->k1;
    slot_t *scratch = (slot_t *)malloc(sizeof(slot_t) * (size_t)k);
    for (long long r = j->row_begin; r < j->row_end; ++r) {
        slot_t *heap = j->heaps ? j->heaps + (size_t)r * k : scratch;
        for (int i = 0; i < k; ++i) {
            heap[i].id = UINT_MAX;
            heap[i].dist = INFINITY;
        }
        const int *ai = j->assign + (size_t)r * k1;
        const float *di = j->dis + (size_t)r * k1;
        for (int c = 1; c < k1; ++c) {
            slot_t x;
            x.id = (uint32_t)ai[c];
            x.dist = di[c];
            offer(heap, k, x);
        }
        slot_t *work = scratch;
        if (j->heaps) memcpy(work, heap, sizeof(slot_t) * (size_t)k); /* keep the heap image */
        else work = heap;
        drain_ascending(work, k, j->out_ids + (size_t)r * k, j->out_dists + (size_t)r * k);
    }
    free(scratch);
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

static void *init_device(void *arg) {
    (void)arg;
    if (pqh_knn_init(NULL, 0) != PQH_OK) {
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

    /* ---- exact kNN on the GPU: k1 = min(num_nn, n-1) + 1 columns, column 0 is the row itself
     *      (run_block, compute_nn_fast.c:469-483) ---- */
    const int k = o.num_nn;
    const int nn_real = (int)(n - 1 < k ? n - 1 : k);
    const int k1 = nn_real + 1;
    int *assign = (int *)malloc(sizeof(int) * (size_t)n * k1);
    float *dis = (float *)malloc(sizeof(float) * (size_t)n * k1);
    if (o.with_blocks_stat) {
        char *p = join(o.tpl, "blocks_stat.txt");
        FILE *bs = fopen(p, "w");
        if (bs) {
            fprintf(bs, "%lld ", n);
            fclose(bs);
        }
        free(p);
    }
    pthread_join(init_thread, NULL);
    const double t_init = now_s();
    if (n > 0) {
        int rc = pqh_knn_search_rows_ld((int)n, (int)n, d, k1, image + 1, d + 1, image + 1, d + 1,
                                        1, 0, (int)n, assign, dis);
        if (rc != PQH_OK) {
            fprintf(stderr, "pqh_knn_search_rows_ld failed: %s\n", pqh_last_error());
            return 1;
        }
    }
    const double t_knn = now_s();
    free(image);

    /* ---- merge + export, as the reference does through its temp file ---- */
    uint32_t *out_ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n * k);
    float *out_dists = (float *)malloc(sizeof(float) * (size_t)n * k);
    slot_t *heaps = o.delete_temp ? NULL : (slot_t *)malloc(sizeof(slot_t) * (size_t)n * k);
    int nt = o.num_threads;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < ncpu) nt = (int)(ncpu > 64 ? 64 : ncpu); /* host merge is ours: use the cores */
    if (nt > n) nt = n > 0 ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
    merge_job_t *jobs = (merge_job_t *)malloc(sizeof(merge_job_t) * nt);
    const long long per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        merge_job_t *j = jobs + t;
        j->row_begin = t * per;
        j->row_end = (t + 1) * per < n ? (t + 1) * per : n;
        j->k = k;
        j->k1 = k1;
        j->assign = assign;
        j->dis = dis;
        j->heaps = heaps;
        j->out_ids = out_ids;
        j->out_dists = out_dists;
        pthread_create(th + t, NULL, merge_rows, j);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    const double t_merge = now_s();

    char *pi = join(o.tpl, "nn_indices.ivecsl"), *pd = join(o.tpl, "nn_dist.fvecsl");
    write_light(pi, out_ids, sizeof(uint32_t), n, k);
    write_light(pd, out_dists, sizeof(float), n, k);
    if (heaps) {
        FILE *tf = fopen(o.temp_file, "wb");
        if (tf) {
            fwrite(heaps, sizeof(slot_t), (size_t)n * k, tf);
            fclose(tf);
        }
    }
    const double t_end = now_s();
    pqh_knn_stats_t s;
    memset(&s, 0, sizeof(s));
    pqh_knn_last_stats(&s);
    printf("done: read %.3f s (+%.3f s waiting for the CUDA context), knn %.3f s (h2d %.1f ms, kernels %.1f ms, d2h %.1f ms; %lld rows by "
           "tensor cores, %lld by the exact kernel), merge %.3f s, write %.3f s, total %.3f s\n",
           t_read - t0, t_init - t_read, t_knn - t_init, s.ms_h2d, s.ms_kernels, s.ms_d2h, s.rows_tensor,
           s.rows_exact, t_merge - t_knn, t_end - t_merge, t_end - t0);
    pqh_knn_shutdown();
    return 0;
}
