/*
 * pqh_partition.h -- geometry of compute_nn_fast's approximate mode: the data set is cut along
 * its first `num_splits` coordinates into blocks_per_dim^num_splits OVERLAPPING boxes ("blocks"),
 * each of which is then searched on its own (fast_nn_blocks_info.h:8-22 for the description kept
 * per split, fast_nn_blocks_info.c:52-85 for the quantile rule, :199-227 for membership).
 *
 * This is host logic of the native driver (pqh_compute_nn_fast) and has no CUDA in it; it is also
 * built as libpqh_host.so so that tests can hold it against the reference's own objects.
 */
#ifndef PQH_PARTITION_H_
#define PQH_PARTITION_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pqh_split {
    int dimension;          /* coordinate this split cuts */
    long long digit_weight; /* blocks_per_dim^dimension: block_id / weight % blocks_per_dim = box */
    float *lo, *hi;         /* [blocks_per_dim] closed interval of each box; hi = lo + blocks_per_dim */
} pqh_split_t;

typedef struct pqh_partition {
    int num_splits;
    double overlap;         /* fraction of the data set two neighbouring boxes share */
    int blocks_per_dim;
    long long num_blocks;   /* blocks_per_dim^num_splits (1 when num_splits == 0) */
    pqh_split_t *splits;    /* split s cuts coordinate num_splits-1-s */
} pqh_partition_t;

/* rows: n vectors, vector i at rows + i*ld (ld = d+1 for an .fvecs image).  threads: sort the
 * split columns concurrently.  Returns 0, or -1 on bad arguments / out of memory. */
int pqh_partition_build(pqh_partition_t *p, const float *rows, long long n, long long ld,
                        int num_splits, int blocks_per_dim, double overlap, int threads);
/* the reference's --blocks-info-cache file (fast_nn_blocks_info.c:118-165), same bytes */
int pqh_partition_save(const pqh_partition_t *p, const char *path);
int pqh_partition_load(pqh_partition_t *p, const char *path); /* 0 = loaded, -1 = no such file */
void pqh_partition_free(pqh_partition_t *p);

/* cols[i*num_splits + s] = coordinate splits[s].dimension of vector i (malloc'ed; free() it) */
float *pqh_partition_columns(const pqh_partition_t *p, const float *rows, long long n, long long ld);

/* Does the vector whose split coordinates are `c` (one row of pqh_partition_columns) lie in
 * block `block_id`, looking only at splits [s_begin, s_end)? */
int pqh_partition_contains(const pqh_partition_t *p, const float *c, long long block_id,
                           int s_begin, int s_end);

/* Row ids of block `block_id` among rows [row_begin, row_end), ascending, written to out;
 * returns how many.  When prefix_len > 0 a row must also lie in the first prefix_len splits of
 * block prefix_block (the loader's shortcut for passes of several blocks,
 * fast_nn_block_loader2.c:111-118). */
long long pqh_partition_block_rows(const pqh_partition_t *p, const float *cols, long long row_begin,
                                   long long row_end, long long block_id, long long prefix_block,
                                   int prefix_len, int *out);

#ifdef __cplusplus
}
#endif
#endif /* PQH_PARTITION_H_ */
