/* oracle/ref_est_shim.c -- test infrastructure only.  Gives the tests a handle on a function
 * that is static in the reference: collect_stats_at_kth_nn (nn_huffman_estimator.c:69-107).  The
 * reference's source file is included where it lies (-I/root/reference/src at build time,
 * nothing is copied), its main renamed; the result is oracle/_ref/libref_est.so. */
#include <string.h>

#define main ref_estimator_main
#include "nn_huffman_estimator.c"
#undef main

long long ref_collect_stats(long long n, int m, int k_star, int num_nn, const unsigned char *pq,
                            const unsigned *nn, double *stats, int nn_index) {
    config_t c;
    memset(&c, 0, sizeof(c));
    c.m = m;
    c.k_star = k_star;
    c.num_vectors = n;
    c.num_nn = num_nn;
    return collect_stats_at_kth_nn(&c, pq, nn, stats, nn_index);
}
