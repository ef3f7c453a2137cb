/*
 * Stand-in for Yael v438's <yael/nn.h>, which the reference includes at
 * compute_nn_fast.c:11 but does not vendor.  Only the one prototype the
 * reference uses is declared (call site compute_nn_fast.c:478-481).  Written
 * from the call site and Yael's documented signature; no Yael source is copied.
 */
#ifndef PQH_YAEL_NN_STUB_H_
#define PQH_YAEL_NN_STUB_H_
#ifdef __cplusplus
extern "C" {
#endif
void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *v, const float *b_weights, int *assign, float *dis,
                     int n_thread);
#ifdef __cplusplus
}
#endif
#endif
