// nn_stats.cu -- joint code histograms of the kNN graph for nn_huffman_estimator.
//
// The reference's estimator counts, for every neighbour rank j and every PQ sub-quantiser i, how
// often a vector with code a has a j-th neighbour with code b (collect_stats_at_kth_nn,
// nn_huffman_estimator.c:69-107): stats[i][a][b] += 1 over all vectors whose j-th neighbour is
// known (index != (vector_id_t)-1), and returns how many were.  The counts are integers far
// below 2^53, so the doubles the reference accumulates are exact whatever the order: here all
// ranks are counted at once with integer atomics and converted at the end.
#include <cstdint>

#include "kernels.h"

namespace pqh {

namespace {

constexpr uint32_t NN_UNKNOWN = 0xffffffffu;  // (vector_id_t)-1, misc.h:4

// one thread per (vector, rank)
__global__ void __launch_bounds__(256)
nn_hist_kernel(const uint8_t *__restrict__ pq, const uint32_t *__restrict__ nn, long long n, int m,
               int k_star, int num_nn, int nn_first, int nn_count, uint32_t *__restrict__ cnt,
               unsigned long long *__restrict__ known, uint32_t *__restrict__ flags) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * nn_count) return;
    const long long row = g / nn_count;
    const int j = (int)(g % nn_count);
    const uint32_t t = nn[row * num_nn + nn_first + j];
    if (t == NN_UNKNOWN) return;
    if ((long long)t >= n) {  // the reference would read past its array
        atomicOr(flags, 1u);
        return;
    }
    atomicAdd(known + j, 1ull);
    const uint8_t *src = pq + row * m, *dst = pq + (long long)t * m;
    for (int i = 0; i < m; ++i) {
        const uint32_t a = src[i], b = dst[i];
        if (a >= (uint32_t)k_star || b >= (uint32_t)k_star) {
            atomicOr(flags, 2u);
            continue;
        }
        atomicAdd(cnt + (((size_t)j * m + i) * k_star + a) * k_star + b, 1u);
    }
}

__global__ void __launch_bounds__(256)
nn_hist_to_double_kernel(const uint32_t *__restrict__ cnt, size_t total, double *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = (double)cnt[i];
}

}  // namespace

// d_cnt: nn_count * m * k_star^2 words, d_known: nn_count 64-bit words, d_flags: one word; all
// three are cleared here.  d_out receives the counts as doubles.
cudaError_t launch_nn_histograms(const void *d_pq, const void *d_nn, long long n, int m, int k_star,
                                 int num_nn, int nn_first, int nn_count, uint32_t *d_cnt,
                                 unsigned long long *d_known, uint32_t *d_flags, double *d_out,
                                 cudaStream_t stream, int *launches) {
    const size_t total = (size_t)nn_count * m * k_star * k_star;
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_cnt, 0, total * sizeof(uint32_t), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_known, 0, (size_t)nn_count * sizeof(unsigned long long), stream)) !=
        cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(d_flags, 0, sizeof(uint32_t), stream)) != cudaSuccess) return e;
    const long long work = n * nn_count;
    if (work > 0) {
        nn_hist_kernel<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(
            static_cast<const uint8_t *>(d_pq), static_cast<const uint32_t *>(d_nn), n, m, k_star,
            num_nn, nn_first, nn_count, d_cnt, d_known, d_flags);
        ++*launches;
    }
    nn_hist_to_double_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_cnt, total, d_out);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pqh
