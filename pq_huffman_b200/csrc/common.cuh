// common.cuh -- shared device helpers for the kNN kernels (sm_100a).
//
// Candidate bookkeeping shared by the exact FP32 kernel (knn_exact.cu) and the
// tensor-core kernel (knn_tc.cu): a neighbour candidate is one 64-bit key
//     key = (order-preserving bits of the score) << 32 | base position
// so that "smaller key" == "better neighbour, ties to the lower position".
// That is the oracle's order (oracle/knn_oracle.c: ascending (distance,
// position)); positions are unique per query row, so keys are unique and every
// sort/merge below is over a strict total order -- no tie ambiguity anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqh {

typedef unsigned long long key64_t;
static constexpr key64_t KEY_INF = 0xFFFFFFFFFFFFFFFFull;

// Any float -> uint32 whose unsigned order equals the float order.
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ key64_t make_key(float score, uint32_t pos) {
    return ((key64_t)float_to_ordered(score) << 32) | (key64_t)pos;
}
__host__ __device__ __forceinline__ float key_score(key64_t k) {
    return ordered_to_float((uint32_t)(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_pos(key64_t k) { return (uint32_t)k; }

// ---------------------------------------------------------------------------
// Warp-cooperative bitonic sort of n <= NMAX keys held in shared memory
// (ascending).  NMAX is a power of two; slots [n, NMAX) are filled with KEY_INF.
// All 32 lanes of the warp must call it.
template <int NMAX>
__device__ __forceinline__ void warp_sort_keys(key64_t *s, int n, int lane) {
    for (int i = n + lane; i < NMAX; i += 32) s[i] = KEY_INF;
    __syncwarp();
#pragma unroll 1
    for (int size = 2; size <= NMAX; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < NMAX / 2; t += 32) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                key64_t a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncwarp();
        }
    }
}

// Same, sized at run time: sorts the first `npow2` slots (a power of two >= n, >= 32), so a row
// with few keys does not pay for the capacity of the buffer.
__device__ __forceinline__ void warp_sort_keys_n(key64_t *s, int n, int lane) {
    int npow2 = 32;
    while (npow2 < n) npow2 <<= 1;
    for (int i = n + lane; i < npow2; i += 32) s[i] = KEY_INF;
    __syncwarp();
#pragma unroll 1
    for (int size = 2; size <= npow2; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < npow2 / 2; t += 32) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                key64_t a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncwarp();
        }
    }
}

// number of elements of sorted a[0..n) that are < x
__device__ __forceinline__ int lower_bound_keys(const key64_t *a, int n, key64_t x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Warp-cooperative merge of two sorted key sequences with disjoint keys into
// out[0 .. min(cap, na+nb)).  Returns the output length.  out must not alias.
__device__ __forceinline__ int warp_merge_keys(const key64_t *a, int na, const key64_t *b, int nb,
                                               key64_t *out, int cap, int lane) {
    for (int i = lane; i < na; i += 32) {
        key64_t x = a[i];
        int p = i + lower_bound_keys(b, nb, x);
        if (p < cap) out[p] = x;
    }
    for (int j = lane; j < nb; j += 32) {
        key64_t x = b[j];
        int p = j + lower_bound_keys(a, na, x);
        if (p < cap) out[p] = x;
    }
    int n = na + nb;
    return n < cap ? n : cap;
}

}  // namespace pqh
