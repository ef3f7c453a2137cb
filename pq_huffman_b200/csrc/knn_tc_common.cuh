// knn_tc_common.cuh -- definitions shared by the tensor-core kernels (knn_tc.cu, knn_tc3.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pqh {
namespace {

constexpr int TC_ROWS = 256;    // query rows per CTA: two M=128 accumulator tiles
constexpr int TC_BN = 128;      // base rows per tile (UMMA N)
constexpr int TC_NSTAGE = 4;    // B-tile ring
constexpr int TC_MAXSEG = 6;
constexpr int TC_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_EPI_WARPS = 8;
constexpr int FOLD_COLS = 16;

struct TcSeg {
    int k_off;   // element offset in the fp16 operand row (data segments)
    int bytes;   // row bytes of this K segment: 128 / 64 / 32 (= swizzle span)
    int fold;    // 1: the |y|^2 fold segment (separate tensors)
    int a_off;   // byte offset inside the A region
    int b_off;   // byte offset inside one B stage
};

struct TcPlan {
    int nseg;
    int a_bytes;  // A region bytes (256 rows)
    int b_bytes;  // one B stage (128 rows)
    TcSeg seg[TC_MAXSEG];
};

struct PrepScalars {
    uint32_t maxabs_bits;  // max |x| over base and queries
    uint32_t ymax_bits;    // max scaled base norm |y'| (float bits, rounded up)
    uint32_t inexact;      // some fp16(sigma*x) != sigma*x
    uint32_t fb_count;     // rows that failed the certificate
    uint32_t pad[4];
    unsigned long long cand_total[64];
    unsigned long long dbg[8];  // epilogue event counters: slow chunks, appends, hard/idle shrinks, chunks  // candidates re-ranked (spread to avoid contention)
};

struct TcArgs {
    int n_rows;      // query rows in this launch
    int q_begin;     // absolute query index of row 0 (self-join: also its base position)
    int q_op_begin;  // row of row 0 inside the A operand tensor
    int nb;
    int self_join;
    int n_qtiles, nsplit, tiles_per_split, n_btiles;
    int kcap;        // neighbours needed per row (k, or k-1 for a self-join)
    int nstage;      // pair kernel: depth of the B-tile ring
    int debug_mode;  // timing experiments only: 1 = epilogue does not read TMEM, 2 = no slow path
    int L_exact, L_inexact;  // candidates kept per row (data exactly representable in fp16 / not)
    const PrepScalars *sc;
    key64_t *cand;
    int *cand_cnt;
    float *cand_tau;
    TcPlan plan;
};

struct TcMaps {
    CUtensorMap a[3];  // data operand, box 256 rows: SW128 / SW64 / SW32
    CUtensorMap b[3];  // data operand, box 128 rows
    CUtensorMap fa;    // fold constants [256 x 16]
    CUtensorMap fb;    // fold pieces    [nb x 16]
};

__device__ __forceinline__ int seg_class(int bytes) { return bytes == 128 ? 0 : (bytes == 64 ? 1 : 2); }
__device__ __forceinline__ uint32_t seg_layout(int bytes) {
    return bytes == 128 ? 2u : (bytes == 64 ? 4u : 6u);
}

// power-of-two scale putting maxabs in [2^7, 2^8)
__host__ __device__ __forceinline__ float scale_from_maxabs_bits(uint32_t bits) {
    int e = (int)((bits >> 23) & 0xff) - 127;  // maxabs in [2^e, 2^(e+1))
    if ((bits & 0x7fffffffu) == 0 || e < -100) return 1.0f;
    int se = 7 - e;
    if (se > 100) se = 100;
    if (se < -100) se = -100;
    uint32_t sb = (uint32_t)(se + 127) << 23;
#ifdef __CUDA_ARCH__
    return __uint_as_float(sb);
#else
    union { uint32_t u; float f; } c; c.u = sb; return c.f;
#endif
}

// Warp-cooperative shrink of one row's candidate list (n <= GCAP keys in global memory) to
// its L best: a 32-step binary search on the ordered score bits finds the L-th best score,
// survivors are written back compacted (unsorted).  Returns that score's ordered bits: every
// dropped key has a score that is not better.  No sort, no shared memory.
template <int GCAP>
__device__ __noinline__ uint32_t shrink_row(key64_t *list, int n, int L, int lane) {
    constexpr int PER = GCAP / 32;
    key64_t k[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = i * 32 + lane;
        k[i] = idx < n ? __ldcg(list + idx) : KEY_INF;
    }
    uint32_t lo = 0u, hi = 0xFFFFFFFFu;
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) c += ((uint32_t)(k[i] >> 32) <= mid) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= L) hi = mid; else lo = mid + 1;
    }
    const uint32_t x = lo;  // L-th smallest ordered score
    int c_lt = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) c_lt += ((uint32_t)(k[i] >> 32) < x) ? 1 : 0;
    c_lt = __reduce_add_sync(0xffffffffu, c_lt);
    int quota = L - c_lt;  // how many of the keys with score == x survive
    __syncwarp();          // all loads done before the list is overwritten
    int base = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const uint32_t h = (uint32_t)(k[i] >> 32);
        const unsigned eq = __ballot_sync(0xffffffffu, h == x);
        const int eq_rank = __popc(eq & lt_mask);
        const bool keep = (h < x) || (h == x && eq_rank < quota);
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        if (keep) __stcg(list + base + __popc(kb & lt_mask), k[i]);
        base += __popc(kb);
        const int taken = __popc(eq);
        quota -= taken < quota ? taken : quota;
    }
    __syncwarp();
    return x;
}


}  // namespace
}  // namespace pqh
