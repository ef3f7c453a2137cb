// knn_tc_i8.cuh -- integer flavour of the tensor-core filter (launch_knn_tc takes it whenever
// the data qualifies; $PQH_TC_I8=0 disables it).
//
// For data whose values are integers in [0, 255] (SIFT-shaped uint8 stored as float32,
// vecs_io.c:43-68 reads it that way) the score can be computed EXACTLY on the INT8 tensor cores
// (tcgen05.mma kind::i8, u8 x s8 -> s32, twice the FP16 rate):
//     A (queries) = x as u8,  B (base) = y' = y - 128 as s8,
//     32 extra K columns fold G(y) = ceil(|y|^2 / 2):  A side constants (255 x30, 1, 1),
//     B side s8 digits b with sum(c_i b_i) = -G   (capacity: |y|^2 <= 1958398),
//     accumulator T = x.y' - G,   dist = |x|^2 - 256 sum(x) - 2 T - (2G - |y|^2),  2G-|y|^2 in {0,1}.
// Larger T == smaller distance within a row; the only slack the certificate needs is that 0/1.
// The epilogue is the float one (knn_tc_common.cuh) with integer compares; candidate keys keep
// the same layout (high word: order-reversing map of the score, low word: base position), so
// shrink_row and the re-rank kernel are shared.
#pragma once
#include "knn_tc_common.cuh"

namespace pqh {
namespace {

constexpr int I8_FOLD_FULL = 30;      // fold columns with A-side constant 255 (the last two: 1)
constexpr int I8_FOLD_UNIT = 255 * 128;  // what one full column can subtract

__device__ __forceinline__ uint32_t i32_score_to_hi(int s) { return (uint32_t)s ^ 0x7fffffffu; }
__device__ __forceinline__ int i32_hi_to_score(uint32_t h) { return (int)(h ^ 0x7fffffffu); }

__device__ __forceinline__ void append_if_i32(bool p, key64_t *list, int &cnt, int score, int id) {
    const key64_t key = ((key64_t)i32_score_to_hi(score) << 32) | (key64_t)(uint32_t)id;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %0, 0;\n\t"
        "@q st.global.b64 [%1], %2;\n\t"
        "}\n"
        :
        : "r"((int)p), "l"(list + cnt), "l"(key)
        : "memory");
    cnt += p ? 1 : 0;
}

__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }

// filter_tile_n<GCAP, 128> with s32 scores; same structure, no instrumentation.
template <int GCAP, class Release>
__device__ __forceinline__ void filter_tile_i32(uint32_t tbase, int col_base, bool special,
                                                int ib_hi, int self, int L, int lane, int &tau,
                                                int &cnt, key64_t *list, Release release) {
    constexpr int NCOLS = TC_BN, NG = NCOLS / 8;
    unsigned groups = 0;
    {
        uint32_t ra[64], rb[64];
        ptx::tmem_ld_32x32b_x64(tbase, ra);
        ptx::tmem_ld_32x32b_x64(tbase + 64, rb);
        ptx::tmem_wait_ld();
        int gm[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) {
            const uint32_t *buf = (q < 8) ? &ra[q * 8] : &rb[(q - 8) * 8];
            int m = imax3((int)buf[0], (int)buf[1], (int)buf[2]);
            m = imax3(m, (int)buf[3], (int)buf[4]);
            m = imax3(m, (int)buf[5], (int)buf[6]);
            gm[q] = max(m, (int)buf[7]);
        }
        int rm = gm[0];
#pragma unroll
        for (int q = 1; q + 1 < NG; q += 2) rm = imax3(rm, gm[q], gm[q + 1]);
        rm = max(rm, gm[NG - 1]);
        if (__any_sync(0xffffffffu, rm > tau)) {
            unsigned gmask = 0;
#pragma unroll
            for (int q = 0; q < NG; ++q) gmask |= (gm[q] > tau) ? (1u << q) : 0u;
            groups = __reduce_or_sync(0xffffffffu, gmask);
        }
    }
    if (groups == 0) {
        release();
        return;
    }
    const int id_lim = special ? ib_hi : 0x7fffffff;
    const bool any_special = __any_sync(0xffffffffu, special);
    auto scan8 = [&](const uint32_t(&v8)[8], int q) {
        const int col0 = col_base + q * 8;
        if (any_special) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int id = col0 + j;
                const bool p = ((int)v8[j] > tau) & (id < id_lim) & (id != self);
                append_if_i32(p, list, cnt, (int)v8[j], id);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                append_if_i32((int)v8[j] > tau, list, cnt, (int)v8[j], col0 + j);
        }
    };
    const int q1 = __ffs(groups) - 1;
    const unsigned rest = groups & (groups - 1);
    if ((rest & (rest - 1)) == 0) {
        const int q2 = rest ? __ffs(rest) - 1 : q1;
        uint32_t va[8], vb[8];
        ptx::tmem_ld_32x32b_x8(tbase + q1 * 8, va);
        ptx::tmem_ld_32x32b_x8(tbase + q2 * 8, vb);
        ptx::tmem_wait_ld();
        release();
        scan8(va, q1);
        if (rest) scan8(vb, q2);
    } else {
#pragma unroll 1
        do {
            const int q = __ffs(groups) - 1;
            groups &= groups - 1;
            uint32_t v8[8];
            ptx::tmem_ld_32x32b_x8(tbase + q * 8, v8);
            ptx::tmem_wait_ld();
            scan8(v8, q);
        } while (groups != 0);
        release();
    }
    if (cnt > GCAP) __trap();
    unsigned need = __ballot_sync(0xffffffffu, cnt > GCAP - NCOLS);
#pragma unroll 1
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        key64_t *l = reinterpret_cast<key64_t *>(
            __shfl_sync(0xffffffffu, (unsigned long long)list, src));
        const int n = __shfl_sync(0xffffffffu, cnt, src);
        int kept;
        const uint32_t x = shrink_row<GCAP>(l, n, L, lane, kept);
        if (lane == src) {
            cnt = kept;
            tau = max(tau, i32_hi_to_score(x));
        }
    }
}

template <int GCAP>
__device__ __forceinline__ void wait_acc_with_idle_shrinks_i32(uint64_t *bar, uint32_t parity, int L,
                                                               int lane, int &tau, int &cnt,
                                                               key64_t *list) {
    const int soft = (L + L / 2 < GCAP - TC_BN) ? L + L / 2 : GCAP - TC_BN;
    while (!ptx::mbar_test_wait(bar, parity)) {
        const unsigned need = __ballot_sync(0xffffffffu, cnt > soft);
        if (need == 0) {
            ptx::mbar_wait(bar, parity);
            return;
        }
        const int src = __ffs(need) - 1;
        key64_t *l = reinterpret_cast<key64_t *>(
            __shfl_sync(0xffffffffu, (unsigned long long)list, src));
        const int n = __shfl_sync(0xffffffffu, cnt, src);
        int kept;
        const uint32_t x = n <= 64    ? shrink_row<64>(l, n, L, lane, kept)
                           : n <= 128 ? shrink_row<128>(l, n, L, lane, kept)
                                      : shrink_row<GCAP>(l, n, L, lane, kept);
        if (lane == src) {
            cnt = kept;
            tau = max(tau, i32_hi_to_score(x));
        }
    }
}

}  // namespace
}  // namespace pqh
