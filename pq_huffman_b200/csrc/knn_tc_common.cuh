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
    uint32_t tile_cursor;  // base tile most recently started by some CTA (L2-friendly order)
    uint32_t pad[3];
    unsigned long long cand_total[64];
    unsigned long long dbg[8];  // epilogue event counters: slow chunks, appends, hard/idle shrinks, chunks  // candidates re-ranked (spread to avoid contention)
};

struct TcArgs {
    int n_rows;      // query rows in this launch
    int q_begin;     // absolute query index of row 0 (self-join: also its base position)
    int q_op_begin;  // row of row 0 inside the A operand tensor
    int nb;
    int self_join;
    const int *row_ids;  // second pass: absolute query id of each row (null: q_begin + row)
    int n_qtiles, nsplit, tiles_per_split, n_btiles;
    // Tail of the work list: the last n_qtiles - q_tail0 query tiles (the partly filled last
    // round of the persistent grid) are cut into ns_tail base splits each so that every SM gets
    // a piece; their rows own ns_tail candidate lists instead of nsplit.
    int q_tail0, ns_tail, tiles_per_split_tail;
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

struct ItemInfo {
    int qt, sp, t_lo, t_hi;
};
__device__ __forceinline__ int total_items(const TcArgs &a) {
    return a.q_tail0 * a.nsplit + (a.n_qtiles - a.q_tail0) * a.ns_tail;
}
__device__ __forceinline__ ItemInfo decode_item(const TcArgs &a, int item) {
    ItemInfo it;
    const int n_main = a.q_tail0 * a.nsplit;
    int tps;
    if (item < n_main) {
        it.qt = item / a.nsplit;
        it.sp = item % a.nsplit;
        tps = a.tiles_per_split;
    } else {
        const int j = item - n_main;
        it.qt = a.q_tail0 + j / a.ns_tail;
        it.sp = j % a.ns_tail;
        tps = a.tiles_per_split_tail;
    }
    it.t_lo = it.sp * tps;
    it.t_hi = it.t_lo + tps;
    if (it.t_hi > a.n_btiles) it.t_hi = a.n_btiles;
    return it;
}
// candidate-list slot of (query row r of this launch, base split sp)
__host__ __device__ __forceinline__ size_t list_slot(int r, int sp, int tail_row0, int nsplit,
                                                      int ns_tail) {
    return r < tail_row0 ? (size_t)r * nsplit + sp
                         : (size_t)tail_row0 * nsplit + (size_t)(r - tail_row0) * ns_tail + sp;
}

// Base-tile order of one work item.  Every CTA sweeps ALL tiles of its range, so the order is
// free: CTA 0 never restarts -- it keeps cycling through the tiles and publishes its position --
// and every other CTA starts a new item at CTA 0's current tile and wraps around.  All CTAs then
// stream the same window of the base operand at any moment, and a tile is fetched from HBM once
// per sweep of the pack instead of once per CTA.  (Single writer: a cursor written by every CTA
// never converges.)
__device__ __forceinline__ int item_start_tile(const PrepScalars *sc, int t_lo, int t_hi,
                                               int my_next) {
    int c = my_next;
    if (blockIdx.x != 0)
        c = (int)*reinterpret_cast<const volatile uint32_t *>(&sc->tile_cursor);
    return (c >= t_lo && c < t_hi) ? c : t_lo;
}
__device__ __forceinline__ int item_tile(int i, int t0, int t_lo, int t_hi) {
    int t = t0 + i;
    return t >= t_hi ? t - (t_hi - t_lo) : t;
}
__device__ __forceinline__ void publish_cursor(const PrepScalars *sc, int t) {
    if (blockIdx.x == 0)
        *reinterpret_cast<volatile uint32_t *>(&const_cast<PrepScalars *>(sc)->tile_cursor) =
            (uint32_t)t;
}

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

// Warp-cooperative shrink of one row's candidate list (n <= NMAX keys in global memory) to its
// best L .. L+L/2 keys.  A bisection on the ordered score bits, bracketed by the list's own
// min/max and counted with ballots, stops at the first pivot that keeps between L and 1.5 L keys
// (typically < 10 steps); only when equal scores make that impossible does it run to the exact
// L-th score and break the tie by position in the list.  Survivors are written back compacted
// (unsorted); new_cnt receives their number.  Returns the pivot's ordered bits: every dropped
// key has a strictly worse score.  Registers, ballots and two warp reductions only -- no
// shared memory, no sort: a shrink has to cost about one tile time, not several.
template <int NMAX>
__device__ __noinline__ uint32_t shrink_row(key64_t *list, int n, int L, int lane, int &new_cnt) {
    constexpr int PER = NMAX / 32;
    key64_t k[PER];
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = i * 32 + lane;
        k[i] = idx < n ? __ldcg(list + idx) : KEY_INF;
        const uint32_t h = (uint32_t)(k[i] >> 32);
        if (idx < n) {
            mn = h < mn ? h : mn;
            mx = h > mx ? h : mx;
        }
    }
    uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = __reduce_max_sync(0xffffffffu, mx);
    const int Lmax = L + L / 2;
    bool early = false;
    uint32_t x = hi;
    int c_keep = n;
#pragma unroll 1
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i)
            c += __popc(__ballot_sync(0xffffffffu, (uint32_t)(k[i] >> 32) <= mid));
        if (c >= L) {
            if (c <= Lmax) {
                early = true;
                x = mid;
                c_keep = c;
                break;
            }
            hi = mid;
        } else {
            lo = mid + 1;
        }
    }
    int quota = 0x7fffffff;  // keys with score == x that survive (all of them on the early path)
    if (!early) {
        x = lo;  // exact L-th smallest ordered score; ties at x are cut by list position
        int c_lt = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i)
            c_lt += __popc(__ballot_sync(0xffffffffu, (uint32_t)(k[i] >> 32) < x));
        quota = L - c_lt;
        c_keep = L;
    }
    __syncwarp();  // all loads done before the list is overwritten
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
        if (!early) quota -= taken < quota ? taken : quota;
    }
    __syncwarp();
    new_cnt = c_keep < base ? c_keep : base;
    return x;
}

// Predicated append of (ordered(-score), id) to a row's candidate list: no branch.
__device__ __forceinline__ void append_if(bool p, key64_t *list, int &cnt, uint32_t score_bits,
                                          int id) {
    // float_to_ordered(-score) straight from the bits of score: two instructions
    // hi = bits ^ (~sign & 0x7fffffff)
    uint32_t sign, hi;
    asm("shr.s32 %0, %1, 31;" : "=r"(sign) : "r"(score_bits));
    asm("lop3.b32 %0, %1, %2, 0x7fffffff, 0xD2;" : "=r"(hi) : "r"(score_bits), "r"(sign));
    const key64_t key = ((key64_t)hi << 32) | (key64_t)(uint32_t)id;
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

// Epilogue of one accumulator tile for one warp: 32 query rows (one per thread, the TMEM lane
// it may read) x TC_BN columns starting at TMEM address tbase / base position col_base.
//
// Hot path (kept small and contiguous -- it must live in the instruction cache): copy the
// 32 x 128 slice to registers with two 64-column tcgen05.ld in flight together, reduce it to
// 16 group maxima (groups of 8 columns) with 3-input max, compare with the row's threshold tau
// and OR the 16-bit result over the warp.  No bit set (the common case): release the
// accumulator stage and return.
// Slow path (one small loop, no unrolled copies): for every flagged group re-read its 8 columns
// from TMEM by address and append the scores above tau to the row's candidate list with
// predicated stores (no divergent branches); then release the stage, then -- rarely -- shrink a
// list that reached its hard capacity.
// tau_shared (pair kernel): ordered bits of the threshold shared by the two threads that
// filter the two column halves of the same row.
template <int GCAP, int NCOLS, class Release>
__device__ __forceinline__ void filter_tile_n(uint32_t tbase, int col_base, bool special, int ib_hi,
                                            int self, int L, int lane, int debug_mode,
                                            uint32_t *tau_shared, float &tau, int &cnt,
                                            key64_t *list, unsigned (&ctr)[4], Release release,
                                            long long *tprof = nullptr) {
    const long long tp0 = tprof ? clock64() : 0;
    static_assert(NCOLS == 128 || NCOLS == 80, "tile widths: 64 + 64 or 64 + 16 columns");
    constexpr int NG = NCOLS / 8;  // groups of 8 columns
    unsigned groups = 0;
    {
        uint32_t ra[64], rb[NCOLS - 64];
        ptx::tmem_ld_32x32b_x64(tbase, ra);
        if constexpr (NCOLS == 128) ptx::tmem_ld_32x32b_x64(tbase + 64, rb);
        else ptx::tmem_ld_32x32b_x16(tbase + 64, rb);
        ptx::tmem_wait_ld();
        // group maxima (groups of 8 columns), then the row maximum
        float gm[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) {
            const uint32_t *buf = (q < 8) ? &ra[q * 8] : &rb[(q - 8) * 8];
            float m = ptx::fmax3(__uint_as_float(buf[0]), __uint_as_float(buf[1]),
                                 __uint_as_float(buf[2]));
            m = ptx::fmax3(m, __uint_as_float(buf[3]), __uint_as_float(buf[4]));
            m = ptx::fmax3(m, __uint_as_float(buf[5]), __uint_as_float(buf[6]));
            gm[q] = fmaxf(m, __uint_as_float(buf[7]));
        }
        float rm = gm[0];
#pragma unroll
        for (int q = 1; q + 1 < NG; q += 2) rm = ptx::fmax3(rm, gm[q], gm[q + 1]);
        if ((NG & 1) == 0) rm = fmaxf(rm, gm[NG - 1]);
        // common case: no row of the warp has anything above its threshold -> one vote, done
        if (__any_sync(0xffffffffu, rm > tau)) {
            unsigned gmask = 0;
#pragma unroll
            for (int q = 0; q < NG; ++q) gmask |= (gm[q] > tau) ? (1u << q) : 0u;
            groups = __reduce_or_sync(0xffffffffu, gmask);
        }
    }
    const long long tp1 = tprof ? clock64() : 0;
    if (groups == 0 || debug_mode == 2) {
        release();
        if (tprof) {
            tprof[0] += tp1 - tp0;
            tprof[1] += clock64() - tp1;
        }
        return;
    }
    ++ctr[0];
    const int id_lim = special ? ib_hi : 0x7fffffff;
    // the id checks (ragged last tile, the row's own column) are only compiled into the copy of
    // the scan that "special" tiles take; the choice is made warp-uniform
    const bool any_special = __any_sync(0xffffffffu, special);
    auto scan8 = [&](const uint32_t(&v8)[8], int q) {
        const int col0 = col_base + q * 8;
        if (any_special) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int id = col0 + j;
                const bool p = (__uint_as_float(v8[j]) > tau) & (id < id_lim) & (id != self);
                append_if(p, list, cnt, v8[j], id);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                append_if(__uint_as_float(v8[j]) > tau, list, cnt, v8[j], col0 + j);
        }
    };
    const int q1 = __ffs(groups) - 1;
    const unsigned rest = groups & (groups - 1);
    if ((rest & (rest - 1)) == 0) {
        // one or two flagged groups (the usual slow tile): fetch them, give the accumulator
        // stage back, and only then scan -- the scan is off the tensor cores' critical path
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
    if (cnt > GCAP) __trap();  // list overflow would silently corrupt a neighbouring row's list
    const long long tp2 = tprof ? clock64() : 0;
    // hard capacity: a tile can append up to NCOLS entries before this check runs again
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
        ++ctr[1];
        if (lane == src) {
            cnt = kept;
            tau = fmaxf(tau, -ordered_to_float(x));
            if (tau_shared) atomicMax(tau_shared, float_to_ordered(tau));
        }
    }
    if (tprof) {
        tprof[0] += tp1 - tp0;          // TMEM -> registers, group maxima
        tprof[2] += tp2 - tp1;          // slow path up to the release
        tprof[3] += clock64() - tp2;    // hard shrinks
    }
}

template <int GCAP, class Release>
__device__ __forceinline__ void filter_tile(uint32_t tbase, int col_base, bool special, int ib_hi,
                                            int self, int L, int lane, int debug_mode,
                                            uint32_t *tau_shared, float &tau, int &cnt,
                                            key64_t *list, unsigned (&ctr)[4], Release release,
                                            long long *tprof = nullptr) {
    filter_tile_n<GCAP, TC_BN>(tbase, col_base, special, ib_hi, self, L, lane, debug_mode,
                               tau_shared, tau, cnt, list, ctr, release, tprof);
}

// While the tensor cores have not finished the next accumulator tile, use the idle time to
// shrink lists that have grown past 2L: the threshold then stays fresh without a shrink ever
// standing between a finished accumulator stage and its release (mandatory shrinks at the hard
// capacity GCAP-32 remain in filter_tile, but become rare).
template <int GCAP>
__device__ __forceinline__ void wait_acc_with_idle_shrinks(uint64_t *bar, uint32_t parity, int L,
                                                           int lane, uint32_t *tau_shared,
                                                           float &tau, int &cnt, key64_t *list,
                                                           unsigned (&ctr)[4]) {
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
        ++ctr[2];
        if (lane == src) {
            cnt = kept;
            tau = fmaxf(tau, -ordered_to_float(x));
            if (tau_shared) atomicMax(tau_shared, float_to_ordered(tau));
        }
    }
}


}  // namespace
}  // namespace pqh
