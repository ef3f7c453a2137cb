// knn_tc_common.cuh -- definitions shared by the tensor-core kernels (knn_tc.cu: query tile
// resident in shared memory; knn_tcs.cu: both operands streamed): work decomposition, the
// candidate-list bookkeeping and the whole filter epilogue.
//
// Scores.  FP16 operands give an FP32 score S' (larger = nearer); the integer flavour for data
// made of integers in [0, 255] (tcgen05.mma kind::i8, u8 x s8 -> s32) gives the EXACT score
//     T = x.(y - 128) - ceil(|y|^2 / 2),   dist = |x|^2 - 256 sum(x) - 2 T - (0 or 1),
// so the only slack its certificate needs is that 0/1.  Score<I8> hides the difference; the
// candidate keys (high word: order-reversing map of the score, low word: base position) have the
// same layout in both, so shrink_row and the re-rank kernel are shared.
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
constexpr int FOLD_COLS = 16;
constexpr int TC_EPI_WARPS = 8;  // filter warps: each owns 32 query rows (a TMEM lane quadrant)
// warp 0 TMA, then `nmma` MMA issuers, then the filter warps
__host__ __device__ constexpr int tc_threads(int nmma) { return (1 + nmma + TC_EPI_WARPS) * 32; }

constexpr int I8_FOLD_FULL = 30;         // fold columns with A-side constant 255 (the last two: 1)
constexpr int I8_FOLD_UNIT = 255 * 128;  // what one full column can subtract

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
    uint32_t tile_cursor;  // base tile most recently started by CTA 0 (L2-friendly order)
    uint32_t work_next;    // work items handed out beyond the first one of every CTA (reset to 0
                           // before each launch)
    uint32_t eps_ratio_bits;  // PQH_TC_CHECK_EPS: max |S'_tensor - S'_fp64| / eps (float bits)
    uint32_t fb2_count;    // rows that failed the certificate of the second pass
    uint32_t i8_reject;    // integer path: bit 0 = a value is not an integer in [0, 255],
                           // bit 1 = |y|^2 beyond the capacity of the fold columns
    unsigned long long cand_total[32];  // candidates re-ranked (spread to avoid contention)
    unsigned long long dbg[16];         // [0],[1] PQH_TC_DEBUG: MMA-thread cycle counters;
                                        // [2] (warp, tile) pairs that took the slow path
};

struct TcArgs {
    int n_rows;      // query rows in this launch
    int q_begin;     // absolute query index of row 0 (self-join: also its base position)
    int q_op_begin;  // row of row 0 inside the A operand tensor
    int nb;
    int self_join;
    const int *row_ids;  // second pass: absolute query id of each row (null: q_begin + row)
    // Locality order (knn_order.cu): operand row p of the base is data row base_perm[p], query
    // row r of this launch is data row q_perm[r].  q_self[r] = position of query r's own vector
    // in the base operand (self-join; -1 otherwise) and qt_start[qt] = window of base tiles with
    // which query tile qt starts its sweep (the cells its rows belong to): first tile in the low
    // 24 bits, length above.  All null: natural order.
    const int *q_self;
    const int *qt_start;
    int n_qtiles, nsplit, tiles_per_split, n_btiles;
    // Tail of the work list: the last n_qtiles - q_tail0 query tiles (the partly filled last
    // round of the persistent grid) are cut into ns_tail base splits each so that every SM gets
    // a piece; their rows own ns_tail candidate lists instead of nsplit.
    int q_tail0, ns_tail, tiles_per_split_tail;
    int kcap;        // neighbours needed per row (k, or k-1 for a self-join)
    int debug_mode;  // timing experiments only: 1 = epilogue does not read TMEM, 2 = no slow path
    int dynamic;     // work items are handed out by a counter (sc->work_next) instead of round robin
    int L_exact, L_inexact;  // candidates kept per row (data exactly representable in fp16 / not)
    const PrepScalars *sc;
    key64_t *cand;
    int *cand_cnt;
    float *cand_tau;
    // arg-max mode (cell assignment of the locality order): the column with the best score of
    // every row and a histogram of those columns, instead of candidate lists
    int *pick_cell, *pick_hist;
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
// free, and two things are wanted from it:
//  * Locality order (qt_start given): the sweep STARTS with a window of a few tiles at
//    the cell the query tile's own rows belong to, so a row meets its near neighbours first and
//    its threshold is tight for the rest of the sweep.
//  * The rest of the range is swept cyclically from wherever CTA 0 currently is: CTA 0 never
//    restarts -- it keeps cycling through the tiles and publishes its position -- and every
//    other CTA joins the sweep there and wraps around.  All CTAs then stream the same window of
//    the base operand at any moment, and a tile is fetched from HBM once per sweep of the pack
//    instead of once per CTA (measured without it, every CTA starting at its own cell: 146 GB
//    of DRAM reads per SIFT1M launch instead of 5).  (Single writer: a cursor written by every
//    CTA never converges.)
// An order is (w0, uc): first tile of the own window and the start of the cyclic part, as an
// index into the range with the window cut out.
struct ItemOrder {
    int w0, uc;
};
__device__ __forceinline__ int own_window_len(const TcArgs &a, int qt, int t_lo, int t_hi) {
    if (!a.qt_start) return 0;
    const int n = t_hi - t_lo, w = (int)((uint32_t)a.qt_start[qt] >> 24);
    return n < w ? n : w;
}
// producer side: my_next = where CTA 0's own cyclic sweep continues
__device__ __forceinline__ ItemOrder item_order(const TcArgs &a, int qt, int t_lo, int t_hi,
                                                int my_next) {
    const int n = t_hi - t_lo, wlen = own_window_len(a, qt, t_lo, t_hi);
    ItemOrder o;
    o.w0 = t_lo;
    if (wlen) {
        int s = a.qt_start[qt] & 0xffffff;
        if (s < t_lo || s >= t_hi) s = t_lo;  // the cell lies in another base split
        o.w0 = s < t_hi - wlen ? s : t_hi - wlen;
    }
    int c = my_next;
    if (blockIdx.x != 0)
        c = (int)*reinterpret_cast<const volatile uint32_t *>(&a.sc->tile_cursor);
    if (c < t_lo || c >= t_hi) c = t_lo;
    o.uc = c < o.w0 ? c - t_lo : (c >= o.w0 + wlen ? c - t_lo - wlen : o.w0 - t_lo);
    if (o.uc >= n - wlen) o.uc = 0;
    return o;
}
// tile of step i of the item
__device__ __forceinline__ int item_tile(int i, ItemOrder o, int wlen, int t_lo, int t_hi) {
    if (i < wlen) return o.w0 + i;
    const int m = t_hi - t_lo - wlen;
    int u = o.uc + (i - wlen);
    if (u >= m) u -= m;
    const int t = t_lo + u;
    return t >= o.w0 ? t + wlen : t;
}
// CTA 0 publishes its position in the cyclic part
__device__ __forceinline__ void publish_cursor(const TcArgs &a, int t) {
    if (blockIdx.x == 0)
        *reinterpret_cast<volatile uint32_t *>(&const_cast<PrepScalars *>(a.sc)->tile_cursor) =
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

// ------------------------------------------------------------------------------------
// score flavours
template <bool I8>
struct Score;
template <>
struct Score<false> {  // FP32 accumulator of FP16 operands
    typedef float T;
    static __device__ __forceinline__ T lowest() { return __int_as_float(0xff800000); }
    static __device__ __forceinline__ T highest() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ T from_bits(uint32_t b) { return __uint_as_float(b); }
    static __device__ __forceinline__ T max2(T a, T b) { return fmaxf(a, b); }
    static __device__ __forceinline__ T max3(T a, T b, T c) { return ptx::fmax3(a, b, c); }
    // float_to_ordered(-score) straight from the accumulator bits: hi = bits ^ (~sign & 0x7fffffff)
    static __device__ __forceinline__ uint32_t hi(uint32_t bits) {
        uint32_t sign, h;
        asm("shr.s32 %0, %1, 31;" : "=r"(sign) : "r"(bits));
        asm("lop3.b32 %0, %1, %2, 0x7fffffff, 0xD2;" : "=r"(h) : "r"(bits), "r"(sign));
        return h;
    }
    static __device__ __forceinline__ T from_hi(uint32_t h) { return -ordered_to_float(h); }
    static __device__ __forceinline__ float as_float(T t) { return t; }
};
template <>
struct Score<true> {  // s32 accumulator of u8 x s8 operands
    typedef int T;
    static __device__ __forceinline__ T lowest() { return (int)0x80000000; }
    static __device__ __forceinline__ T highest() { return 0x7fffffff; }
    static __device__ __forceinline__ T from_bits(uint32_t b) { return (int)b; }
    static __device__ __forceinline__ T max2(T a, T b) { return max(a, b); }
    static __device__ __forceinline__ T max3(T a, T b, T c) { return max(max(a, b), c); }
    static __device__ __forceinline__ uint32_t hi(uint32_t bits) { return bits ^ 0x7fffffffu; }
    static __device__ __forceinline__ T from_hi(uint32_t h) { return (int)(h ^ 0x7fffffffu); }
    // |score| < 2^24: exact as a float; "no threshold yet" = -inf
    static __device__ __forceinline__ float as_float(T t) {
        return t == lowest() ? __int_as_float(0xff800000) : (float)t;
    }
};
// the same order-reversing maps on the host / in the re-rank kernel
__host__ __device__ __forceinline__ uint32_t i32_score_to_hi(int s) { return (uint32_t)s ^ 0x7fffffffu; }
__host__ __device__ __forceinline__ int i32_hi_to_score(uint32_t h) { return (int)(h ^ 0x7fffffffu); }

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

// Shrink the lists of the lanes flagged in `need` (warp-uniform), one after the other; the lane
// that owns a list takes its new length and raises its threshold to the pivot's score.
template <int GCAP, bool I8, bool SIZED>
__device__ __forceinline__ void shrink_flagged(unsigned need, int L, int lane,
                                               typename Score<I8>::T &tau, int &cnt,
                                               key64_t *list) {
    typedef Score<I8> S;
#pragma unroll 1
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        key64_t *l = reinterpret_cast<key64_t *>(
            __shfl_sync(0xffffffffu, (unsigned long long)list, src));
        const int n = __shfl_sync(0xffffffffu, cnt, src);
        int kept;
        uint32_t x;
        if (SIZED)  // idle-time shrinks run early, on short lists: size the register copy to fit
            x = n <= 64    ? shrink_row<64>(l, n, L, lane, kept)
                : n <= 128 ? shrink_row<128>(l, n, L, lane, kept)
                           : shrink_row<GCAP>(l, n, L, lane, kept);
        else
            x = shrink_row<GCAP>(l, n, L, lane, kept);
        if (lane == src) {
            cnt = kept;
            tau = S::max2(tau, S::from_hi(x));
        }
    }
}

// Predicated append of (hi(score), id) to a row's candidate list: no branch.
template <bool I8>
__device__ __forceinline__ void append_if(bool p, key64_t *list, int &cnt, uint32_t score_bits,
                                          int id) {
    const key64_t key = ((key64_t)Score<I8>::hi(score_bits) << 32) | (key64_t)(uint32_t)id;
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

// Cycle accounting of a filter warp (diagnostic builds only, -DPQH_TC_PROF: MODE 4 of
// knn_tc_kernel): with ON = false every call is empty and the object has no state.
template <bool ON>
struct EpiProf {
    uint32_t t[ON ? 8 : 1];
    uint32_t last;
    __device__ __forceinline__ void init() {
        if constexpr (ON) {
            for (int i = 0; i < 8; ++i) t[i] = 0;
            last = (uint32_t)clock();
        }
    }
    __device__ __forceinline__ void lap(int i) {
        if constexpr (ON) {
            const uint32_t now = (uint32_t)clock();
            t[i] += now - last;
            last = now;
        }
    }
    __device__ __forceinline__ void add(int i, uint32_t v) {
        if constexpr (ON) t[i] += v;
    }
};
enum { PROF_WAIT = 0, PROF_LD, PROF_HOT, PROF_SCAN, PROF_HARD, PROF_IDLE, PROF_N_IDLE, PROF_N_HARD };

// Epilogue of one accumulator tile for one warp: 32 query rows (one per thread, the TMEM lane
// it may read) x TC_BN columns starting at TMEM address tbase / base position col_base.
// The 32 x 128 slice is copied to registers with one 128-column tcgen05.ld.
//
// EARLY: the accumulator stage is released as soon as the copy has landed -- the tile then lives
//   in registers and the tensor cores never wait for the filtering, only for the copy.  (With a
//   short K the accumulator hand-offs dominate: one 256 x 128 kind::i8 tile is 640 cycles of MMA
//   against ~900 cycles of commit / wake-up / release latencies, and the two stages TMEM has room
//   for cannot hide a filter on top of that.)  Hot path: four independent running maxima with
//   3-input max, one per 32-column quarter, give the row maximum (the 128 values have to stay
//   live, so there is no register room for per-group maxima); one compare with the row's
//   threshold tau and one vote.  Nothing above tau in any row of the warp (the common case):
//   done.  Slow path: the quarters that hold something (one vote each; usually just one) are
//   gone through group by group of 8 columns (warp-uniform) and the scores above tau are
//   appended, straight from the registers, with predicated stores.
// !EARLY: 16 group maxima level by level, vote, then the flagged groups are re-read from TMEM by
//   address in one small loop and the stage is released after that (before the scan when at
//   most two groups are flagged).
// Either way a list that reached its hard capacity is shrunk (rare).
template <int GCAP, bool I8, bool EARLY, bool PROF, class Release>
__device__ __forceinline__ void filter_tile(uint32_t tbase, int col_base, bool special, int ib_hi,
                                            int self, int L, int lane, int debug_mode,
                                            typename Score<I8>::T &tau, int &cnt, key64_t *list,
                                            unsigned &slow_tiles, EpiProf<PROF> &prof,
                                            Release release) {
    typedef Score<I8> S;
    typedef typename S::T T;
    constexpr int NG = TC_BN / 8;  // groups of 8 columns
    // one 128-column tcgen05.ld (measured against two 64-column loads in flight: -1 % at 1M rows)
    uint32_t rr[128];
    ptx::tmem_ld_32x32b_x128(tbase, rr);
    uint32_t *const ra = rr, *const rb = rr + 64;
    ptx::tmem_wait_ld();
#define PQH_V(q, j) S::from_bits(((q) < 8) ? ra[(q) * 8 + (j)] : rb[((q) - 8) * 8 + (j)])
    if constexpr (EARLY) {
        release();
        prof.lap(PROF_LD);
        // "special" tiles (ragged last tile, the row's own column): take the forbidden columns
        // out here, once, so that the scan below tests scores only
        if (__any_sync(0xffffffffu, special)) {
            const int id_lim = special ? ib_hi : 0x7fffffff;
            uint32_t low;
            if constexpr (I8) low = 0x80000000u;
            else low = 0xff800000u;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                if (col_base + c >= id_lim || col_base + c == self) ra[c] = low;
                if (col_base + 64 + c >= id_lim || col_base + 64 + c == self) rb[c] = low;
            }
        }
        T m0 = S::max3(PQH_V(0, 0), PQH_V(0, 1), PQH_V(0, 2));
        T m1 = S::max3(PQH_V(4, 0), PQH_V(4, 1), PQH_V(4, 2));
        T m2 = S::max3(PQH_V(8, 0), PQH_V(8, 1), PQH_V(8, 2));
        T m3 = S::max3(PQH_V(12, 0), PQH_V(12, 1), PQH_V(12, 2));
#pragma unroll
        for (int c = 3; c < 31; c += 2) {  // columns 3..30 of each 32-column quarter, pairwise
            m0 = S::max3(m0, PQH_V(c >> 3, c & 7), PQH_V((c + 1) >> 3, (c + 1) & 7));
            m1 = S::max3(m1, PQH_V(4 + (c >> 3), c & 7), PQH_V(4 + ((c + 1) >> 3), (c + 1) & 7));
            m2 = S::max3(m2, PQH_V(8 + (c >> 3), c & 7), PQH_V(8 + ((c + 1) >> 3), (c + 1) & 7));
            m3 = S::max3(m3, PQH_V(12 + (c >> 3), c & 7), PQH_V(12 + ((c + 1) >> 3), (c + 1) & 7));
        }
        // maxima of the four 32-column quarters (the slow path only scans quarters that hold
        // something), then of the row
        m0 = S::max2(m0, PQH_V(3, 7));
        m1 = S::max2(m1, PQH_V(7, 7));
        m2 = S::max2(m2, PQH_V(11, 7));
        m3 = S::max2(m3, PQH_V(15, 7));
        const T rm = S::max3(S::max3(m0, m1, m2), m3, S::lowest());
        const bool hot = !__any_sync(0xffffffffu, rm > tau) || debug_mode == 2;
        prof.lap(PROF_HOT);
        if (hot) return;
        ++slow_tiles;
        const T qm[4] = {m0, m1, m2, m3};
#pragma unroll
        for (int Q = 0; Q < 4; ++Q) {
            if (!__any_sync(0xffffffffu, qm[Q] > tau)) continue;
#pragma unroll
            for (int q = Q * 4; q < Q * 4 + 4; ++q) {
                T m = S::max3(PQH_V(q, 0), PQH_V(q, 1), PQH_V(q, 2));
                m = S::max3(m, PQH_V(q, 3), PQH_V(q, 4));
                m = S::max3(m, PQH_V(q, 5), PQH_V(q, 6));
                m = S::max2(m, PQH_V(q, 7));
                if (__any_sync(0xffffffffu, m > tau)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t bits = (q < 8) ? ra[q * 8 + j] : rb[(q - 8) * 8 + j];
                        append_if<I8>(S::from_bits(bits) > tau, list, cnt, bits,
                                      col_base + q * 8 + j);
                    }
                }
            }
        }
    } else {
        T gm[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) gm[q] = S::max3(PQH_V(q, 0), PQH_V(q, 1), PQH_V(q, 2));
#pragma unroll
        for (int q = 0; q < NG; ++q) gm[q] = S::max3(gm[q], PQH_V(q, 3), PQH_V(q, 4));
#pragma unroll
        for (int q = 0; q < NG; ++q) gm[q] = S::max3(gm[q], PQH_V(q, 5), PQH_V(q, 6));
#pragma unroll
        for (int q = 0; q < NG; ++q) gm[q] = S::max2(gm[q], PQH_V(q, 7));
        const T r0 = S::max3(gm[0], gm[1], gm[2]), r1 = S::max3(gm[3], gm[4], gm[5]);
        const T r2 = S::max3(gm[6], gm[7], gm[8]), r3 = S::max3(gm[9], gm[10], gm[11]);
        const T r4 = S::max3(gm[12], gm[13], gm[14]);
        const T rm = S::max3(S::max3(r0, r1, r2), S::max3(r3, r4, gm[15]), S::lowest());
        // common case: no row of the warp has anything above its threshold -> one vote, done
        if (!__any_sync(0xffffffffu, rm > tau) || debug_mode == 2) {
            release();
            return;
        }
        unsigned gmask = 0;
#pragma unroll
        for (int q = 0; q < NG; ++q) gmask |= (gm[q] > tau) ? (1u << q) : 0u;
        unsigned groups = __reduce_or_sync(0xffffffffu, gmask);
        ++slow_tiles;
        // ragged last tile / the row's own column: only "special" tiles carry a real limit
        const int id_lim = special ? ib_hi : 0x7fffffff;
        auto scan8 = [&](const uint32_t(&v8)[8], int q) {
            const int col0 = col_base + q * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int id = col0 + j;
                const bool p = (S::from_bits(v8[j]) > tau) & (id < id_lim) & (id != self);
                append_if<I8>(p, list, cnt, v8[j], id);
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
    }
#undef PQH_V
    prof.lap(PROF_SCAN);
    if (cnt > GCAP) __trap();  // list overflow would silently corrupt a neighbouring row's list
    // hard capacity: a tile can append up to TC_BN entries before this check runs again
    const unsigned full = __ballot_sync(0xffffffffu, cnt > GCAP - TC_BN);
    shrink_flagged<GCAP, I8, false>(full, L, lane, tau, cnt, list);
    prof.add(PROF_N_HARD, __popc(full));
    prof.lap(PROF_HARD);
}

// Wait for the next accumulator tile.  If some list of the warp has grown past 1.5 L and the
// tensor cores are not done yet, the idle time goes into shrinking it: the threshold then stays
// fresh without a shrink ever standing between a finished accumulator stage and its release
// (mandatory shrinks at the hard capacity remain in filter_tile, but become rare).  The common
// case -- nothing to shrink -- goes straight to the suspending wait, which resumes ~60 cycles
// after the barrier flips (a non-blocking probe costs ~150 cycles even when the phase is
// already complete).  Loop conditions are warp-uniform: the body runs full-mask collectives.
template <int GCAP, bool I8, bool PROF>
__device__ __forceinline__ void wait_acc_with_idle_shrinks(uint64_t *bar, uint32_t parity, int L,
                                                           int lane, typename Score<I8>::T &tau,
                                                           int &cnt, key64_t *list,
                                                           EpiProf<PROF> &prof) {
    const int soft = (L + L / 2 < GCAP - TC_BN) ? L + L / 2 : GCAP - TC_BN;
    unsigned need = __ballot_sync(0xffffffffu, cnt > soft);
    while (need != 0 && !__all_sync(0xffffffffu, ptx::mbar_test_wait(bar, parity))) {
        const uint32_t t0 = PROF ? (uint32_t)clock() : 0u;
        shrink_flagged<GCAP, I8, true>(need & (0u - need), L, lane, tau, cnt, list);
        need = __ballot_sync(0xffffffffu, cnt > soft);
        prof.add(PROF_IDLE, (uint32_t)clock() - t0);
        prof.add(PROF_N_IDLE, 1);
    }
    ptx::mbar_wait(bar, parity);
    prof.lap(PROF_WAIT);
}

// Barriers and hand-over words the filter warps use, in the shared-memory tail of both kernels.
struct EpiBars {
    uint64_t acc_full[2], acc_empty[2];
    uint64_t t0_ready;  // producer -> epilogue: the order of the next item is written
    uint32_t tmem_base;
    int item_w0[2], item_uc[2];  // tile order of the current / next work item (ItemOrder)
    int item_id[2];              // ... and which item it is (-1: no more work)
};

// Work items of a CTA.  The producer thread decides: the first item of CTA b is item b; after
// that either every gridDim-th item (round robin) or -- args.dynamic -- the next one nobody has
// taken yet, from a counter in global memory.  Items differ in how long they take (slow tiles,
// shrinks), and with a fixed share of ~27 items per CTA the slowest CTA of a 1M-row search ran
// 7 % longer than the average one; handing the items out on demand ends all CTAs within one
// (short, last) item of each other.  The producer asks for the following item when it starts
// one, so the atomic's latency is hidden behind a whole sweep; everyone else learns the item
// from eb->item_id when the t0_ready barrier flips.
__device__ __forceinline__ int next_item(const TcArgs &a, int item) {
    if (!a.dynamic) return item + (int)gridDim.x;
    return (int)gridDim.x +
           (int)atomicAdd(&const_cast<PrepScalars *>(a.sc)->work_next, 1u);
}
// consumer side: the item of this round, -1 when the CTA is done
__device__ __forceinline__ int take_item(EpiBars *eb, int it_par) {
    ptx::mbar_wait(&eb->t0_ready, (uint32_t)it_par);
    return *reinterpret_cast<volatile int *>(&eb->item_id[it_par]);
}

// The filter warps of one CTA: filter warp e (0 .. 7) owns TMEM lane quadrant `quad` (= its
// warp index & 3: the quadrant the hardware lets it read; four consecutive warps cover the
// four quadrants whatever the index of the first filter warp is) of query tile g = e >> 2,
// i.e. 32 query rows, one per lane, and filters every base tile of the sweep for them.
// Steps are numbered over the whole life of the CTA (gstep): accumulator stage = step & 1,
// barrier phase = (step >> 1) & 1.
template <int GCAP, bool I8, bool EARLY, bool PROF = false>
__device__ __forceinline__ void epilogue_warps(const TcArgs &args, int e, int quad, int lane,
                                               EpiBars *eb) {
    typedef Score<I8> S;
    typedef typename S::T T;
    const int g = e >> 2;
    const int row_in_cta = g * 128 + quad * 32 + lane;
    const uint32_t tmem_row = eb->tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 128);
    const int L = args.sc->inexact ? args.L_inexact : args.L_exact;
    const int n_items = total_items(args);
    uint32_t gstep = 0;
    int it_par = 0;
    unsigned slow_tiles = 0;
    EpiProf<PROF> prof;
    prof.init();
    for (;;) {
        const int item = take_item(eb, it_par);
        if (item < 0) break;
        const ItemInfo it = decode_item(args, item);
        const int qt = it.qt, sp = it.sp, t_lo = it.t_lo, t_hi = it.t_hi;
        const int r = qt * TC_ROWS + row_in_cta;
        const bool valid = r < args.n_rows;
        int self = -1;
        if (valid && args.self_join)
            self = args.q_self ? args.q_self[r] : (args.row_ids ? args.row_ids[r] : args.q_begin + r);
        const size_t slot =
            valid ? list_slot(r, sp, args.q_tail0 * TC_ROWS, args.nsplit, args.ns_tail) : 0;
        key64_t *list = args.cand + slot * GCAP;
        T tau = valid ? S::lowest() : S::highest();
        int cnt = 0;
        long long b_hi = (long long)t_hi * TC_BN;
        if (b_hi > args.nb) b_hi = args.nb;
        const int ib_hi = (int)b_hi;

        const int nst = t_hi - t_lo;
        const int wlen = own_window_len(args, qt, t_lo, t_hi);
        // the producer wrote the tile order of this item before it loaded anything
        ItemOrder ord;
        ord.w0 = *reinterpret_cast<volatile int *>(&eb->item_w0[it_par]);
        ord.uc = *reinterpret_cast<volatile int *>(&eb->item_uc[it_par]);
        for (int i = 0; i < nst; ++i) {
            const uint32_t s = gstep + (uint32_t)i;
            const int as = (int)(s & 1u);
            const uint32_t acc_ph = (s >> 1) & 1u;
            // everything that does not need the accumulator is computed BEFORE the wait: what
            // follows the barrier flip is on the critical path of the accumulator ring
            int tile0 = item_tile(i, ord, wlen, t_lo, t_hi) * TC_BN;
            uint32_t tbase = tmem_row + (uint32_t)(as * 256);
            // id checks are only needed in the tile holding this row's own column and in
            // the (possibly ragged) last tile
            int special = (tile0 == (t_hi - 1) * TC_BN) || (self >= tile0 && self < tile0 + TC_BN);
            asm volatile("" : "+r"(tile0), "+r"(tbase), "+r"(special));
            wait_acc_with_idle_shrinks<GCAP, I8, PROF>(&eb->acc_full[as], acc_ph, L, lane, tau, cnt,
                                                       list, prof);
            ptx::tc_fence_after_sync();
            auto release = [&]() {
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&eb->acc_empty[as]);
            };
            if (args.debug_mode != 1)
                filter_tile<GCAP, I8, EARLY, PROF>(tbase, tile0, special != 0, ib_hi, self, L, lane,
                                                   args.debug_mode, tau, cnt, list, slow_tiles,
                                                   prof, release);
            else
                release();
        }
        gstep += (uint32_t)nst;
        it_par ^= 1;
        if (valid) {
            args.cand_cnt[slot] = cnt;
            args.cand_tau[slot] = S::as_float(tau);
        }
    }
    if (lane == 0 && slow_tiles)
        atomicAdd(const_cast<unsigned long long *>(&args.sc->dbg[2]), (unsigned long long)slow_tiles);
    if constexpr (PROF) {
        if (lane == 0)
            for (int i = 0; i < 8; ++i)
                atomicAdd(const_cast<unsigned long long *>(&args.sc->dbg[3 + i]),
                          (unsigned long long)prof.t[i]);
    }
}


// Arg-max flavour of the filter warps: the nearest pivot of every row (knn_order.cu).  The
// "base" is the pivot set, a sweep is a few dozen tiles, and all that is wanted per row is the
// column with the best score -- no candidate lists, no thresholds to certify.  Same geometry and
// barriers as epilogue_warps; the accumulator stage is released as soon as the tile is in
// registers.  Ties keep the lower column.  One work item per query tile (no base splits).
template <bool I8>
__device__ __forceinline__ void epilogue_warps_argmax(const TcArgs &args, int e, int quad, int lane,
                                                      EpiBars *eb) {
    typedef Score<I8> S;
    typedef typename S::T T;
    const int g = e >> 2;
    const int row_in_cta = g * 128 + quad * 32 + lane;
    const uint32_t tmem_row = eb->tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 128);
    const int n_items = total_items(args);
    uint32_t gstep = 0;
    int it_par = 0;
    for (;;) {
        const int item = take_item(eb, it_par);
        if (item < 0) break;
        const ItemInfo it = decode_item(args, item);
        const int t_lo = it.t_lo, t_hi = it.t_hi;
        const int r = it.qt * TC_ROWS + row_in_cta;
        T best = S::lowest();
        int best_col = 0;
        const int nst = t_hi - t_lo;
        const int wlen = own_window_len(args, it.qt, t_lo, t_hi);
        ItemOrder ord;
        ord.w0 = *reinterpret_cast<volatile int *>(&eb->item_w0[it_par]);
        ord.uc = *reinterpret_cast<volatile int *>(&eb->item_uc[it_par]);
        for (int i = 0; i < nst; ++i) {
            const uint32_t s = gstep + (uint32_t)i;
            const int as = (int)(s & 1u);
            const int col_base = item_tile(i, ord, wlen, t_lo, t_hi) * TC_BN;
            const uint32_t tbase = tmem_row + (uint32_t)(as * 256);
            ptx::mbar_wait(&eb->acc_full[as], (s >> 1) & 1u);
            ptx::tc_fence_after_sync();
            uint32_t ra[64], rb[64];
            ptx::tmem_ld_32x32b_x64(tbase, ra);
            ptx::tmem_ld_32x32b_x64(tbase + 64, rb);
            ptx::tmem_wait_ld();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&eb->acc_empty[as]);
#define PQH_V(q, j) S::from_bits(((q) < 8) ? ra[(q) * 8 + (j)] : rb[((q) - 8) * 8 + (j)])
#pragma unroll
            for (int q = 0; q < TC_BN / 8; ++q) {
                T m = S::max3(PQH_V(q, 0), PQH_V(q, 1), PQH_V(q, 2));
                m = S::max3(m, PQH_V(q, 3), PQH_V(q, 4));
                m = S::max3(m, PQH_V(q, 5), PQH_V(q, 6));
                m = S::max2(m, PQH_V(q, 7));
                if (__any_sync(0xffffffffu, m > best)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const T v = PQH_V(q, j);
                        const bool p = v > best && col_base + q * 8 + j < args.nb;
                        best = p ? v : best;
                        best_col = p ? col_base + q * 8 + j : best_col;
                    }
                }
            }
#undef PQH_V
        }
        gstep += (uint32_t)nst;
        it_par ^= 1;
        if (r < args.n_rows) {
            args.pick_cell[r] = best_col;
            atomicAdd(args.pick_hist + best_col, 1);
        }
    }
}

}  // namespace
}  // namespace pqh
