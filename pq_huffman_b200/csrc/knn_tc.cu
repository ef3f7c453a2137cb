// knn_tc.cu -- tensor-core kNN: tcgen05/TMEM GEMM with a fused per-row candidate
// filter, followed by an exact FP32 re-rank with a completeness certificate.
//
// Score.  For query x and base y the squared L2 distance is |x|^2 - 2 S with
// S = x.y - |y|^2/2, so "k smallest distances" == "k largest S" per row.  The
// kernel computes S' = sigma^2 * S on the tensor cores from FP16 operands:
//   * prep: x' = fp16(sigma * x) with sigma a power of two putting max|x| in
//     [2^7, 2^8) (exact for integer-valued data such as SIFT-shaped uint8), and
//     |y'|^2/2 split into three FP16 pieces q1,q2,q3 with |y'|^2/2 =
//     4096 q1 + 2 q2 + 2^-10 q3 (+ <2^-30 relative);
//   * the GEMM K dimension is [x'(d16) | fold(16)]: the A (query) side holds the
//     constants (4096, 2, 2^-10, 0..), the B (base) side (-q1,-q2,-q3, 0..), so
//     the accumulator IS S' and the epilogue is a bare threshold compare.
// Data made of integers in [0, 255] takes the integer flavour instead (u8 x s8 -> s32 on
// tcgen05.mma kind::i8, exact scores; knn_tc_common.cuh).
// Nothing of the nq x nb score matrix reaches HBM: an epilogue thread owns one
// query row, copies its 128 accumulator columns from TMEM to registers
// (tcgen05.ld), reduces them to group maxima and only touches memory for the
// rare elements above the row's running threshold (filter_tile in
// knn_tc_common.cuh).
//
// Exactness.  Candidates (a superset of the top-L by S') are re-ranked by
// knn_rerank_kernel with the scalar FP32 sequence of get_real_dist
// (compute_nn_fast.c:304-311), giving distances bit-identical to the oracle.
// A row is accepted only if a proven bound shows no non-candidate can beat its
// k-th exact distance (see certificate in knn_rerank_kernel); other rows get a
// second pass with longer candidate lists and finally the exact kernel
// (knn_exact.cu).  So the result is always the oracle's, and the tensor cores
// only decide how fast it is found.
//
// Kernels in this file: the prep kernels, knn_tc_kernel (d <= ~144: query tile resident in
// shared memory), the re-rank kernel, and the host-side orchestration launch_knn_tc.
// knn_tcs.cu holds the streaming kernel for larger d, knn_order.cu the locality order.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include "knn_tc_common.cuh"

namespace pqh {

namespace {

// ------------------------------------------------------------------------------------
// environment knobs, read once per context (tc_reload_env)
struct TcEnv {
    int i8 = 1;        // PQH_TC_I8: integer tensor-core path for uint8-valued data
    int order = 1;     // PQH_TC_ORDER: locality order of the sweep (knn_order.cu)
    int nmma = 2;      // PQH_TC_MMA: MMA-issuing threads per CTA, 1 or 2
    int early = 1;     // PQH_TC_EARLY: release the accumulator stage before filtering (0: after)
    int own_tiles = 8; // PQH_TC_OWN: least number of tiles in the window a sweep starts with
    int l_exact = 0, l_inexact = 0;  // PQH_TC_LEXACT / PQH_TC_LINEXACT: candidates kept per row
    int gcap = 0;      // PQH_TC_GCAP: list capacity, 256 or 512
    int nsplit = 0;    // PQH_TC_NSPLIT: base splits
    int no_tail = 0, no_pass2 = 0, check_eps = 0;
    int debug = 0;     // PQH_TC_DEBUG: 1 = MMA only, 2 = no slow path (timing experiments),
                       // 3 = normal run + counters printed
    int streaming = 0; // PQH_TC_STREAM=1: streaming kernel for every d
    int dynamic = 1;   // PQH_TC_DYN=0: work items round robin instead of on demand
    int wide_d = 64;   // PQH_TC_WIDE_D: re-rank rows of at least this many floats are fetched by
                       // the whole warp and handed to their lanes through shared memory
                       // (measured: d = 128 6.3 -> 4.9 ms at 1M rows, d = 96 k = 32 38 -> 25 ms at 2M)
    int slice_mb = 2048;  // PQH_TC_SLICE_MB: candidate-list scratch per row slice
    bool loaded = false;
};
TcEnv g_env;

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
const TcEnv &env() {
    if (!g_env.loaded) tc_reload_env();
    return g_env;
}

// ------------------------------------------------------------------------------------
// prep 1: max |x| over a matrix (grid-stride, float4), atomicMax on the IEEE bits
__global__ void __launch_bounds__(256) maxabs_kernel(const float *__restrict__ x, size_t n,
                                                     uint32_t *__restrict__ out_bits) {
    float m = 0.0f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)x) & 15) == 0) {
        const size_t n4 = n / 4;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        for (size_t j = i; j < n4; j += stride) {
            float4 v = __ldg(x4 + j);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        for (size_t j = n4 * 4 + i; j < n; j += stride) m = fmaxf(m, fabsf(x[j]));
    } else {
        for (size_t j = i; j < n; j += stride) m = fmaxf(m, fabsf(x[j]));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out_bits, __float_as_uint(m));
}

// prep 2: one warp per operand row: fp16 operand row, exact-ish FP64 squared norm, fold pieces.
// Operand row `row` is built from data row perm[row] - perm_off (perm null: row itself); the
// norm is stored under the DATA row.
__global__ void __launch_bounds__(256)
prep_rows_kernel(const float *__restrict__ x, int n, int d, int d16, __half *__restrict__ xh,
                 double *__restrict__ norm64, __half *__restrict__ fold, int write_fold,
                 const int *__restrict__ perm, int perm_off, PrepScalars *__restrict__ sc) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int src = perm ? perm[row] - perm_off : row;
    const float sigma = scale_from_maxabs_bits(sc->maxabs_bits);
    const float *xr = x + (size_t)src * d;
    __half *hr = xh + (size_t)row * d16;
    double acc = 0.0;
    bool inexact = false;
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        // four consecutive elements per lane: 16-byte loads, 8-byte stores (d16 is a multiple of 16)
        for (int i4 = lane; i4 < d16 / 4; i4 += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * i4 < d) v = __ldg(reinterpret_cast<const float4 *>(xr) + i4);
            const float s0 = v.x * sigma, s1 = v.y * sigma, s2 = v.z * sigma, s3 = v.w * sigma;
            const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
            if (__low2float(h01) != s0 || __high2float(h01) != s1 || __low2float(h23) != s2 ||
                __high2float(h23) != s3)
                inexact = true;
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t *>(&h01);
            pk.y = *reinterpret_cast<const uint32_t *>(&h23);
            reinterpret_cast<uint2 *>(hr)[i4] = pk;
            acc += (double)v.x * (double)v.x;
            acc += (double)v.y * (double)v.y;
            acc += (double)v.z * (double)v.z;
            acc += (double)v.w * (double)v.w;
        }
    } else {
        for (int i = lane; i < d16; i += 32) {
            float v = i < d ? xr[i] : 0.0f;
            float s = v * sigma;  // exact: power-of-two scale
            __half h = __float2half_rn(s);
            if (__half2float(h) != s) inexact = true;
            hr[i] = h;
            acc += (double)v * (double)v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (__any_sync(0xffffffffu, inexact) && lane == 0) atomicOr(&sc->inexact, 1u);
    if (lane == 0 && norm64) norm64[src] = acc;
    if (write_fold) {
        // n' = sigma^2 |y|^2 / 2 = 4096 q1 + 2 q2 + 2^-10 q3 (+ residual)
        const double np = acc * (double)sigma * (double)sigma * 0.5;
        __half q1 = __double2half(np / 4096.0);
        double r1 = np - 4096.0 * (double)__half2float(q1);
        __half q2 = __double2half(r1 * 0.5);
        double r2 = r1 - 2.0 * (double)__half2float(q2);
        __half q3 = __double2half(r2 * 1024.0);
        if (lane < FOLD_COLS) {
            __half v = __float2half_rn(0.0f);
            if (lane == 0) v = __hneg(q1);
            if (lane == 1) v = __hneg(q2);
            if (lane == 2) v = __hneg(q3);
            fold[(size_t)row * FOLD_COLS + lane] = v;
        }
        if (lane == 0) {
            // scaled norm, rounded up so that it is a valid upper bound
            float yn = __double2float_ru(sqrt(acc) * (double)sigma * (1.0 + 1e-7));
            atomicMax(&sc->ymax_bits, __float_as_uint(yn));
        }
    }
}

// second pass: collect the query vectors of the rows that failed the first certificate
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float *__restrict__ src, int src_row0, const int *__restrict__ ids, int n,
                   int d, float *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *s = src + (size_t)(ids[row] - src_row0) * d;
    float *o = dst + (size_t)row * d;
    for (int i = lane; i < d; i += 32) o[i] = s[i];
}

__global__ void fold_const_kernel(__half *__restrict__ fa) {
    // [256 x 16]: every query row multiplies the base pieces by (4096, 2, 2^-10)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_ROWS * FOLD_COLS) return;
    const int c = i % FOLD_COLS;
    float v = c == 0 ? 4096.0f : (c == 1 ? 2.0f : (c == 2 ? 0.0009765625f : 0.0f));
    fa[i] = __float2half_rn(v);
}

// Integer path, one warp per operand row: the row as 8-bit integers -- u8 x for the A (query)
// operand, s8 x - 128 for the B (base) operand -- zero-padded to d32 bytes; FP64 squared norm
// (stored under the data row); for base rows the 32 s8 fold digits of G = ceil(|y|^2 / 2).
// sc->i8_reject collects the reasons the integer path cannot be used.  out == fold == null: only
// check.
__global__ void __launch_bounds__(256)
prep_rows_i8_kernel(const float *__restrict__ x, int n, int d, int d32, uint8_t *__restrict__ out,
                    double *__restrict__ norm64, int8_t *__restrict__ fold, int as_base,
                    const int *__restrict__ perm, int perm_off, PrepScalars *__restrict__ sc) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int src = perm ? perm[row] - perm_off : row;
    const float *xr = x + (size_t)src * d;
    uint8_t *o = out ? out + (size_t)row * d32 : nullptr;
    double acc = 0.0;
    bool bad = false;
    if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        // four consecutive elements per lane: 16-byte loads, 4-byte stores (d32 is a multiple of 32)
        const int sub = as_base ? 128 : 0;
        for (int i4 = lane; i4 < d32 / 4; i4 += 32) {
            const bool in = 4 * i4 < d;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) v = __ldg(reinterpret_cast<const float4 *>(xr) + i4);
            const int i0 = (int)v.x, i1 = (int)v.y, i2 = (int)v.z, i3 = (int)v.w;
            if (!(v.x >= 0.0f && v.x <= 255.0f) || (float)i0 != v.x) bad = true;
            if (!(v.y >= 0.0f && v.y <= 255.0f) || (float)i1 != v.y) bad = true;
            if (!(v.z >= 0.0f && v.z <= 255.0f) || (float)i2 != v.z) bad = true;
            if (!(v.w >= 0.0f && v.w <= 255.0f) || (float)i3 != v.w) bad = true;
            uint32_t pk = 0;
            if (in)
                pk = ((uint32_t)(i0 - sub) & 0xffu) | (((uint32_t)(i1 - sub) & 0xffu) << 8) |
                     (((uint32_t)(i2 - sub) & 0xffu) << 16) | (((uint32_t)(i3 - sub) & 0xffu) << 24);
            if (o) reinterpret_cast<uint32_t *>(o)[i4] = pk;
            acc += (double)v.x * (double)v.x;
            acc += (double)v.y * (double)v.y;
            acc += (double)v.z * (double)v.z;
            acc += (double)v.w * (double)v.w;
        }
    } else {
        for (int i = lane; i < d32; i += 32) {
            const float v = i < d ? xr[i] : 0.0f;
            const int iv = (int)v;
            if (!(v >= 0.0f && v <= 255.0f) || (float)iv != v) bad = true;
            uint8_t b = 0;
            if (i < d) b = as_base ? (uint8_t)(int8_t)(iv - 128) : (uint8_t)iv;
            if (o) o[i] = b;
            acc += (double)v * (double)v;
        }
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
    unsigned flags = __any_sync(0xffffffffu, bad) ? 1u : 0u;
    if (norm64 && lane == 0) norm64[src] = acc;
    if (as_base) {
        const long long G = ((long long)acc + 1) / 2;
        const long long q = G / I8_FOLD_UNIT;
        const int r = (int)(G % I8_FOLD_UNIT);
        if (q >= I8_FOLD_FULL) flags |= 2u;
        int b = 0;
        if (lane < I8_FOLD_FULL) {
            if (lane < q) b = -128;
            else if (lane == q) b = -(r / 255);
        } else {
            const int r2 = r % 255, first = r2 < 127 ? r2 : 127;
            b = lane == I8_FOLD_FULL ? -first : -(r2 - first);
        }
        if (fold) fold[(size_t)row * 32 + lane] = (int8_t)b;
    }
    if (flags && lane == 0) atomicOr(&sc->i8_reject, flags);
}

__global__ void fold_const_i8_kernel(uint8_t *__restrict__ fa) {
    // [256 x 32]: what every query row multiplies the fold digits by
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_ROWS * 32) return;
    fa[i] = (i % 32) < I8_FOLD_FULL ? 255 : 1;
}

// ------------------------------------------------------------------------------------
// main kernel
struct TcSmemTail {
    uint64_t a_full, a_empty;
    uint64_t b_full[TC_NSTAGE], b_empty[TC_NSTAGE];
    EpiBars eb;
};

// both barriers of an MMA step probed together: their latencies overlap instead of adding up
__device__ __forceinline__ void mbar_wait2(uint64_t *b0, uint32_t p0, uint64_t *b1, uint32_t p1) {
    const bool ok0 = ptx::mbar_try_wait(b0, p0);
    const bool ok1 = ptx::mbar_try_wait(b1, p1);
    if (!ok0) ptx::mbar_wait_slow(b0, p0);
    if (!ok1) ptx::mbar_wait_slow(b1, p1);
}

// Warp roles: warp 0 = TMA producer, warps 1 .. NMMA = MMA issuers (one elected thread each),
// then the 8 filter warps.  With NMMA == 2 the two issuers take alternate steps (issuer m fills
// accumulator stage m), so the barrier round trips of one overlap the MMAs of the other.
// Steps are numbered over the life of the CTA: B ring stage = step % 4, accumulator stage =
// step & 1.
// MODE: 0 = filter, stage released after the filtering; 1 = filter, stage released early;
// 2 = arg-max (cell assignment of the locality order); 4 = mode 1 with cycle accounting in the
// filter warps (diagnostic builds, -DPQH_TC_PROF).
template <int GCAP, bool I8, int NMMA, int MODE>
__global__ void __launch_bounds__(tc_threads(NMMA), 1)
knn_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcPlan &plan = args.plan;
    uint8_t *smA = smem;
    uint8_t *smB = smA + plan.a_bytes;
    TcSmemTail &tail = *reinterpret_cast<TcSmemTail *>(smB + TC_NSTAGE * plan.b_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int EPI0 = 1 + NMMA;  // first filter warp

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 3; ++i) {
            ptx::prefetch_tmap(&maps.a[i]);
            ptx::prefetch_tmap(&maps.b[i]);
        }
        ptx::prefetch_tmap(&maps.fa);
        ptx::prefetch_tmap(&maps.fb);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(&tail.a_full, 1);
        ptx::mbar_init(&tail.a_empty, NMMA);
        for (int i = 0; i < TC_NSTAGE; ++i) {
            ptx::mbar_init(&tail.b_full[i], 1);
            ptx::mbar_init(&tail.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tail.eb.acc_full[i], 1);
            ptx::mbar_init(&tail.eb.acc_empty[i], TC_EPI_WARPS);
        }
        ptx::mbar_init(&tail.eb.t0_ready, 1);
        ptx::fence_barrier_init();
    }
    if (warp == EPI0) ptx::tmem_alloc<512>(&tail.eb.tmem_base);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = tail.eb.tmem_base;

    const int n_items = total_items(args);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (ptx::elect_one()) {
            int stage = 0, it_par = 0, my_next = 0;
            uint32_t ph = 0, a_ph = 0;
            for (int item = blockIdx.x;;) {
                // the MMAs of the previous item are done: the query tile may be overwritten and
                // every consumer has seen the previous flip of t0_ready
                ptx::mbar_wait(&tail.a_empty, a_ph ^ 1);
                if (item >= n_items) {
                    tail.eb.item_id[it_par] = -1;
                    ptx::mbar_arrive(&tail.eb.t0_ready);
                    break;
                }
                const int following = next_item(args, item);
                const ItemInfo it = decode_item(args, item);
                const int qt = it.qt, t_lo = it.t_lo, t_hi = it.t_hi;
                const ItemOrder ord = item_order(args, qt, t_lo, t_hi, my_next);
                const int wlen = own_window_len(args, qt, t_lo, t_hi);
                tail.eb.item_w0[it_par] = ord.w0;
                tail.eb.item_uc[it_par] = ord.uc;
                tail.eb.item_id[it_par] = item;
                ptx::mbar_arrive(&tail.eb.t0_ready);
                ptx::mbar_arrive_expect_tx(&tail.a_full, plan.a_bytes);
                for (int s = 0; s < plan.nseg; ++s) {
                    const TcSeg &sg = plan.seg[s];
                    if (sg.fold)
                        ptx::tma_load_2d(smA + sg.a_off, &maps.fa, &tail.a_full, 0, 0);
                    else
                        ptx::tma_load_2d(smA + sg.a_off, &maps.a[seg_class(sg.bytes)],
                                         &tail.a_full, sg.k_off, args.q_op_begin + qt * TC_ROWS);
                }
                a_ph ^= 1;
                it_par ^= 1;
                for (int i = 0; i < t_hi - t_lo; ++i) {
                    const int t = item_tile(i, ord, wlen, t_lo, t_hi);
                    if (i == wlen) my_next = t;  // the cyclic part ends where it started
                    if (i >= wlen && (i & 7) == 0) publish_cursor(args, t);
                    ptx::mbar_wait(&tail.b_empty[stage], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&tail.b_full[stage], plan.b_bytes);
                    uint8_t *dst = smB + stage * plan.b_bytes;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        if (sg.fold)
                            ptx::tma_load_2d(dst + sg.b_off, &maps.fb, &tail.b_full[stage], 0,
                                             t * TC_BN);
                        else
                            ptx::tma_load_2d(dst + sg.b_off, &maps.b[seg_class(sg.bytes)],
                                             &tail.b_full[stage], sg.k_off, t * TC_BN);
                    }
                    if (++stage == TC_NSTAGE) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
                item = following;
            }
        }
    } else if (warp < EPI0) {
        // ===================== MMA issuer(s) =====================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc =
                I8 ? ptx::make_idesc_i8_u8s8(128, TC_BN) : ptx::make_idesc_f16(128, TC_BN);
            const uint32_t m = (uint32_t)(warp - 1);
            uint32_t gstep = 0, a_ph = 0;
            long long t_wait = 0;
            const bool prof = args.debug_mode != 0;
            const long long t_begin = prof ? clock64() : 0;
            for (int it_par = 0;; it_par ^= 1) {
                const int item = take_item(&tail.eb, it_par);
                if (item < 0) break;
                const ItemInfo it = decode_item(args, item);
                const int nst = it.t_hi - it.t_lo;
                ptx::mbar_wait(&tail.a_full, a_ph);
                a_ph ^= 1;
                bool issued = false;
                for (int i = (NMMA == 2) ? (int)((m ^ gstep) & 1u) : 0; i < nst; i += NMMA) {
                    const uint32_t s = gstep + (uint32_t)i;
                    const int stage = (int)(s % TC_NSTAGE), as = (int)(s & 1u);
                    const uint32_t ph = (s / TC_NSTAGE) & 1u, acc_ph = (s >> 1) & 1u;
                    const long long w0 = prof ? clock64() : 0;
                    mbar_wait2(&tail.b_full[stage], ph, &tail.eb.acc_empty[as], acc_ph ^ 1);
                    if (prof) t_wait += clock64() - w0;
                    ptx::tc_fence_after_sync();
                    const uint32_t b_base = ptx::smem_u32(smB + stage * plan.b_bytes);
                    const uint32_t a_base = ptx::smem_u32(smA);
                    // The two query tiles are independent accumulation chains: alternate
                    // between them so that consecutive MMAs never depend on each other.
                    uint32_t acc = 0;
                    for (int sgi = 0; sgi < plan.nseg; ++sgi) {
                        const TcSeg &sg = plan.seg[sgi];
                        const uint32_t b_addr = b_base + sg.b_off;
                        const uint32_t sbo = 8 * sg.bytes, lay = seg_layout(sg.bytes);
                        for (int kk = 0; kk < sg.bytes / 32; ++kk) {
                            const uint64_t bd = ptx::make_smem_desc(b_addr + kk * 32, sbo, lay);
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                const uint32_t a_addr = a_base + sg.a_off + g * 128 * sg.bytes;
                                const uint64_t ad = ptx::make_smem_desc(a_addr + kk * 32, sbo, lay);
                                const uint32_t dcol = tmem_base + (uint32_t)(as * 256 + g * 128);
                                if constexpr (I8) ptx::umma_i8(dcol, ad, bd, idesc, acc);
                                else ptx::umma_f16(dcol, ad, bd, idesc, acc);
                            }
                            acc = 1;
                        }
                    }
                    ptx::umma_commit(&tail.b_empty[stage]);
                    ptx::umma_commit(&tail.eb.acc_full[as]);
                    issued = true;
                }
                // the query tile may be overwritten once the MMAs of every issuer are done
                if (issued) ptx::umma_commit(&tail.a_empty);
                else ptx::mbar_arrive(&tail.a_empty);
                gstep += (uint32_t)nst;
            }
            if (prof) {
                PrepScalars *scw = const_cast<PrepScalars *>(args.sc);
                atomicAdd(&scw->dbg[0], (unsigned long long)t_wait);
                atomicAdd(&scw->dbg[1], (unsigned long long)(clock64() - t_begin));
            }
        }
    } else {
        // ===================== epilogue: threshold filter =====================
        if constexpr (MODE == 2)
            epilogue_warps_argmax<I8>(args, warp - EPI0, warp & 3, lane, &tail.eb);
        else
            epilogue_warps<GCAP, I8, MODE == 1 || MODE == 4, MODE == 4>(args, warp - EPI0, warp & 3,
                                                                        lane, &tail.eb);
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == EPI0) ptx::tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------
// exact FP32 re-rank + certificate.  One warp per query row.
constexpr int RR_WARPS = 4;

struct RerankArgs {
    const float *base;
    const float *queries;  // FP32 rows of this pass's queries: row ql (below) is the query's vector
    int q_begin;           // absolute query index of row 0 when rows are in natural order
    const int *row_ids;    // second pass: absolute query index of every row instead
    const int *q_perm;     // locality order: absolute query index of every row; the query's vector
                           // is then row (q - q_begin) of `queries`
    const int *base_perm;  // locality order: candidate position -> base row (null: identity);
                           // applied while the candidates are gathered
    int out_q_begin;       // output row = q - out_q_begin
    int n_rows, nb, d, k, kcap, self_join;
    int nlists, gcap;        // candidate lists per row (base splits), capacity of one
    int tail_row0, ns_tail;  // rows >= tail_row0 own ns_tail splits (knn_tc_common.cuh list_slot)
    const key64_t *cand;
    const int *cand_cnt;
    const float *cand_tau;
    const double *qnorm64;  // |x|^2 of the query rows, indexed like `queries`
    const double *bnorm64;  // |y|^2 of the base rows (by base row)
    int check_eps;          // debug: measure |S'_tensor - S'_fp64| / eps over all candidates
    const PrepScalars *sc;
    int ksteps;             // UMMA K steps per accumulator (for the accumulation bound)
    int i8;                 // scores came from the integer path
    int *assign;
    float *dis;
    int *fb_rows;           // absolute query ids of rows that failed the certificate
    uint32_t *fb_count;
};

// WIDE (d >= 64, rows 16-byte aligned): a lane still owns one candidate and sums its d terms in
// index order, but the candidate rows are FETCHED by the whole warp -- 256 contiguous bytes of
// two rows per load instruction -- and handed to their lanes through a shared-memory tile.  With
// one scattered 16-byte load per lane (the narrow path) the re-rank is bound by the rate of
// sector requests (32 different rows per instruction), not by bytes: d = 960 213 -> 59 ms,
// d = 128 6.3 -> 4.9 ms at 1M rows.
constexpr int RR_DC = 64;        // floats of a row per tile chunk
constexpr int RR_TSTR = RR_DC + 4;  // tile row stride: 16-byte aligned, conflict-free LDS.128

template <int NMAX, bool WIDE>
__global__ void __launch_bounds__(RR_WARPS * 32) knn_rerank_kernel(const RerankArgs a) {
    extern __shared__ __align__(16) unsigned char rr_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    key64_t *keys = reinterpret_cast<key64_t *>(rr_smem) + (size_t)warp * NMAX;
    float *qrow = reinterpret_cast<float *>(rr_smem + (size_t)RR_WARPS * NMAX * sizeof(key64_t)) +
                  (size_t)warp * a.d;
    float *tile = reinterpret_cast<float *>(rr_smem + (size_t)RR_WARPS * NMAX * sizeof(key64_t)) +
                  (size_t)RR_WARPS * a.d + (size_t)warp * 32 * RR_TSTR;  // WIDE only
    const int r = blockIdx.x * RR_WARPS + warp;
    if (r >= a.n_rows) return;
    const int q = a.q_perm ? a.q_perm[r] : (a.row_ids ? a.row_ids[r] : a.q_begin + r);
    const int ql = a.q_perm ? q - a.q_begin : r;  // row of `queries` / `qnorm64`

    // the row's threshold: the largest over its lists (each is a statement about the whole row)
    const int my_lists = r < a.tail_row0 ? a.nlists : a.ns_tail;
    const size_t slot0 = list_slot(r, 0, a.tail_row0, a.nlists, a.ns_tail);
    float tau = __int_as_float(0xff800000);
    for (int s = 0; s < my_lists; ++s) tau = fmaxf(tau, a.cand_tau[slot0 + s]);
    // candidates scoring below it cannot be among the row's best L: drop them unseen
    uint32_t hi_max = 0xffffffffu;
    if (tau > __int_as_float(0xff800000))
        hi_max = a.i8 ? i32_score_to_hi((int)tau) : float_to_ordered(-tau);
    int n = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int s = 0; s < my_lists; ++s) {
        const int c = a.cand_cnt[slot0 + s];
        const key64_t *l = a.cand + (slot0 + s) * a.gcap;
        for (int i0 = 0; i0 < c; i0 += 32) {
            const int i = i0 + lane;
            key64_t kv = i < c ? __ldcg(l + i) : 0;
            const bool keep = i < c && (uint32_t)(kv >> 32) <= hi_max;
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                // locality order: the position in the permuted base operand -> the base row
                if (a.base_perm)
                    kv = (kv & 0xFFFFFFFF00000000ull) | (uint32_t)a.base_perm[(uint32_t)kv];
                keys[n + __popc(kb & lt_mask)] = kv;
            }
            n += __popc(kb);
        }
    }
    const float *qsrc = a.queries + (size_t)ql * a.d;
    for (int i = lane; i < a.d; i += 32) qrow[i] = qsrc[i];
    __syncwarp();

    // error bound of the tensor-core score for this row (see the certificate below)
    double eps = 0.0;
    const double sigma = (double)scale_from_maxabs_bits(a.sc->maxabs_bits);
    const double xn2 = a.qnorm64[ql];
    {
        const double ymax = (double)__uint_as_float(a.sc->ymax_bits);
        const double xn = sqrt(xn2) * sigma * (1.0 + 1e-12);
        const double u = 1.0 / 2048.0;  // fp16 unit roundoff 2^-11
        if (a.sc->inexact)
            eps += (2.0 * u + u * u) * xn * ymax + (double)a.d * (1.0 / 32768.0);
        eps += (double)a.ksteps * (1.0 / 4194304.0) * (xn * ymax + 0.5 * ymax * ymax);
        eps += ymax * ymax * (1.0 / 536870912.0);
    }
    if (a.check_eps && !a.i8) {
        // debug: the bound must dominate the real error of every score the kernel produced
        float worst = 0.0f;
        for (int i = lane; i < n; i += 32) {
            const uint32_t id = key_pos(keys[i]);
            const float s_tc = -key_score(keys[i]);
            const float *brow = a.base + (size_t)id * a.d;
            double dot = 0.0;
            for (int j = 0; j < a.d; ++j) dot += (double)qrow[j] * (double)brow[j];
            const double s_true = sigma * sigma * (dot - 0.5 * a.bnorm64[id]);
            const double err = fabs((double)s_tc - s_true);
            const float ratio = eps > 0.0 ? (float)(err / eps) : (err > 0.0 ? 1e30f : 0.0f);
            worst = fmaxf(worst, ratio);
        }
        for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        if (lane == 0)
            atomicMax(const_cast<uint32_t *>(&a.sc->eps_ratio_bits), __float_as_uint(worst));
    }

    // exact distances: the scalar sequence of get_real_dist, one candidate per lane; keys are
    // rebuilt on (distance, base row) so that ties order by the caller's row numbers
    if constexpr (WIDE) {
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int nbatch = n - b0 < 32 ? n - b0 : 32;
            const uint32_t my_id = lane < nbatch ? key_pos(keys[b0 + lane]) : 0u;
            float acc = 0.0f;
            for (int j0 = 0; j0 < a.d; j0 += RR_DC) {
                float4 stage[16];
                const int off = j0 + (lane & 15) * 4;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = 2 * i + (lane >> 4);
                    const uint32_t cid = __shfl_sync(0xffffffffu, my_id, c);
                    stage[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < nbatch && off < a.d)
                        stage[i] = __ldg(reinterpret_cast<const float4 *>(a.base + (size_t)cid * a.d + off));
                }
                __syncwarp();  // the previous chunk has been consumed
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    *reinterpret_cast<float4 *>(tile + (2 * i + (lane >> 4)) * RR_TSTR + (lane & 15) * 4) =
                        stage[i];
                __syncwarp();
                if (lane < nbatch) {
                    const float *tr = tile + lane * RR_TSTR;
                    const int jn = a.d - j0 < RR_DC ? a.d - j0 : RR_DC;
                    for (int j = 0; j < jn; j += 4) {
                        const float4 bv = *reinterpret_cast<const float4 *>(tr + j);
                        const float4 qv = *reinterpret_cast<const float4 *>(qrow + j0 + j);
                        float t;
                        t = __fsub_rn(qv.x, bv.x); acc = __fadd_rn(acc, __fmul_rn(t, t));
                        t = __fsub_rn(qv.y, bv.y); acc = __fadd_rn(acc, __fmul_rn(t, t));
                        t = __fsub_rn(qv.z, bv.z); acc = __fadd_rn(acc, __fmul_rn(t, t));
                        t = __fsub_rn(qv.w, bv.w); acc = __fadd_rn(acc, __fmul_rn(t, t));
                    }
                }
            }
            if (lane < nbatch) keys[b0 + lane] = make_key(acc, my_id);
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            const uint32_t id = key_pos(keys[i]);
            const float *brow = a.base + (size_t)id * a.d;
            float acc = 0.0f;
            if ((a.d & 3) == 0) {
                const float4 *b4 = reinterpret_cast<const float4 *>(brow);
                for (int j = 0; j < a.d / 4; ++j) {
                    const float4 bv = __ldg(b4 + j);
                    float t;
                    t = __fsub_rn(qrow[4 * j + 0], bv.x); acc = __fadd_rn(acc, __fmul_rn(t, t));
                    t = __fsub_rn(qrow[4 * j + 1], bv.y); acc = __fadd_rn(acc, __fmul_rn(t, t));
                    t = __fsub_rn(qrow[4 * j + 2], bv.z); acc = __fadd_rn(acc, __fmul_rn(t, t));
                    t = __fsub_rn(qrow[4 * j + 3], bv.w); acc = __fadd_rn(acc, __fmul_rn(t, t));
                }
            } else {
                for (int j = 0; j < a.d; ++j) {
                    float t = __fsub_rn(qrow[j], __ldg(brow + j));
                    acc = __fadd_rn(acc, __fmul_rn(t, t));
                }
            }
            keys[i] = make_key(acc, id);
        }
    }
    __syncwarp();
    warp_sort_keys_n(keys, n, lane);  // n <= NMAX, NMAX a power of two

    // certificate: every non-candidate j has S'_j <= tau, hence
    //   true dist_j >= |x|^2 - 2 (tau + eps) / sigma^2   and the oracle's FP32 value
    //   >= true * (1 - gamma).  Accept iff that lower bound > k-th exact distance.
    const int m = n < a.kcap ? n : a.kcap;
    bool ok = true;
    if (tau > __int_as_float(0xff800000)) {
        if (m < a.kcap) {
            ok = false;  // cannot happen (a threshold implies >= L >= kcap candidates)
        } else {
            const double gamma = (double)(a.d + 2) * (1.0 / 16777216.0) * 1.0001;
            double lb;
            if (a.i8) {
                // tau is an exact integer score T = x.y' - ceil(|y|^2/2); a non-candidate has
                // dist = |x|^2 - 256 sum(x) - 2 T - {0,1} >= |x|^2 - 256 sum(x) - 2 tau - 1
                double x1 = 0.0;
                for (int i = lane; i < a.d; i += 32) x1 += (double)qrow[i];
                for (int o = 16; o > 0; o >>= 1) x1 += __shfl_xor_sync(0xffffffffu, x1, o);
                lb = (xn2 - 256.0 * x1 - 2.0 * (double)tau - 1.0) * (1.0 - gamma);
            } else {
                lb = (xn2 * (1.0 - 1e-12) - 2.0 * ((double)tau + eps) / (sigma * sigma)) *
                     (1.0 - gamma);
            }
            const double dk = (double)key_score(keys[a.kcap - 1]);
            ok = lb > dk;
        }
    } else if (n < a.kcap && n < a.nb - (a.self_join ? 1 : 0)) {
        // no threshold and fewer candidates than the base could give: scores the filter could
        // not compare (NaN / Inf input); the exact kernel decides
        ok = false;
    }

    const size_t orow = (size_t)(q - a.out_q_begin) * a.k;
    int first = 0;
    if (a.self_join) {
        if (lane == 0) {
            a.assign[orow] = q;
            a.dis[orow] = 0.0f;
        }
        first = 1;
    }
    for (int i = lane; i < a.k - first; i += 32) {
        if (i < m) {
            a.assign[orow + first + i] = (int)key_pos(keys[i]);
            a.dis[orow + first + i] = key_score(keys[i]);
        } else {
            a.assign[orow + first + i] = -1;
            a.dis[orow + first + i] = __int_as_float(0x7f800000);
        }
    }
    if (lane == 0) {
        atomicAdd(const_cast<unsigned long long *>(&a.sc->cand_total[blockIdx.x & 31]),
                  (unsigned long long)n);
        if (!ok) {
            const uint32_t p = atomicAdd(a.fb_count, 1u);
            a.fb_rows[p] = q;
        }
    }
}

// ------------------------------------------------------------------------------------
// host side
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode_fn() {
    static encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<encode_tiled_fn>(p);
    return fn;
}

// fp16 matrix [rows x cols] row-major (row pitch = cols), box = [box_cols x box_rows]
bool make_map(CUtensorMap *m, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_cols,
              uint32_t box_rows) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(__half)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols * 2 == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool build_plan(int d16, TcPlan *p) {
    p->nseg = 0;
    int a_off = 0, b_off = 0, k = 0;
    auto add = [&](int bytes, int fold) {
        if (p->nseg >= TC_MAXSEG) return false;
        TcSeg &s = p->seg[p->nseg++];
        s.k_off = fold ? 0 : k;
        s.bytes = bytes;
        s.fold = fold;
        s.a_off = a_off;
        s.b_off = b_off;
        a_off += TC_ROWS * bytes;
        b_off += TC_BN * bytes;
        if (!fold) k += bytes / 2;
        return true;
    };
    int rem = d16;
    while (rem >= 64) {
        if (!add(128, 0)) return false;
        rem -= 64;
    }
    if (rem >= 32) {
        if (!add(64, 0)) return false;
        rem -= 32;
    }
    if (rem >= 16) {
        if (!add(32, 0)) return false;
        rem -= 16;
    }
    if (!add(32, 1)) return false;
    p->a_bytes = a_off;
    p->b_bytes = b_off;
    return true;
}

// candidates kept per row (slack for the certificate): ~1.5x the neighbours asked for when the
// FP16 operands are exact (integer-valued data), max(64, 2x) when they are rounded
int pick_L(int kcap, bool inexact) {
    if (!inexact) {
        // exact operands: the certificate only has to absorb the accumulation bound, a
        // margin of half the neighbours is plenty; a short list keeps the threshold tight
        if (env().l_exact > 0) return env().l_exact;
        return (kcap + kcap / 2 + 7) / 8 * 8;
    }
    if (env().l_inexact > 0) return env().l_inexact;
    // rounded operands: 64 candidates, or twice the neighbours asked for if that is more.
    // Measured (profiles/README.md): GIST-shaped k=16 needs 64 (at 32, 10 % of the rows fail the
    // certificate); Deep-shaped k=32 certifies every row at 64 and runs 1.5x faster than at 128.
    int L = 2 * kcap;
    L = (L + 31) / 32 * 32;
    if (L < 64) L = 64;
    return L;
}
// Integer path: scores are exact, the certificate's only slack is the parity bit of |y|^2 (a
// non-candidate is at least dist(tau) - 1 away), so a quarter more than the neighbours asked for
// is enough; a row with ties at its k-th distance goes to the second pass either way.  Measured
// at L = 20 instead of 24 for k = 16: no row lost, 5 % off short sweeps, 25 % off the re-rank.
int pick_L_i8(int kcap) {
    if (env().l_exact > 0) return env().l_exact;
    const int slack = kcap / 4 > 4 ? kcap / 4 : 4;
    return (kcap + slack + 3) / 4 * 4;
}

// List capacity.  The filter checks the count once per tile, after up to TC_BN (128) appends,
// and a shrink leaves up to L + L/2 entries: L + L/2 + TC_BN must fit.
int pick_gcap(int L) {
    const int need = L + L / 2 + TC_BN;
    const int e = env().gcap;
    if ((e == 256 || e == 512) && need <= e) return e;
    return need <= 256 ? 256 : 512;
}

size_t tc_smem_bytes(const TcPlan &p) {
    return 1024 + (size_t)p.a_bytes + (size_t)TC_NSTAGE * p.b_bytes + sizeof(TcSmemTail);
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// does the resident-A plan fit in shared memory for this operand width?
bool resident_fits(int op_cols) {
    TcPlan p;
    if (!build_plan(op_cols, &p)) return false;
    return tc_smem_bytes(p) <= 227 * 1024;
}

constexpr int P2_GCAP = 512, P2_MAX_LISTS = 4;

// work decomposition of one launch: n_rows query rows against n_btiles base tiles
struct TcGeom {
    int n_qtiles, nsplit, tiles_per_split;
    int q_tail0, ns_tail, tiles_per_split_tail;
    size_t n_slots;  // (row, split) pairs: candidate lists
};

void tc_geom(int n_rows, int n_btiles, int gcap, int num_sms, TcGeom *g) {
    g->n_qtiles = (n_rows + TC_ROWS - 1) / TC_ROWS;
    // Base splits: a work item is (query tile, base split).  A split costs a second candidate list
    // per row with its own warm-up (measured: 100k x 128 takes 5.0 / 6.7 / 8.6 ms with 1 / 2 / 4
    // splits although 4 splits fill the last round of the grid best), so the base is only split
    // when the query tiles alone cannot occupy every SM once.
    const int max_lists = 2048 / gcap;
    int nsplit = 1;
    while (g->n_qtiles * nsplit < num_sms && nsplit * 2 <= max_lists && nsplit < 4 &&
           n_btiles / (nsplit * 2) >= 16)
        nsplit *= 2;
    if (env().nsplit >= 1 && env().nsplit <= max_lists && n_btiles / env().nsplit >= 8)
        nsplit = env().nsplit;
    g->nsplit = nsplit;
    g->tiles_per_split = (n_btiles + nsplit - 1) / nsplit;
    // The last round of the persistent grid: if the query tiles left over after the full rounds
    // occupy at most half of the SMs, cut each of them into 2..4 base splits so the round
    // finishes in a fraction of an item time (8 GPUs x 125k rows: 489 tiles = 3 rounds + 45).
    g->q_tail0 = g->n_qtiles;
    g->ns_tail = 1;
    g->tiles_per_split_tail = n_btiles;
    if (nsplit == 1 && g->n_qtiles > num_sms && !env().no_tail) {
        const int rem = g->n_qtiles % num_sms;
        int ns = rem > 0 ? num_sms / rem : 1;
        if (ns > 4) ns = 4;
        while (ns > 1 && (ns > max_lists || n_btiles / ns < 16)) --ns;
        if (ns > 1) {
            g->q_tail0 = g->n_qtiles - rem;
            g->ns_tail = ns;
            g->tiles_per_split_tail = (n_btiles + ns - 1) / ns;
        }
    }
    const long long tail_row0 = (long long)g->q_tail0 * TC_ROWS;
    const long long rows_main = tail_row0 < n_rows ? tail_row0 : n_rows;
    g->n_slots = (size_t)rows_main * nsplit + (size_t)(n_rows - rows_main) * g->ns_tail;
}

// Device scratch of one search.  Everything that scales with the candidate lists is sized for
// ONE ROW SLICE: the first pass runs slice by slice over the query rows, so the lists never
// exceed $PQH_TC_SLICE_MB (2 GiB by default; a slice is a whole number of rounds of the
// persistent grid), whatever n is.
struct TcLayout {
    int d16, d32, L, L_exact, L_i8, gcap, nmma, n_btiles;
    bool streaming, order;
    int slice_rows, slice_qtiles;
    size_t scalars, xh_b, fold_b, fold_a, norm_b, norm_q, xq, cand, cnt, tau, fb_rows, order_ws;
    size_t piv_op, piv_fold;  // locality order: operand rows and fold pieces of the pivots
    int cells;
    // second pass (rows that failed the first certificate, redone with longer lists)
    size_t p2_q, p2_xh, p2_norm, p2_cand, p2_cnt, p2_tau, p2_fb, p2_self;
    int p2_rows;  // rows per second-pass batch
    size_t total;
};

void tc_layout(int n_rows, int nb, int d, int k, int self_join, int num_sms, TcLayout *lo) {
    const TcEnv &ev = env();
    const int kcap = self_join ? k - 1 : k;
    lo->d16 = (d + 15) / 16 * 16;
    lo->d32 = (d + 31) / 32 * 32;
    lo->L = pick_L(kcap, true);
    lo->L_exact = pick_L(kcap, false);
    lo->L_i8 = pick_L_i8(kcap);
    lo->gcap = pick_gcap(lo->L);
    lo->streaming = ev.streaming || !resident_fits(lo->d16);
    lo->nmma = lo->streaming ? 1 : ev.nmma;
    lo->cells = ev.order ? order_num_cells(nb) : 0;
    lo->order = lo->cells > 0;
    lo->n_btiles = (nb + TC_BN - 1) / TC_BN;
    // row slice: as many whole rounds of the grid as the list budget allows
    {
        const size_t per_row = (size_t)lo->gcap * sizeof(key64_t);
        long long rows = (long long)((size_t)ev.slice_mb * 1024 * 1024 / per_row);
        const long long round = (long long)num_sms * TC_ROWS;
        rows = rows / round * round;
        if (rows < round) rows = round;
        if (rows > n_rows) rows = n_rows;
        lo->slice_rows = (int)rows;
        lo->slice_qtiles = (lo->slice_rows + TC_ROWS - 1) / TC_ROWS;
    }
    // list slots of the largest slice (base splits only happen for small slices)
    size_t max_slots = 0;
    {
        TcGeom g;
        tc_geom(lo->slice_rows, lo->n_btiles, lo->gcap, num_sms, &g);
        max_slots = g.n_slots;
        const int last = n_rows % lo->slice_rows;
        if (last) {
            tc_geom(last, lo->n_btiles, lo->gcap, num_sms, &g);
            if (g.n_slots > max_slots) max_slots = g.n_slots;
        }
    }
    const size_t n_lists = max_slots;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    lo->scalars = take(sizeof(PrepScalars));
    lo->xh_b = take((size_t)nb * lo->d16 * sizeof(__half));
    lo->fold_b = take((size_t)nb * FOLD_COLS * sizeof(__half));
    lo->fold_a = take((size_t)TC_ROWS * FOLD_COLS * sizeof(__half));
    lo->norm_b = take((size_t)nb * sizeof(double));
    lo->norm_q = self_join ? lo->norm_b : take((size_t)n_rows * sizeof(double));
    lo->xq = take((size_t)lo->slice_rows * lo->d16 * sizeof(__half));
    lo->cand = take(n_lists * lo->gcap * sizeof(key64_t));
    lo->cnt = take(n_lists * sizeof(int));
    lo->tau = take(n_lists * sizeof(float));
    lo->fb_rows = take((size_t)n_rows * sizeof(int));
    lo->order_ws = lo->order ? take(order_workspace_bytes(lo->slice_rows, nb, lo->slice_qtiles)) : 0;
    lo->piv_op = lo->order ? take((size_t)lo->cells * lo->d16 * sizeof(__half)) : 0;
    lo->piv_fold = lo->order ? take((size_t)lo->cells * FOLD_COLS * sizeof(__half)) : 0;
    {
        int p2 = n_rows / 16;
        if (p2 < 2048) p2 = 2048;
        if (p2 > 65536) p2 = 65536;
        if (p2 > n_rows) p2 = n_rows;
        lo->p2_rows = p2;
        lo->p2_q = take((size_t)p2 * d * sizeof(float));
        lo->p2_xh = take((size_t)p2 * lo->d16 * sizeof(__half));
        lo->p2_norm = take((size_t)p2 * sizeof(double));
        lo->p2_cand = take((size_t)p2 * P2_MAX_LISTS * P2_GCAP * sizeof(key64_t));
        lo->p2_cnt = take((size_t)p2 * P2_MAX_LISTS * sizeof(int));
        lo->p2_tau = take((size_t)p2 * P2_MAX_LISTS * sizeof(float));
        lo->p2_self = take((size_t)p2 * sizeof(int));
        lo->p2_fb = take((size_t)n_rows * sizeof(int));
    }
    lo->total = off + 1024;
}

template <int GCAP, bool I8, int NMMA, int MODE>
cudaError_t launch_tc_one(const TcMaps &maps, const TcArgs &args, size_t smem, int grid,
                          cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<GCAP, I8, NMMA, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    knn_tc_kernel<GCAP, I8, NMMA, MODE><<<grid, tc_threads(NMMA), smem, stream>>>(maps, args);
    return cudaGetLastError();
}
template <int GCAP, bool I8>
cudaError_t launch_tc_cfg(const TcMaps &maps, const TcArgs &args, size_t smem, int grid, int nmma,
                          bool early, cudaStream_t stream) {
#ifdef PQH_TC_PROF  // diagnostic build: cycle accounting of the filter warps (PQH_TC_DEBUG=3)
    if (early && nmma == 2 && GCAP == 256 && args.debug_mode == 3)
        return launch_tc_one<256, I8, 2, 4>(maps, args, smem, grid, stream);
#endif
    if (early)
        return nmma == 2 ? launch_tc_one<GCAP, I8, 2, 1>(maps, args, smem, grid, stream)
                         : launch_tc_one<GCAP, I8, 1, 1>(maps, args, smem, grid, stream);
    return nmma == 2 ? launch_tc_one<GCAP, I8, 2, 0>(maps, args, smem, grid, stream)
                     : launch_tc_one<GCAP, I8, 1, 0>(maps, args, smem, grid, stream);
}

}  // namespace

void tc_reload_env() {
    TcEnv e;
    e.i8 = env_int("PQH_TC_I8", 1) != 0;
    e.order = env_int("PQH_TC_ORDER", 1) != 0;
    e.nmma = env_int("PQH_TC_MMA", 2) == 1 ? 1 : 2;
    e.early = env_int("PQH_TC_EARLY", 1) != 0;
    e.own_tiles = env_int("PQH_TC_OWN", 8);
    if (e.own_tiles < 1) e.own_tiles = 1;
    if (e.own_tiles > 64) e.own_tiles = 64;
    e.l_exact = env_int("PQH_TC_LEXACT", 0);
    e.l_inexact = env_int("PQH_TC_LINEXACT", 0);
    e.gcap = env_int("PQH_TC_GCAP", 0);
    e.nsplit = env_int("PQH_TC_NSPLIT", 0);
    e.no_tail = env_int("PQH_TC_NO_TAIL", 0);
    e.no_pass2 = env_int("PQH_TC_NO_PASS2", 0);
    e.check_eps = env_int("PQH_TC_CHECK_EPS", 0);
    e.debug = env_int("PQH_TC_DEBUG", 0);
    e.streaming = env_int("PQH_TC_STREAM", 0) != 0;
    e.dynamic = env_int("PQH_TC_DYN", 1) != 0;
    e.wide_d = env_int("PQH_TC_WIDE_D", 64);
    e.slice_mb = env_int("PQH_TC_SLICE_MB", 2048);
    if (e.slice_mb < 64) e.slice_mb = 64;
    e.loaded = true;
    g_env = e;
}

bool tc_supported(int n_rows, int nb, int d, int k, int self_join) {
    const int kcap = self_join ? k - 1 : k;
    if (kcap < 1 || kcap > 64) return false;  // list capacity 512 holds L = 128 candidates
    if (nb < 2048 || n_rows < 1) return false;
    if (d < 1 || d > 4096) return false;
    return true;  // the streaming kernel takes every d the resident kernel cannot
}

size_t tc_workspace_bytes(int n_rows, int nb, int d, int k, int self_join, int num_sms) {
    TcLayout lo;
    tc_layout(n_rows, nb, d, k, self_join, num_sms, &lo);
    return lo.total;
}

// Tensor-core search of query rows [q_begin, q_begin + n_rows).
//   d_queries: self-join -> same matrix as d_base (row q is query q);
//              otherwise a matrix whose row (q - q_row0) is query q.
// Rows that fail the certificate are written to *h_fb_count / d_fb_rows (absolute ids);
// the caller redoes them with the exact kernel.  Synchronises the stream: once to learn which
// operand flavour the data allows, once for the number of rows that failed (and once more after
// a second pass).
cudaError_t launch_knn_tc(const float *d_base, const float *d_queries, int q_row0, int q_begin,
                          int out_q_begin, int n_rows, int nb, int d, int k, int self_join,
                          int *d_assign, float *d_dis, void *d_work, size_t work_bytes,
                          int num_sms, cudaStream_t stream, int *launches, int **d_fb_rows_out,
                          unsigned *h_fb_count, TcTimes *times) {
    *h_fb_count = 0;
    if (n_rows <= 0) return cudaSuccess;
    const TcEnv &ev = env();
    const int kcap = self_join ? k - 1 : k;
    TcLayout lo;
    tc_layout(n_rows, nb, d, k, self_join, num_sms, &lo);
    if (lo.total > work_bytes) return cudaErrorMemoryAllocation;
    unsigned char *w = reinterpret_cast<unsigned char *>(
        ((uintptr_t)d_work + 1023) & ~(uintptr_t)1023);
    PrepScalars *sc = reinterpret_cast<PrepScalars *>(w + lo.scalars);
    __half *xh_b = reinterpret_cast<__half *>(w + lo.xh_b);
    __half *fold_b = reinterpret_cast<__half *>(w + lo.fold_b);
    __half *fold_a = reinterpret_cast<__half *>(w + lo.fold_a);
    double *norm_b = reinterpret_cast<double *>(w + lo.norm_b);
    double *norm_q = reinterpret_cast<double *>(w + lo.norm_q);
    void *xq = w + lo.xq;
    key64_t *cand = reinterpret_cast<key64_t *>(w + lo.cand);
    int *cnt = reinterpret_cast<int *>(w + lo.cnt);
    float *tau = reinterpret_cast<float *>(w + lo.tau);
    int *fb_rows = reinterpret_cast<int *>(w + lo.fb_rows);
    void *order_ws = w + lo.order_ws;
    *d_fb_rows_out = fb_rows;
    const int debug_mode = ev.debug;

    // events: [0],[1] bracket the preparation; per slice (gemm start, gemm end, re-rank end)
    const int n_slices = (n_rows + lo.slice_rows - 1) / lo.slice_rows;
    std::vector<cudaEvent_t> evs;
    if (times) {
        evs.resize(2 + 3 * (size_t)n_slices);
        for (auto &x : evs) cudaEventCreate(&x);
    }
    struct EvGuard {
        std::vector<cudaEvent_t> &v;
        ~EvGuard() {
            for (auto x : v) cudaEventDestroy(x);
        }
    } ev_guard{evs};

    cudaError_t e;
    if (times) cudaEventRecord(evs[0], stream);
    if ((e = cudaMemsetAsync(sc, 0, sizeof(PrepScalars), stream)) != cudaSuccess) return e;

    // queries of this call: row (q - q_begin + q_src0) of q_mat
    const float *q_mat = self_join ? d_base : d_queries;
    const int q_src0 = self_join ? q_begin : q_begin - q_row0;

    // ---- which operand flavour does the data allow?  Integers in [0, 255] whose norms fit the
    //      fold columns take the exact INT8 path; one flag read-back decides ----
    bool used_i8 = false;
    const bool i8_possible = ev.i8 && !lo.streaming && (debug_mode >= 0 && debug_mode <= 3) &&
                             resident_fits(lo.d32 / 2);
    if (i8_possible) {
        prep_rows_i8_kernel<<<(nb + 7) / 8, 256, 0, stream>>>(d_base, nb, d, lo.d32, nullptr,
                                                              nullptr, nullptr, 1, nullptr, 0, sc);
        ++*launches;
        if (!self_join) {
            prep_rows_i8_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(
                q_mat + (size_t)q_src0 * d, n_rows, d, lo.d32, nullptr, nullptr, nullptr, 0,
                nullptr, 0, sc);
            ++*launches;
        }
        PrepScalars probe;
        if ((e = cudaMemcpyAsync(&probe, sc, sizeof(PrepScalars), cudaMemcpyDeviceToHost,
                                 stream)) != cudaSuccess)
            return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        used_i8 = probe.i8_reject == 0;
    }
    // FP16 flavour: the common scale of all operands
    auto prep_scale_fp16 = [&]() {
        const int grid_ma = num_sms * 8;
        maxabs_kernel<<<grid_ma, 256, 0, stream>>>(d_base, (size_t)nb * d, &sc->maxabs_bits);
        ++*launches;
        if (!self_join) {
            maxabs_kernel<<<grid_ma, 256, 0, stream>>>(q_mat + (size_t)q_src0 * d,
                                                       (size_t)n_rows * d, &sc->maxabs_bits);
            ++*launches;
        }
    };
    if (!used_i8) prep_scale_fp16();
    if (used_i8)
        fold_const_i8_kernel<<<(TC_ROWS * 32 + 255) / 256, 256, 0, stream>>>(
            reinterpret_cast<uint8_t *>(fold_a));
    else
        fold_const_kernel<<<(TC_ROWS * FOLD_COLS + 255) / 256, 256, 0, stream>>>(fold_a);
    ++*launches;
    // operand rows of a set of data rows (perm: operand row p is data row perm[p] - perm_off)
    auto prep_rows = [&](bool i8, const float *src, int n, void *op, double *norms, void *fold,
                         bool as_base, const int *perm, int perm_off) {
        if (i8)
            prep_rows_i8_kernel<<<(n + 7) / 8, 256, 0, stream>>>(
                src, n, d, lo.d32, static_cast<uint8_t *>(op), norms, static_cast<int8_t *>(fold),
                as_base ? 1 : 0, perm, perm_off, sc);
        else
            prep_rows_kernel<<<(n + 7) / 8, 256, 0, stream>>>(
                src, n, d, lo.d16, static_cast<__half *>(op), norms, static_cast<__half *>(fold),
                as_base ? 1 : 0, perm, perm_off, sc);
        ++*launches;
    };

    const int *base_perm = nullptr;  // set once the locality order of the base exists

    // ---- one pass = tensor-core kernel + (re-rank | cell pick) over a set of query rows ----
    struct Pass {
        TcGeom g;
        int L_exact, L_inexact, gcap, nmma, kcap, self_join;
        bool streaming;
        const void *a_op;        // query operand: row (q_op_begin + r) is query row r
        uint64_t a_rows;         // rows of that operand
        int q_op_begin;
        const void *b_op;        // base operand and its fold pieces (nb_rows rows)
        const void *b_fold;
        int nb_rows, n_btiles;
        const float *queries;    // FP32 rows of the pass's queries (see RerankArgs)
        const double *qnorm64;   // indexed like `queries`
        int q_begin;             // absolute query id of row 0 (natural order) / of queries[0]
        const int *row_ids, *q_perm, *q_self, *qt_start;
        int n_rows;
        key64_t *cand;
        int *cnt;
        float *tau;
        int *fb_rows;
        uint32_t *fb_count;
        bool i8;                 // integer operands
        int op_cols;             // operand row length in 2-byte units (the tensor maps' view)
        cudaEvent_t *ev3;        // gemm start / gemm end / re-rank end, or null
        int *pick_cell, *pick_hist;  // locality order: take each row's nearest pivot instead of
                                     // re-ranking (the base operand is the pivot set)
    };
    auto run_pass = [&](const Pass &p) -> cudaError_t {
        TcMaps maps;
        TcArgs args;
        if (!build_plan(p.op_cols, &args.plan) && !p.streaming) return cudaErrorInvalidValue;
        const uint32_t boxc[3] = {64, 32, 16};
        for (int c = 0; c < 3; ++c) {
            // classes a small d never uses still need a legal box
            const uint32_t bc = boxc[c] <= (uint32_t)p.op_cols ? boxc[c] : 16;
            if (!make_map(&maps.a[c], p.a_op, p.a_rows, p.op_cols, bc, TC_ROWS) ||
                !make_map(&maps.b[c], p.b_op, p.nb_rows, p.op_cols, bc, TC_BN))
                return cudaErrorInvalidValue;
        }
        if (!make_map(&maps.fa, fold_a, TC_ROWS, FOLD_COLS, FOLD_COLS, TC_ROWS) ||
            !make_map(&maps.fb, p.b_fold, p.nb_rows, FOLD_COLS, FOLD_COLS, TC_BN))
            return cudaErrorInvalidValue;

        args.n_rows = p.n_rows;
        args.q_begin = p.q_begin;
        args.q_op_begin = p.q_op_begin;
        args.nb = p.nb_rows;
        args.self_join = p.self_join;
        args.row_ids = p.row_ids;
        args.q_self = p.q_self;
        args.qt_start = p.qt_start;
        args.n_qtiles = p.g.n_qtiles;
        args.nsplit = p.g.nsplit;
        args.q_tail0 = p.g.q_tail0;
        args.ns_tail = p.g.ns_tail;
        args.tiles_per_split_tail = p.g.tiles_per_split_tail;
        args.tiles_per_split = p.g.tiles_per_split;
        args.n_btiles = p.n_btiles;
        args.debug_mode = p.pick_cell ? 0 : debug_mode;
        args.dynamic = (ev.dynamic && !p.streaming) ? 1 : 0;
        args.kcap = p.kcap;
        args.L_exact = p.L_exact;
        args.L_inexact = p.i8 ? p.L_exact : p.L_inexact;  // integer scores are exact
        args.sc = sc;
        args.cand = p.cand;
        args.cand_cnt = p.cnt;
        args.cand_tau = p.tau;
        args.pick_cell = p.pick_cell;
        args.pick_hist = p.pick_hist;

        const int n_items =
            p.g.q_tail0 * p.g.nsplit + (p.g.n_qtiles - p.g.q_tail0) * p.g.ns_tail;
        cudaError_t pe;
        if (args.dynamic &&
            (pe = cudaMemsetAsync(&sc->work_next, 0, sizeof(uint32_t), stream)) != cudaSuccess)
            return pe;
        if (p.ev3) cudaEventRecord(p.ev3[0], stream);
        if (p.streaming) {
            if ((pe = launch_tcs(&maps, &args, p.gcap, lo.d16, n_items, num_sms, stream)) !=
                cudaSuccess)
                return pe;
        } else if (p.pick_cell) {
            const int grid = n_items < num_sms ? n_items : num_sms;
            const size_t smem = tc_smem_bytes(args.plan);
            pe = p.i8 ? launch_tc_one<256, true, 2, 2>(maps, args, smem, grid, stream)
                      : launch_tc_one<256, false, 2, 2>(maps, args, smem, grid, stream);
            if (pe != cudaSuccess) return pe;
        } else {
            const int grid = n_items < num_sms ? n_items : num_sms;
            const size_t smem = tc_smem_bytes(args.plan);
            if (p.i8)
                pe = p.gcap == 256 ? launch_tc_cfg<256, true>(maps, args, smem, grid, p.nmma,
                                                               ev.early != 0, stream)
                                   : launch_tc_cfg<512, true>(maps, args, smem, grid, p.nmma,
                                                               ev.early != 0, stream);
            else
                pe = p.gcap == 256 ? launch_tc_cfg<256, false>(maps, args, smem, grid, p.nmma,
                                                                ev.early != 0, stream)
                                   : launch_tc_cfg<512, false>(maps, args, smem, grid, p.nmma,
                                                                ev.early != 0, stream);
            if (pe != cudaSuccess) return pe;
        }
        ++*launches;
        if (p.ev3) cudaEventRecord(p.ev3[1], stream);
        if (p.pick_cell) {
            if (!p.streaming) return cudaSuccess;  // the arg-max epilogue wrote the cells itself
            return launch_order_pick(p.cand, p.cnt, p.n_rows, p.gcap, p.g.nsplit,
                                     p.g.q_tail0 * TC_ROWS, p.g.ns_tail, p.pick_cell, p.pick_hist,
                                     stream, launches);
        }

        // ---- re-rank + certificate ----
        RerankArgs ra;
        ra.base = d_base;
        ra.queries = p.queries;
        ra.q_begin = p.q_begin;
        ra.row_ids = p.row_ids;
        ra.q_perm = p.q_perm;
        ra.base_perm = base_perm;
        ra.out_q_begin = out_q_begin;
        ra.n_rows = p.n_rows;
        ra.nb = nb;
        ra.d = d;
        ra.k = k;
        ra.kcap = kcap;
        ra.self_join = self_join;
        ra.nlists = p.g.nsplit;
        ra.tail_row0 = p.g.q_tail0 * TC_ROWS;
        ra.ns_tail = p.g.ns_tail;
        ra.gcap = p.gcap;
        ra.cand = p.cand;
        ra.cand_cnt = p.cnt;
        ra.cand_tau = p.tau;
        ra.qnorm64 = p.qnorm64;
        ra.bnorm64 = norm_b;
        ra.check_eps = ev.check_eps;
        ra.sc = sc;
        ra.ksteps = (lo.d16 + FOLD_COLS) / 16;
        ra.i8 = p.i8 ? 1 : 0;
        ra.assign = d_assign;
        ra.dis = d_dis;
        ra.fb_rows = p.fb_rows;
        ra.fb_count = p.fb_count;
        const int splits = p.g.nsplit > p.g.ns_tail ? p.g.nsplit : p.g.ns_tail;
        const int nmax = p.gcap * splits;
        const int rgrid = (p.n_rows + RR_WARPS - 1) / RR_WARPS;
        const size_t qbytes = (size_t)RR_WARPS * d * sizeof(float);
        // wide rows whose start is 16-byte aligned go through the coalescing path
        const bool wide = d >= ev.wide_d && (d & 3) == 0 && (reinterpret_cast<uintptr_t>(d_base) & 15) == 0;
        const size_t tbytes = wide ? (size_t)RR_WARPS * 32 * RR_TSTR * sizeof(float) : 0;
#define LAUNCH_RR1(N, W)                                                                         \
    do {                                                                                         \
        size_t sm = (size_t)RR_WARPS * N * sizeof(key64_t) + qbytes + tbytes;                    \
        if ((pe = cudaFuncSetAttribute(knn_rerank_kernel<N, W>,                                  \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)) != \
            cudaSuccess)                                                                         \
            return pe;                                                                           \
        knn_rerank_kernel<N, W><<<rgrid, RR_WARPS * 32, sm, stream>>>(ra);                       \
    } while (0)
#define LAUNCH_RR(N)                                                                             \
    do {                                                                                         \
        if (wide) LAUNCH_RR1(N, true);                                                           \
        else LAUNCH_RR1(N, false);                                                               \
    } while (0)
        if (nmax <= 256) LAUNCH_RR(256);
        else if (nmax <= 512) LAUNCH_RR(512);
        else if (nmax <= 1024) LAUNCH_RR(1024);
        else LAUNCH_RR(2048);
#undef LAUNCH_RR1
#undef LAUNCH_RR
        ++*launches;
        if (p.ev3) cudaEventRecord(p.ev3[2], stream);
        return cudaGetLastError();
    };

    // ---- locality order: the nearest pivot of every row is a k = 1 sweep of the rows against
    //      the pivot set on the tensor cores; its candidate lists give the cells ----
    OrderBufs ob;
    ob.C = 0;
    OrderOut oo = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void *piv_op = w + lo.piv_op, *piv_fold = w + lo.piv_fold;
    // cells of `n` rows of `src` (row i of src) into cell_out / hist, slice by slice
    auto assign_cells = [&](const float *src, int n, int *cell_out, int *hist) -> cudaError_t {
        for (int r0 = 0; r0 < n; r0 += lo.slice_rows) {
            const int rows = n - r0 < lo.slice_rows ? n - r0 : lo.slice_rows;
            prep_rows(used_i8, src + (size_t)r0 * d, rows, xq, nullptr, nullptr, false, nullptr, 0);
            Pass pa;
            tc_geom(rows, lo.cells / TC_BN, lo.gcap, num_sms, &pa.g);
            if (!lo.streaming) {  // arg-max epilogue: one work item per query tile, no base splits
                pa.g.nsplit = pa.g.ns_tail = 1;
                pa.g.q_tail0 = pa.g.n_qtiles;
                pa.g.tiles_per_split = pa.g.tiles_per_split_tail = lo.cells / TC_BN;
                pa.g.n_slots = (size_t)rows;
            }
            pa.L_exact = pa.L_inexact = 8;
            pa.gcap = lo.gcap;
            pa.nmma = lo.nmma;
            pa.kcap = 1;
            pa.self_join = 0;
            pa.streaming = lo.streaming;
            pa.a_op = xq;
            pa.a_rows = (uint64_t)rows;
            pa.q_op_begin = 0;
            pa.b_op = piv_op;
            pa.b_fold = piv_fold;
            pa.nb_rows = lo.cells;
            pa.n_btiles = lo.cells / TC_BN;
            pa.queries = nullptr;
            pa.qnorm64 = nullptr;
            pa.q_begin = 0;
            pa.row_ids = pa.q_perm = pa.q_self = pa.qt_start = nullptr;
            pa.n_rows = rows;
            pa.cand = cand;
            pa.cnt = cnt;
            pa.tau = tau;
            pa.fb_rows = nullptr;
            pa.fb_count = nullptr;
            pa.i8 = used_i8;
            pa.op_cols = used_i8 ? lo.d32 / 2 : lo.d16;
            pa.ev3 = nullptr;
            pa.pick_cell = cell_out + r0;
            pa.pick_hist = hist;
            cudaError_t pe = run_pass(pa);
            if (pe != cudaSuccess) return pe;
        }
        return cudaSuccess;
    };
    if (lo.order) {
        order_carve(order_ws, lo.slice_rows, nb, lo.slice_qtiles, &ob);
        if ((e = launch_order_pivots(ob, d_base, nb, d, stream, launches)) != cudaSuccess) return e;
        prep_rows(used_i8, d_base, ob.C, piv_op, nullptr, piv_fold, true, ob.piv_ids, 0);
        if ((e = assign_cells(d_base, nb, ob.cell_b, ob.hist_b)) != cudaSuccess) return e;
        if ((e = launch_order_sort_base(ob, nb, stream, launches)) != cudaSuccess) return e;
        base_perm = ob.base_perm;
    }

    // ---- base operand (rows in cell order when the locality order is on) ----
    auto prep_base = [&](bool i8) {
        prep_rows(i8, d_base, nb, xh_b, norm_b, fold_b, true, base_perm, 0);
    };
    prep_base(used_i8);
    if (times) cudaEventRecord(evs[1], stream);

    // ---- first pass: the rows of this call, slice by slice ----
    for (int si = 0; si < n_slices; ++si) {
        const int s_begin = q_begin + si * lo.slice_rows;  // absolute id of the slice's row 0
        const int rows = (n_rows - si * lo.slice_rows) < lo.slice_rows
                             ? (n_rows - si * lo.slice_rows)
                             : lo.slice_rows;
        const float *q_slice = q_mat + (size_t)(q_src0 + si * lo.slice_rows) * d;
        double *qn_slice = (self_join ? norm_b + s_begin : norm_q + (size_t)si * lo.slice_rows);
        Pass p1;
        tc_geom(rows, lo.n_btiles, lo.gcap, num_sms, &p1.g);
        if (lo.order) {
            if (!self_join) {  // queries that are not base rows get their cells like the base did
                if ((e = launch_order_clear_queries(ob, stream)) != cudaSuccess) return e;
                if ((e = assign_cells(q_slice, rows, ob.cell_q, ob.hist_q)) != cudaSuccess) return e;
            }
            if ((e = launch_order_sort_queries(ob, s_begin, rows, self_join, p1.g.n_qtiles,
                                               ev.own_tiles, stream, launches, &oo)) != cudaSuccess)
                return e;
        }
        // query operand of the slice
        if (used_i8 || lo.order || !self_join) {
            prep_rows(used_i8, q_slice, rows, xq, self_join ? nullptr : qn_slice, nullptr, false,
                      oo.q_perm, s_begin);
            p1.a_op = xq;
            p1.a_rows = (uint64_t)rows;
            p1.q_op_begin = 0;
        } else {
            p1.a_op = xh_b;  // FP16 self-join in natural order: query q is operand row q
            p1.a_rows = (uint64_t)nb;
            p1.q_op_begin = s_begin;
        }
        p1.L_exact = used_i8 ? lo.L_i8 : lo.L_exact;
        p1.L_inexact = lo.L;
        p1.gcap = lo.gcap;
        p1.nmma = lo.nmma;
        p1.kcap = kcap;
        p1.self_join = self_join;
        p1.streaming = lo.streaming;
        p1.b_op = xh_b;
        p1.b_fold = fold_b;
        p1.nb_rows = nb;
        p1.n_btiles = lo.n_btiles;
        p1.i8 = used_i8;
        p1.op_cols = used_i8 ? lo.d32 / 2 : lo.d16;
        p1.queries = q_slice;
        p1.qnorm64 = qn_slice;
        p1.q_begin = s_begin;
        p1.row_ids = nullptr;
        p1.q_perm = oo.q_perm;
        p1.q_self = oo.q_self;
        p1.qt_start = oo.qt_start;
        p1.n_rows = rows;
        p1.cand = cand;
        p1.cnt = cnt;
        p1.tau = tau;
        p1.fb_rows = fb_rows;
        p1.fb_count = &sc->fb_count;
        p1.ev3 = times ? &evs[2 + 3 * (size_t)si] : nullptr;
        p1.pick_cell = p1.pick_hist = nullptr;
        if ((e = run_pass(p1)) != cudaSuccess) return e;
    }

    PrepScalars hsc;
    if ((e = cudaMemcpyAsync(&hsc, sc, sizeof(PrepScalars), cudaMemcpyDeviceToHost, stream)) !=
        cudaSuccess)
        return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    *h_fb_count = hsc.fb_count;
    if (times && hsc.fb_count > 0) {
        const size_t keep = hsc.fb_count < (1u << 20) ? hsc.fb_count : (1u << 20);
        times->pass1_failed.resize(keep);
        if ((e = cudaMemcpyAsync(times->pass1_failed.data(), fb_rows, keep * sizeof(int),
                                 cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
            return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    }

    // ---- second pass: rows whose certificate failed, redone with four times longer lists.
    //      Their cost on the exact kernel would be ~1000x a tensor-core row, so a data set on
    //      which a few per cent of the rows fail must not fall off a cliff.  Skipped when most
    //      rows fail (FP16 cannot resolve this data at all) or when there are only a handful.
    const unsigned n_fb1 = hsc.fb_count;
    if (n_fb1 >= 64 && n_fb1 <= (unsigned)n_rows / 2 && !ev.no_pass2 && debug_mode == 0) {
        const int L1 = used_i8 ? lo.L_i8 : (hsc.inexact ? lo.L : lo.L_exact);
        int L2 = 4 * L1;
        if (L2 > 192) L2 = 192;
        if (L2 > L1) {
            if (used_i8) {  // the second pass runs on FP16 operands: build them now
                if ((e = cudaMemsetAsync(&sc->maxabs_bits, 0, 3 * sizeof(uint32_t), stream)) !=
                    cudaSuccess)
                    return e;
                prep_scale_fp16();
                fold_const_kernel<<<(TC_ROWS * FOLD_COLS + 255) / 256, 256, 0, stream>>>(fold_a);
                ++*launches;
                prep_base(false);
            }
            float *q2 = reinterpret_cast<float *>(w + lo.p2_q);
            __half *xh2 = reinterpret_cast<__half *>(w + lo.p2_xh);
            double *norm2 = reinterpret_cast<double *>(w + lo.p2_norm);
            int *fb2 = reinterpret_cast<int *>(w + lo.p2_fb);
            int *self2 = reinterpret_cast<int *>(w + lo.p2_self);
            const float *src = self_join ? d_base : d_queries;
            const int src_row0 = self_join ? 0 : q_row0;
            const bool stream2 = lo.streaming;
            for (unsigned done = 0; done < n_fb1; done += (unsigned)lo.p2_rows) {
                const int n2 = (int)(n_fb1 - done < (unsigned)lo.p2_rows ? n_fb1 - done
                                                                          : (unsigned)lo.p2_rows);
                const int *ids = fb_rows + done;
                Pass p2;
                tc_geom(n2, lo.n_btiles, P2_GCAP, num_sms, &p2.g);
                if (p2.g.nsplit > P2_MAX_LISTS || p2.g.ns_tail > P2_MAX_LISTS)
                    return cudaErrorInvalidValue;
                gather_rows_kernel<<<(n2 + 7) / 8, 256, 0, stream>>>(src, src_row0, ids, n2, d, q2);
                prep_rows_kernel<<<(n2 + 7) / 8, 256, 0, stream>>>(q2, n2, d, lo.d16, xh2, norm2,
                                                                   nullptr, 0, nullptr, 0, sc);
                *launches += 2;
                p2.q_self = nullptr;
                if (lo.order && self_join) {
                    if ((e = launch_order_self_of(ids, n2, ob.base_inv, self2, stream, launches)) !=
                        cudaSuccess)
                        return e;
                    p2.q_self = self2;
                }
                if ((e = cudaMemsetAsync(&sc->tile_cursor, 0, sizeof(uint32_t), stream)) !=
                    cudaSuccess)
                    return e;
                p2.L_exact = p2.L_inexact = L2;
                p2.gcap = P2_GCAP;
                p2.nmma = 1;
                p2.kcap = kcap;
                p2.self_join = self_join;
                p2.streaming = stream2;
                p2.b_op = xh_b;
                p2.b_fold = fold_b;
                p2.nb_rows = nb;
                p2.n_btiles = lo.n_btiles;
                p2.pick_cell = p2.pick_hist = nullptr;
                p2.a_op = xh2;
                p2.a_rows = (uint64_t)n2;
                p2.q_op_begin = 0;
                p2.queries = q2;
                p2.qnorm64 = norm2;
                p2.q_begin = q_begin;
                p2.row_ids = ids;
                p2.q_perm = nullptr;
                p2.qt_start = nullptr;
                p2.n_rows = n2;
                p2.cand = reinterpret_cast<key64_t *>(w + lo.p2_cand);
                p2.cnt = reinterpret_cast<int *>(w + lo.p2_cnt);
                p2.tau = reinterpret_cast<float *>(w + lo.p2_tau);
                p2.fb_rows = fb2;
                p2.fb_count = &sc->fb2_count;
                p2.i8 = false;
                p2.op_cols = lo.d16;
                p2.ev3 = nullptr;
                if ((e = run_pass(p2)) != cudaSuccess) return e;
            }
            if ((e = cudaMemcpyAsync(&hsc, sc, sizeof(PrepScalars), cudaMemcpyDeviceToHost,
                                     stream)) != cudaSuccess)
                return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            *h_fb_count = hsc.fb2_count;
            *d_fb_rows_out = fb2;
            if (times) times->rows_pass2 = (long long)n_fb1;
        }
    }

    if (times) {
        times->candidates = 0;
        for (int i = 0; i < 32; ++i) times->candidates += (double)hsc.cand_total[i];
        times->inexact = used_i8 ? 2 : (int)hsc.inexact;
        times->slices = n_slices;
        {
            union { uint32_t u; float f; } cv;
            cv.u = hsc.eps_ratio_bits;
            times->eps_ratio_max = cv.f;
        }
        if (debug_mode)
            fprintf(stderr,
                    "tc dbg: mma threads: barrier waits %.4g cyc of %.4g cyc (sums over CTAs and "
                    "issuers) | slow (warp, tile) pairs %.4g | pass-1 failures %u | mma %d order %d "
                    "slices %d\n",
                    (double)hsc.dbg[0], (double)hsc.dbg[1], (double)hsc.dbg[2], n_fb1, lo.nmma,
                    (int)lo.order, n_slices);
#ifdef PQH_TC_PROF
        if (debug_mode == 3)
            fprintf(stderr,
                    "tc prof: filter warps (cycles summed over warps): wait %.4g (of which idle "
                    "shrinks %.4g, %.4g of them) | tmem copy %.4g | hot path %.4g | scan+append "
                    "%.4g | shrinks at capacity %.4g (%.4g lists)\n",
                    (double)hsc.dbg[3], (double)hsc.dbg[8], (double)hsc.dbg[9], (double)hsc.dbg[4],
                    (double)hsc.dbg[5], (double)hsc.dbg[6], (double)hsc.dbg[7], (double)hsc.dbg[10]);
#endif
        cudaEventElapsedTime(&times->ms_prep, evs[0], evs[1]);
        times->ms_gemm = times->ms_rerank = 0.f;
        for (int si = 0; si < n_slices; ++si) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, evs[2 + 3 * (size_t)si], evs[3 + 3 * (size_t)si]);
            cudaEventElapsedTime(&b, evs[3 + 3 * (size_t)si], evs[4 + 3 * (size_t)si]);
            times->ms_gemm += a;
            times->ms_rerank += b;
        }
    }
    return cudaSuccess;
}

}  // namespace pqh
