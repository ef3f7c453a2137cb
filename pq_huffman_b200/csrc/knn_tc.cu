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
// Nothing of the nq x nb score matrix reaches HBM: an epilogue thread owns one
// query row, copies its 128 accumulator columns from TMEM to registers
// (tcgen05.ld), reduces them to group maxima and only touches memory for the
// rare elements above the row's running threshold (filter_tile_n in
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
// Kernels in this file: knn_tc_kernel (default for d <= ~144: query tile
// resident in shared memory), knn_tc2_kernel (CTA-pair experiment), the prep,
// gather and re-rank kernels, and the host-side orchestration launch_knn_tc.
// knn_tcs.cu holds the streaming kernel for larger d; knn_tc3.cu / knn_tc5.cu
// two further measured experiments (profiles/README.md).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include "knn_tc_common.cuh"
#include "knn_tc_i8.cuh"

#include <type_traits>

namespace pqh {

namespace {

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

// prep 2: one warp per row: fp16 operand row, exact-ish FP64 squared norm, fold pieces
__global__ void __launch_bounds__(256)
prep_rows_kernel(const float *__restrict__ x, int n, int d, int d16, __half *__restrict__ xh,
                 double *__restrict__ norm64, __half *__restrict__ fold, int write_fold,
                 PrepScalars *__restrict__ sc) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float sigma = scale_from_maxabs_bits(sc->maxabs_bits);
    const float *xr = x + (size_t)row * d;
    __half *hr = xh + (size_t)row * d16;
    double acc = 0.0;
    bool inexact = false;
    for (int i = lane; i < d16; i += 32) {
        float v = i < d ? xr[i] : 0.0f;
        float s = v * sigma;  // exact: power-of-two scale
        __half h = __float2half_rn(s);
        if (__half2float(h) != s) inexact = true;
        hr[i] = h;
        acc += (double)v * (double)v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (__any_sync(0xffffffffu, inexact) && lane == 0) atomicOr(&sc->inexact, 1u);
    if (lane == 0 && norm64) norm64[row] = acc;
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

// Integer path (knn_tc_i8.cuh), one warp per row: the row as 8-bit integers -- u8 x for the A
// (query) operand, s8 x - 128 for the B (base) operand -- zero-padded to d32 bytes; FP64 squared
// norm; for base rows the 32 s8 fold digits of G = ceil(|y|^2 / 2).  sc->pad[2] collects the
// reasons the integer path cannot be used: bit 0 = a value is not an integer in [0, 255],
// bit 1 = |y|^2 beyond the capacity of the fold columns.
__global__ void __launch_bounds__(256)
prep_rows_i8_kernel(const float *__restrict__ x, int n, int d, int d32, uint8_t *__restrict__ out,
                    double *__restrict__ norm64, int8_t *__restrict__ fold, int as_base,
                    PrepScalars *__restrict__ sc) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *xr = x + (size_t)row * d;
    uint8_t *o = out + (size_t)row * d32;
    double acc = 0.0;
    bool bad = false;
    for (int i = lane; i < d32; i += 32) {
        const float v = i < d ? xr[i] : 0.0f;
        const int iv = (int)v;
        if (!(v >= 0.0f && v <= 255.0f) || (float)iv != v) bad = true;
        uint8_t b = 0;
        if (i < d) b = as_base ? (uint8_t)(int8_t)(iv - 128) : (uint8_t)iv;
        o[i] = b;
        acc += (double)v * (double)v;
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
    unsigned flags = __any_sync(0xffffffffu, bad) ? 1u : 0u;
    if (norm64 && lane == 0) norm64[row] = acc;
    if (fold) {
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
        fold[(size_t)row * 32 + lane] = (int8_t)b;
    }
    if (flags && lane == 0) atomicOr(&sc->pad[2], flags);
}

__global__ void fold_const_i8_kernel(uint8_t *__restrict__ fa) {
    // [256 x 32]: what every query row multiplies the fold digits by
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_ROWS * 32) return;
    fa[i] = (i % 32) < I8_FOLD_FULL ? 255 : 1;
}

// ------------------------------------------------------------------------------------
// main kernel
template <int GCAP>
struct TcSmemTail {
    uint64_t a_full, a_empty;
    uint64_t b_full[TC_NSTAGE], b_empty[TC_NSTAGE];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    int item_t0[2];  // first base tile of the current / next work item (set by the producer)
    uint64_t t0_ready;  // producer -> epilogue: item_t0 of the next item is written
};

// I8: integer operands and scores (knn_tc_i8.cuh); everything else is shared.
template <int GCAP, bool PROF, bool I8>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcPlan &plan = args.plan;
    uint8_t *smA = smem;
    uint8_t *smB = smA + plan.a_bytes;
    TcSmemTail<GCAP> &tail =
        *reinterpret_cast<TcSmemTail<GCAP> *>(smB + TC_NSTAGE * plan.b_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
        ptx::mbar_init(&tail.a_empty, 1);
        for (int i = 0; i < TC_NSTAGE; ++i) {
            ptx::mbar_init(&tail.b_full[i], 1);
            ptx::mbar_init(&tail.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tail.acc_full[i], 1);
            ptx::mbar_init(&tail.acc_empty[i], TC_EPI_WARPS);
        }
        ptx::mbar_init(&tail.t0_ready, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(&tail.tmem_base);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = tail.tmem_base;

    const int n_items = total_items(args);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (ptx::elect_one()) {
            int stage = 0, it_par = 0, my_next = 0;
            uint32_t ph = 0, a_ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ItemInfo it = decode_item(args, item);
                const int qt = it.qt, t_lo = it.t_lo, t_hi = it.t_hi;
                ptx::mbar_wait(&tail.a_empty, a_ph ^ 1);
                const int t0 = item_start_tile(args.sc, t_lo, t_hi, my_next);
                my_next = t0;  // a full sweep ends where it started
                tail.item_t0[it_par] = t0;
                ptx::mbar_arrive(&tail.t0_ready);
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
                    const int t = item_tile(i, t0, t_lo, t_hi);
                    if ((i & 7) == 0) publish_cursor(args.sc, t);
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
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc =
                I8 ? ptx::make_idesc_i8_u8s8(128, TC_BN) : ptx::make_idesc_f16(128, TC_BN);
            int stage = 0, as = 0;
            uint32_t ph = 0, acc_ph = 0, a_ph = 0;
            long long mma_wait_b = 0, mma_wait_acc = 0;
            const long long mma_t0 = PROF ? clock64() : 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ItemInfo it = decode_item(args, item);
                const int t_lo = it.t_lo, t_hi = it.t_hi;
                ptx::mbar_wait(&tail.a_full, a_ph);
                a_ph ^= 1;
                for (int i = 0; i < t_hi - t_lo; ++i) {
                    const long long m0 = PROF ? clock64() : 0;
                    ptx::mbar_wait(&tail.b_full[stage], ph);
                    const long long m1 = PROF ? clock64() : 0;
                    ptx::mbar_wait(&tail.acc_empty[as], acc_ph ^ 1);
                    const long long m2 = PROF ? clock64() : 0;
                    mma_wait_b += m1 - m0;
                    mma_wait_acc += m2 - m1;
                    ptx::tc_fence_after_sync();
                    const uint32_t b_base = ptx::smem_u32(smB + stage * plan.b_bytes);
                    const uint32_t a_base = ptx::smem_u32(smA);
                    // The two query tiles are independent accumulation chains: alternate
                    // between them so that consecutive MMAs never depend on each other.
                    uint32_t acc = 0;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        const uint32_t b_addr = b_base + sg.b_off;
                        const uint32_t sbo = 8 * sg.bytes, lay = seg_layout(sg.bytes);
                        for (int kk = 0; kk < sg.bytes / 32; ++kk) {
                            const uint64_t bd = ptx::make_smem_desc(b_addr + kk * 32, sbo, lay);
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                const uint32_t a_addr = a_base + sg.a_off + g * 128 * sg.bytes;
                                if constexpr (I8)
                                    ptx::umma_i8(tmem_base + (uint32_t)(as * 256 + g * 128),
                                                 ptx::make_smem_desc(a_addr + kk * 32, sbo, lay),
                                                 bd, idesc, acc);
                                else
                                    ptx::umma_f16(tmem_base + (uint32_t)(as * 256 + g * 128),
                                                  ptx::make_smem_desc(a_addr + kk * 32, sbo, lay),
                                                  bd, idesc, acc);
                            }
                            acc = 1;
                        }
                    }
                    ptx::umma_commit(&tail.b_empty[stage]);
                    ptx::umma_commit(&tail.acc_full[as]);
                    if (++stage == TC_NSTAGE) {
                        stage = 0;
                        ph ^= 1;
                    }
                    as ^= 1;
                    if (as == 0) acc_ph ^= 1;
                }
                ptx::umma_commit(&tail.a_empty);
            }
            if (PROF) {
            PrepScalars *scw = const_cast<PrepScalars *>(args.sc);
            atomicAdd(&scw->dbg[6], (unsigned long long)mma_wait_b);
            atomicAdd(&scw->dbg[7], (unsigned long long)mma_wait_acc);
            atomicAdd(&scw->cand_total[63], (unsigned long long)(clock64() - mma_t0));
            }
        }
    } else {
        // ===================== epilogue: threshold filter =====================
        const int e = warp - 2;
        const int g = e >> 2;
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        const int row_in_cta = g * 128 + quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int L = args.sc->inexact ? args.L_inexact : args.L_exact;
        int as = 0;
        uint32_t acc_ph = 0;
        unsigned ctr[4] = {0, 0, 0, 0};
        long long t_wait = 0, t_work = 0;
        long long tprof[4] = {0, 0, 0, 0};
        int it_par = 0;
        using tau_t = std::conditional_t<I8, int, float>;
        const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
        tau_t TAU_LO, TAU_HI;  // "nothing seen yet" / "row does not exist"
        if constexpr (I8) {
            TAU_LO = (int)0x80000000;
            TAU_HI = 0x7fffffff;
        } else {
            TAU_LO = NEG_INF;
            TAU_HI = POS_INF;
        }
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const ItemInfo it = decode_item(args, item);
            const int qt = it.qt, sp = it.sp, t_lo = it.t_lo, t_hi = it.t_hi;
            const int r = qt * TC_ROWS + row_in_cta;
            const bool valid = r < args.n_rows;
            const int self = (valid && args.self_join)
                                 ? (args.row_ids ? args.row_ids[r] : args.q_begin + r)
                                 : -1;
            const size_t slot = valid ? list_slot(r, sp, args.q_tail0 * TC_ROWS, args.nsplit, args.ns_tail)
                                      : 0;
            key64_t *list = args.cand + slot * GCAP;
            tau_t tau = valid ? TAU_LO : TAU_HI;
            int cnt = 0;
            long long b_hi = (long long)t_hi * TC_BN;
            if (b_hi > args.nb) b_hi = args.nb;
            const int ib_hi = (int)b_hi;

            int t0 = 0;
            for (int i = 0; i < t_hi - t_lo; ++i) {
                const long long tw0 = PROF ? clock64() : 0;
                if constexpr (I8)
                    wait_acc_with_idle_shrinks_i32<GCAP>(&tail.acc_full[as], acc_ph, L, lane, tau,
                                                         cnt, list);
                else
                    wait_acc_with_idle_shrinks<GCAP>(&tail.acc_full[as], acc_ph, L, lane, nullptr,
                                                     tau, cnt, list, ctr);
                const long long tw1 = PROF ? clock64() : 0;
                t_wait += tw1 - tw0;
                ptx::tc_fence_after_sync();
                if (i == 0) {
                    ptx::mbar_wait(&tail.t0_ready, (uint32_t)it_par);
                    t0 = *reinterpret_cast<volatile int *>(&tail.item_t0[it_par]);
                }
                const int t = item_tile(i, t0, t_lo, t_hi);
                const uint32_t tbase = tmem_base + lane_addr + (uint32_t)(as * 256 + g * 128);
                // id checks are only needed in the tile holding this row's own column and in
                // the (possibly ragged) last tile
                const int tile0 = t * TC_BN;
                const bool special = (t == t_hi - 1) || (self >= tile0 && self < tile0 + TC_BN);
                auto release = [&]() {
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tail.acc_empty[as]);
                };
                if constexpr (I8) {
                    if (args.debug_mode != 1)
                        filter_tile_i32<GCAP>(tbase, tile0, special, ib_hi, self, L, lane, tau, cnt,
                                              list, release);
                    else
                        release();
                } else {
                    if (args.debug_mode != 1)
                        filter_tile<GCAP>(tbase, tile0, special, ib_hi, self, L, lane,
                                          args.debug_mode, nullptr, tau, cnt, list, ctr, release,
                                          args.debug_mode == 9 ? tprof : nullptr);
                    else
                        release();
                }
                ctr[3] += TC_BN / 32;
                if (PROF) t_work += clock64() - tw1;
                as ^= 1;
                if (as == 0) acc_ph ^= 1;
            }
            it_par ^= 1;
            if (valid) {
                args.cand_cnt[slot] = cnt;
                if constexpr (I8)  // |score| < 2^24: exact as a float; "no threshold" = -inf
                    args.cand_tau[slot] = tau == TAU_LO ? NEG_INF : (float)tau;
                else
                    args.cand_tau[slot] = tau;
            }
        }
        if (PROF && lane == 0) {
            PrepScalars *scw = const_cast<PrepScalars *>(args.sc);
            for (int i = 0; i < 4; ++i) atomicAdd(&scw->dbg[i], (unsigned long long)ctr[i]);
            atomicAdd(&scw->dbg[4], (unsigned long long)t_wait);
            atomicAdd(&scw->dbg[5], (unsigned long long)t_work);
            if (args.debug_mode == 9)
                for (int i = 0; i < 4; ++i)
                    atomicAdd(&scw->cand_total[40 + i], (unsigned long long)tprof[i]);
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}


// ------------------------------------------------------------------------------------
// Pair kernel: two CTAs of one cluster (one TPC) run tcgen05.mma.cta_group::2.  The pair owns
// 256 query rows (each CTA: 128, its M half, A operand resident) and sweeps the base in tiles of
// 256 rows; each CTA TMA-loads only ITS 128-row half of a base tile and the tensor cores of both
// SMs consume both halves.  Versus the single-CTA kernel this halves the shared-memory operand
// traffic per MMA (the single-CTA M=128,N=128 shape needs the full 128 B/cycle of shared memory
// for operands alone) at the same L2 traffic per SM.
// Synchronisation: "full" barriers live in the leader CTA (rank 0) and count the TMA bytes of
// both CTAs; "empty"/"accumulator full" barriers are per CTA and are signalled by multicast
// tcgen05.commit; "accumulator empty" lives in the leader and is arrived on remotely by the
// epilogue warps of both CTAs.
constexpr int TC2_NSTAGE = 4;

template <int GCAP>
struct Tc2SmemTail {
    uint64_t a_full, a_empty;
    uint64_t b_full[TC2_NSTAGE], b_empty[TC2_NSTAGE];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tau_shared[128];
    uint32_t tmem_base;
};

template <int GCAP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
knn_tc2_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcPlan &plan = args.plan;  // a_off/a_bytes are for 256 rows: this kernel uses half
    const int a_bytes = plan.a_bytes / 2;
    uint8_t *smA = smem;
    uint8_t *smB = smA + a_bytes;
    Tc2SmemTail<GCAP> &tail =
        *reinterpret_cast<Tc2SmemTail<GCAP> *>(smB + TC2_NSTAGE * plan.b_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

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
        ptx::mbar_init(&tail.a_empty, 1);
        for (int i = 0; i < TC2_NSTAGE; ++i) {
            ptx::mbar_init(&tail.b_full[i], 1);
            ptx::mbar_init(&tail.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tail.acc_full[i], 1);
            ptx::mbar_init(&tail.acc_empty[i], 2 * 4);  // 4 filter warps in each CTA
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc_2sm<512>(&tail.tmem_base);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();  // barriers of both CTAs initialised before any remote signal
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = tail.tmem_base;

    const int n_items = args.n_qtiles * args.nsplit;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs, own halves) =====================
        if (ptx::elect_one()) {
            const uint32_t a_full_l = ptx::mapa_u32(ptx::smem_u32(&tail.a_full), 0);
            int stage = 0;
            uint32_t ph = 0, a_ph = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int qt = item / args.nsplit, sp = item % args.nsplit;
                int t_lo = sp * args.tiles_per_split;
                int t_hi = t_lo + args.tiles_per_split;
                if (t_hi > args.n_btiles) t_hi = args.n_btiles;
                ptx::mbar_wait(&tail.a_empty, a_ph ^ 1);
                if (leader) ptx::mbar_arrive_expect_tx(&tail.a_full, 2 * a_bytes);
                for (int s = 0; s < plan.nseg; ++s) {
                    const TcSeg &sg = plan.seg[s];
                    if (sg.fold)
                        ptx::tma_load_2d_2sm(smA + sg.a_off / 2, &maps.fa, a_full_l, 0, 0);
                    else
                        ptx::tma_load_2d_2sm(smA + sg.a_off / 2, &maps.a[seg_class(sg.bytes)],
                                             a_full_l, sg.k_off,
                                             args.q_op_begin + qt * TC_ROWS + (int)rank * 128);
                }
                a_ph ^= 1;
                for (int t = t_lo; t < t_hi; ++t) {
                    ptx::mbar_wait(&tail.b_empty[stage], ph ^ 1);
                    if (leader) ptx::mbar_arrive_expect_tx(&tail.b_full[stage], 2 * plan.b_bytes);
                    const uint32_t b_full_l = ptx::mapa_u32(ptx::smem_u32(&tail.b_full[stage]), 0);
                    uint8_t *dst = smB + stage * plan.b_bytes;
                    const int row0 = t * 256 + (int)rank * 128;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        if (sg.fold)
                            ptx::tma_load_2d_2sm(dst + sg.b_off, &maps.fb, b_full_l, 0, row0);
                        else
                            ptx::tma_load_2d_2sm(dst + sg.b_off, &maps.b[seg_class(sg.bytes)],
                                                 b_full_l, sg.k_off, row0);
                    }
                    if (++stage == TC2_NSTAGE) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(256, 256);
            int stage = 0, as = 0;
            uint32_t ph = 0, acc_ph = 0, a_ph = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int sp = item % args.nsplit;
                int t_lo = sp * args.tiles_per_split;
                int t_hi = t_lo + args.tiles_per_split;
                if (t_hi > args.n_btiles) t_hi = args.n_btiles;
                ptx::mbar_wait(&tail.a_full, a_ph);
                a_ph ^= 1;
                for (int t = t_lo; t < t_hi; ++t) {
                    ptx::mbar_wait(&tail.b_full[stage], ph);
                    ptx::mbar_wait(&tail.acc_empty[as], acc_ph ^ 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t b_base = ptx::smem_u32(smB + stage * plan.b_bytes);
                    const uint32_t a_base = ptx::smem_u32(smA);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
                    uint32_t acc = 0;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        const uint32_t a_addr = a_base + sg.a_off / 2;
                        const uint32_t b_addr = b_base + sg.b_off;
                        const uint32_t sbo = 8 * sg.bytes, lay = seg_layout(sg.bytes);
                        for (int kk = 0; kk < sg.bytes / 32; ++kk) {
                            ptx::umma_f16_2sm(d_tmem, ptx::make_smem_desc(a_addr + kk * 32, sbo, lay),
                                              ptx::make_smem_desc(b_addr + kk * 32, sbo, lay), idesc,
                                              acc);
                            acc = 1;
                        }
                    }
                    ptx::umma_commit_2sm(&tail.b_empty[stage]);
                    ptx::umma_commit_2sm(&tail.acc_full[as]);
                    if (++stage == TC2_NSTAGE) {
                        stage = 0;
                        ph ^= 1;
                    }
                    as ^= 1;
                    if (as == 0) acc_ph ^= 1;
                }
                ptx::umma_commit_2sm(&tail.a_empty);
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue: threshold filter =====================
        // four filter warps (one per TMEM lane quadrant); a thread owns one query row and
        // filters all 256 columns of a tile as two 128-column passes
        const int quad = warp & 3;
        const int row_in_cta = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int L = args.sc->inexact ? args.L_inexact : args.L_exact;
        uint32_t acc_empty_l[2];
        acc_empty_l[0] = ptx::mapa_u32(ptx::smem_u32(&tail.acc_empty[0]), 0);
        acc_empty_l[1] = ptx::mapa_u32(ptx::smem_u32(&tail.acc_empty[1]), 0);
        int as = 0;
        uint32_t acc_ph = 0;
        unsigned ctr[4] = {0, 0, 0, 0};
        const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int qt = item / args.nsplit, sp = item % args.nsplit;
            int t_lo = sp * args.tiles_per_split;
            int t_hi = t_lo + args.tiles_per_split;
            if (t_hi > args.n_btiles) t_hi = args.n_btiles;
            const int r = qt * TC_ROWS + (int)rank * 128 + row_in_cta;
            const bool valid = r < args.n_rows;
            const int self = (valid && args.self_join)
                                 ? (args.row_ids ? args.row_ids[r] : args.q_begin + r)
                                 : -1;
            const size_t slot = valid ? (size_t)r * args.nsplit + sp : 0;
            key64_t *list = args.cand + slot * GCAP;
            float tau = valid ? NEG_INF : POS_INF;
            int cnt = 0;
            long long b_hi = (long long)t_hi * 256;
            if (b_hi > args.nb) b_hi = args.nb;
            const int ib_hi = (int)b_hi;

            for (int t = t_lo; t < t_hi; ++t) {
                wait_acc_with_idle_shrinks<GCAP>(&tail.acc_full[as], acc_ph, L, lane, nullptr, tau,
                                                 cnt, list, ctr);
                ptx::tc_fence_after_sync();
                const uint32_t tbase = tmem_base + lane_addr + (uint32_t)(as * 256);
                const int tile0 = t * 256;
                const bool special = (t == t_hi - 1) || (self >= tile0 && self < tile0 + 256);
                auto release = [&]() {
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(acc_empty_l[as]);
                };
                auto keep = []() {};
                if (args.debug_mode != 1) {
                    filter_tile<GCAP>(tbase, tile0, special, ib_hi, self, L, lane,
                                      args.debug_mode, nullptr, tau, cnt, list, ctr, keep);
                    filter_tile<GCAP>(tbase + 128, tile0 + 128, special, ib_hi, self, L, lane,
                                      args.debug_mode, nullptr, tau, cnt, list, ctr, release);
                } else {
                    release();
                }
                ctr[3] += TC_BN / 32;
                as ^= 1;
                if (as == 0) acc_ph ^= 1;
            }
            if (valid) {
                args.cand_cnt[slot] = cnt;
                args.cand_tau[slot] = tau;
            }
        }
        if (lane == 0) {
            PrepScalars *scw = const_cast<PrepScalars *>(args.sc);
            for (int i = 0; i < 4; ++i) atomicAdd(&scw->dbg[i], (unsigned long long)ctr[i]);
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();  // the peer may still be signalling this CTA's barriers / TMEM
    if (warp == 2) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

// ------------------------------------------------------------------------------------
// exact FP32 re-rank + certificate.  One warp per query row.
constexpr int RR_WARPS = 4;

struct RerankArgs {
    const float *base;
    const float *queries;  // row r of this launch is queries[(q_row_begin + r) * d]
    int q_row_begin;
    int q_begin;           // absolute query index of row 0 (self-join id, output naming)
    const int *row_ids;    // second pass: absolute query index of every row instead
    int out_q_begin;       // output row = q - out_q_begin
    int n_rows, nb, d, k, kcap, self_join;
    int nlists, gcap;  // candidate lists per row (base splits), rows before the tail
    int tail_row0, ns_tail;  // rows >= tail_row0 own ns_tail lists (knn_tc_common.cuh list_slot)
    const key64_t *cand;
    const int *cand_cnt;
    const float *cand_tau;
    const double *qnorm64;  // |x|^2 of the query rows, indexed like queries (q_row_begin + r)
    const double *bnorm64;  // |y|^2 of the base rows
    int check_eps;          // debug: measure |S'_tensor - S'_fp64| / eps over all candidates
    const PrepScalars *sc;
    int ksteps;             // UMMA K steps per accumulator (for the accumulation bound)
    int i8;                 // scores came from the integer path (knn_tc_i8.cuh)
    int *assign;
    float *dis;
    int *fb_rows;           // absolute query ids of rows that failed the certificate
    uint32_t *fb_count;
};

template <int NMAX>
__global__ void __launch_bounds__(RR_WARPS * 32) knn_rerank_kernel(const RerankArgs a) {
    extern __shared__ __align__(16) unsigned char rr_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    key64_t *keys = reinterpret_cast<key64_t *>(rr_smem) + (size_t)warp * NMAX;
    float *qrow = reinterpret_cast<float *>(rr_smem + (size_t)RR_WARPS * NMAX * sizeof(key64_t)) +
                  (size_t)warp * a.d;
    const int r = blockIdx.x * RR_WARPS + warp;
    if (r >= a.n_rows) return;
    const int q = a.row_ids ? a.row_ids[r] : a.q_begin + r;

    // gather the candidate ids of all splits; tau = max over splits
    int n = 0;
    float tau = __int_as_float(0xff800000);
    const int my_lists = r < a.tail_row0 ? a.nlists : a.ns_tail;
    for (int s = 0; s < my_lists; ++s) {
        const size_t slot = list_slot(r, s, a.tail_row0, a.nlists, a.ns_tail);
        const int c = a.cand_cnt[slot];
        const key64_t *l = a.cand + slot * a.gcap;
        for (int i = lane; i < c; i += 32) keys[n + i] = __ldcg(l + i);
        n += c;
        tau = fmaxf(tau, a.cand_tau[slot]);
    }
    const float *qsrc = a.queries + (size_t)(a.q_row_begin + r) * a.d;
    for (int i = lane; i < a.d; i += 32) qrow[i] = qsrc[i];
    __syncwarp();

    // error bound of the tensor-core score for this row (see the certificate below)
    double eps = 0.0;
    const double sigma = (double)scale_from_maxabs_bits(a.sc->maxabs_bits);
    const double xn2 = a.qnorm64[a.q_row_begin + r];
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
            atomicMax(const_cast<uint32_t *>(&a.sc->pad[0]), __float_as_uint(worst));
    }

    // exact distances: the scalar sequence of get_real_dist, one candidate per lane
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
        const char *e = getenv("PQH_TC_LEXACT");
        if (e) return atoi(e);
        return (kcap + kcap / 2 + 7) / 8 * 8;
    }
    if (const char *e = getenv("PQH_TC_LINEXACT")) return atoi(e);
    // rounded operands: 64 candidates, or twice the neighbours asked for if that is more.
    // Measured (profiles/README.md): GIST-shaped k=16 needs 64 (at 32, 10 % of the rows fail the
    // certificate); Deep-shaped k=32 certifies every row at 64 and runs 1.5x faster than at 128.
    int L = 2 * kcap;
    L = (L + 31) / 32 * 32;
    if (L < 64) L = 64;
    return L;
}
// list capacity: the filter checks it once per tile, after up to TC_BN (128) appends
int pick_gcap(int L) {
    static const char *e = getenv("PQH_TC_GCAP");  // experiments: 256 or 512
    if (e && (atoi(e) == 256 || atoi(e) == 512) && L <= 96) return atoi(e);
    return L <= 96 ? 256 : 512;
}

template <int GCAP>
size_t tc2_smem_bytes(const TcPlan &p) {
    return 1024 + (size_t)p.a_bytes / 2 + (size_t)TC2_NSTAGE * p.b_bytes +
           sizeof(Tc2SmemTail<GCAP>);
}
template <int GCAP>
size_t tc_smem_bytes(const TcPlan &p) {
    return 1024 + (size_t)p.a_bytes + (size_t)TC_NSTAGE * p.b_bytes + sizeof(TcSmemTail<GCAP>);
}

struct TcLayout {
    size_t xh_b, xh_q, fold_b, fold_a, norm_b, norm_q, scalars, cand, cnt, tau, fb_rows, total;
    size_t a8;  // integer path: u8 query operand [n_rows x d32] (0 when the path is disabled)
    int d32;
    // second pass (rows that failed the first certificate, redone with longer lists)
    size_t p2_q, p2_xh, p2_norm, p2_cand, p2_cnt, p2_tau, p2_fb;
    int p2_rows;  // rows per second-pass batch
    int d16, L, L_exact, gcap, nsplit, n_qtiles, n_btiles, tiles_per_split;
    int variant, halves, nlists;
    int q_tail0, ns_tail, tiles_per_split_tail, qrows;  // tail of the work list (TcArgs)
    size_t n_slots;                                      // candidate lists in all
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Kernel variants.  1 = single-CTA kernel with the query tile resident in shared memory
// (knn_tc_kernel; default whenever the tile fits, d <= ~144); 4 = streaming kernel for larger d
// (knn_tcs.cu).  PQH_TC_VARIANT=2 / 3 select the CTA-pair experiments (this file / knn_tc3.cu),
// kept for A/B measurements -- see profiles/README.md for why they are not the default.
int tc_variant() {
    const char *v = getenv("PQH_TC_VARIANT");
    const int x = v ? atoi(v) : 1;
    return (x >= 1 && x <= 5) ? x : 1;
}

// does the resident-A plan of `variant` fit in shared memory for this d16 / list capacity?
bool resident_fits(int variant, int d16, int gcap) {
    TcPlan p;
    if (!build_plan(d16, &p)) return false;
    if (variant == 3) return tc3_pick_nstage(&p) >= 2;
    if (variant == 5) return tc5_pick_nstage(&p, d16) >= 3;
    size_t smem;
    if (variant == 2)
        smem = gcap == 128 ? tc2_smem_bytes<128>(p)
                           : (gcap == 256 ? tc2_smem_bytes<256>(p) : tc2_smem_bytes<512>(p));
    else
        smem = gcap == 128 ? tc_smem_bytes<128>(p)
                           : (gcap == 256 ? tc_smem_bytes<256>(p) : tc_smem_bytes<512>(p));
    return smem <= 227 * 1024;
}

int tc_pick_variant(int d16, int gcap) {
    int v = tc_variant();
    if (v != 4 && !resident_fits(v, d16, gcap)) v = 4;
    return v;
}

constexpr int P2_GCAP = 512, P2_MAX_LISTS = 4;

// Integer path for data made of integers in [0, 255] (knn_tc_i8.cuh): taken whenever the data
// qualifies; $PQH_TC_I8=0 keeps such data on the FP16 path (A/B measurements, tests of the FP16
// error bound).  Read on every call so that one process can compare both.
bool i8_enabled() {
    const char *e = getenv("PQH_TC_I8");
    return e ? atoi(e) != 0 : true;
}

void tc_layout(int n_rows, int nq_src_rows, int nb, int d, int k, int self_join, int num_sms,
               TcLayout *lo, int L_override = 0) {
    const int kcap = self_join ? k - 1 : k;
    lo->d16 = (d + 15) / 16 * 16;
    lo->L = L_override ? L_override : pick_L(kcap, true);
    lo->L_exact = L_override ? L_override : pick_L(kcap, false);
    lo->gcap = L_override ? P2_GCAP : pick_gcap(lo->L);
    lo->variant = tc_pick_variant(lo->d16, lo->gcap);
    lo->halves = 1;
    const int bn = lo->variant == 2 ? 256 : (lo->variant == 5 ? tc5_tile_rows() : TC_BN);
    const int units = lo->variant == 1 ? num_sms : num_sms / 2;
    const int qrows = lo->variant == 3 ? 512 : TC_ROWS;  // query rows per work item
    lo->n_qtiles = (n_rows + qrows - 1) / qrows;
    lo->n_btiles = (nb + bn - 1) / bn;
    // Base splits: a work item is (query tile, base split).  A split costs a second candidate list
    // per row with its own warm-up (measured: 100k x 128 takes 5.0 / 6.7 / 8.6 ms with 1 / 2 / 4
    // splits although 4 splits fill the last round of the grid best), so the base is only split
    // when the query tiles alone cannot occupy every SM once.
    const int max_lists = 2048 / lo->gcap;
    int nsplit = 1;
    while (lo->n_qtiles * nsplit < units && nsplit * 2 * lo->halves <= max_lists && nsplit < 4 &&
           lo->n_btiles / (nsplit * 2) >= 16)
        nsplit *= 2;
    if (const char *e = getenv("PQH_TC_NSPLIT")) {
        const int v = atoi(e);
        if (v >= 1 && v * lo->halves <= max_lists && lo->n_btiles / v >= 8) nsplit = v;
    }
    lo->nsplit = nsplit;
    lo->nlists = nsplit * lo->halves;
    lo->tiles_per_split = (lo->n_btiles + nsplit - 1) / nsplit;
    // The last round of the persistent grid: if the query tiles left over after the full rounds
    // occupy at most half of the SMs, cut each of them into 2..4 base splits so the round
    // finishes in a fraction of an item time (8 GPUs x 125k rows: 489 tiles = 3 rounds + 45).
    lo->qrows = qrows;
    lo->q_tail0 = lo->n_qtiles;
    lo->ns_tail = 1;
    lo->tiles_per_split_tail = lo->n_btiles;
    if ((lo->variant == 1 || lo->variant == 4) && nsplit == 1 && lo->n_qtiles > units &&
        !getenv("PQH_TC_NO_TAIL")) {
        const int rem = lo->n_qtiles % units;
        int ns = rem > 0 ? units / rem : 1;
        if (ns > 4) ns = 4;
        while (ns > 1 && (ns > max_lists || lo->n_btiles / ns < 16)) --ns;
        if (ns > 1) {
            lo->q_tail0 = lo->n_qtiles - rem;
            lo->ns_tail = ns;
            lo->tiles_per_split_tail = (lo->n_btiles + ns - 1) / ns;
        }
    }
    {
        const long long tail_row0 = (long long)lo->q_tail0 * qrows;
        const long long rows_main = tail_row0 < n_rows ? tail_row0 : n_rows;
        lo->n_slots = (size_t)rows_main * lo->nlists + (size_t)(n_rows - rows_main) * lo->ns_tail;
    }
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    lo->scalars = take(sizeof(PrepScalars));
    lo->xh_b = take((size_t)nb * lo->d16 * sizeof(__half));
    lo->xh_q = self_join ? lo->xh_b : take((size_t)nq_src_rows * lo->d16 * sizeof(__half));
    lo->fold_b = take((size_t)nb * FOLD_COLS * sizeof(__half));
    lo->fold_a = take((size_t)TC_ROWS * FOLD_COLS * sizeof(__half));
    lo->norm_b = take((size_t)nb * sizeof(double));
    lo->norm_q = self_join ? lo->norm_b : take((size_t)nq_src_rows * sizeof(double));
    lo->cand = take(lo->n_slots * lo->gcap * sizeof(key64_t));
    lo->cnt = take(lo->n_slots * sizeof(int));
    lo->tau = take(lo->n_slots * sizeof(float));
    lo->fb_rows = take((size_t)n_rows * sizeof(int));
    lo->d32 = (d + 31) / 32 * 32;
    lo->a8 = i8_enabled() ? take((size_t)n_rows * lo->d32) : 0;
    {
        int p2 = n_rows / 16;
        if (p2 < 2048) p2 = 2048;
        if (p2 > n_rows) p2 = n_rows;
        lo->p2_rows = p2;
        lo->p2_q = take((size_t)p2 * d * sizeof(float));
        lo->p2_xh = take((size_t)p2 * lo->d16 * sizeof(__half));
        lo->p2_norm = take((size_t)p2 * sizeof(double));
        lo->p2_cand = take((size_t)p2 * P2_MAX_LISTS * P2_GCAP * sizeof(key64_t));
        lo->p2_cnt = take((size_t)p2 * P2_MAX_LISTS * sizeof(int));
        lo->p2_tau = take((size_t)p2 * P2_MAX_LISTS * sizeof(float));
        lo->p2_fb = take((size_t)n_rows * sizeof(int));
    }
    lo->total = off + 1024;
}

}  // namespace

bool tc_supported(int n_rows, int nb, int d, int k, int self_join) {
    const int kcap = self_join ? k - 1 : k;
    if (kcap < 1 || kcap > 64) return false;  // list capacity 512 holds L = 128 candidates
    if (nb < 2048 || n_rows < 1) return false;
    if (d < 1 || d > 4096) return false;
    return true;  // the streaming kernel takes every d the resident kernels cannot
}

size_t tc_workspace_bytes(int n_rows, int nb, int d, int k, int self_join, int num_sms) {
    TcLayout lo;
    tc_layout(n_rows, n_rows, nb, d, k, self_join, num_sms, &lo);
    return lo.total;
}

// Tensor-core search of query rows [q_begin, q_begin + n_rows).
//   d_queries: self-join -> same matrix as d_base (row q is query q);
//              otherwise a matrix whose row (q - q_row0) is query q.
// Rows that fail the certificate are written to *h_fb_count / d_fb_rows (absolute ids);
// the caller redoes them with the exact kernel.  Synchronises the stream once.
cudaError_t launch_knn_tc(const float *d_base, const float *d_queries, int q_row0, int q_begin,
                          int out_q_begin, int n_rows, int nb, int d, int k, int self_join,
                          int *d_assign, float *d_dis, void *d_work, size_t work_bytes,
                          int num_sms, cudaStream_t stream, int *launches, int **d_fb_rows_out,
                          unsigned *h_fb_count, cudaEvent_t *ev, TcTimes *times) {
    *h_fb_count = 0;
    if (n_rows <= 0) return cudaSuccess;
    const int kcap = self_join ? k - 1 : k;
    TcLayout lo;
    tc_layout(n_rows, n_rows, nb, d, k, self_join, num_sms, &lo);
    if (lo.total > work_bytes) return cudaErrorMemoryAllocation;
    unsigned char *w = reinterpret_cast<unsigned char *>(
        ((uintptr_t)d_work + 1023) & ~(uintptr_t)1023);
    PrepScalars *sc = reinterpret_cast<PrepScalars *>(w + lo.scalars);
    __half *xh_b = reinterpret_cast<__half *>(w + lo.xh_b);
    __half *xh_q = reinterpret_cast<__half *>(w + lo.xh_q);
    __half *fold_b = reinterpret_cast<__half *>(w + lo.fold_b);
    __half *fold_a = reinterpret_cast<__half *>(w + lo.fold_a);
    double *norm_b = reinterpret_cast<double *>(w + lo.norm_b);
    double *norm_q = reinterpret_cast<double *>(w + lo.norm_q);
    key64_t *cand = reinterpret_cast<key64_t *>(w + lo.cand);
    int *cnt = reinterpret_cast<int *>(w + lo.cnt);
    float *tau = reinterpret_cast<float *>(w + lo.tau);
    int *fb_rows = reinterpret_cast<int *>(w + lo.fb_rows);
    *d_fb_rows_out = fb_rows;

    cudaError_t e;
    if (ev) cudaEventRecord(ev[0], stream);
    if ((e = cudaMemsetAsync(sc, 0, sizeof(PrepScalars), stream)) != cudaSuccess) return e;

    // queries of this launch: rows [q_src0, q_src0 + n_rows) of d_queries
    const int q_src0 = self_join ? q_begin : q_begin - q_row0;
    const char *dbg_env = getenv("PQH_TC_DEBUG");
    const int debug_mode = dbg_env ? atoi(dbg_env) : 0;

    // ---- prep, integer flavour (opt-in): tried first, kept only if every value qualifies ----
    bool used_i8 = false;
    uint8_t *a8 = reinterpret_cast<uint8_t *>(w + lo.a8);
    if (i8_enabled() && lo.variant == 1 && debug_mode == 0 &&
        resident_fits(1, lo.d32 / 2, lo.gcap)) {
        const float *qsrc = (self_join ? d_base : d_queries) + (size_t)q_src0 * d;
        prep_rows_i8_kernel<<<(nb + 7) / 8, 256, 0, stream>>>(
            d_base, nb, d, lo.d32, reinterpret_cast<uint8_t *>(xh_b), norm_b,
            reinterpret_cast<int8_t *>(fold_b), 1, sc);
        prep_rows_i8_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(
            qsrc, n_rows, d, lo.d32, a8, self_join ? nullptr : norm_q, nullptr, 0, sc);
        fold_const_i8_kernel<<<(TC_ROWS * 32 + 255) / 256, 256, 0, stream>>>(
            reinterpret_cast<uint8_t *>(fold_a));
        *launches += 3;
        PrepScalars probe;
        if ((e = cudaMemcpyAsync(&probe, sc, sizeof(PrepScalars), cudaMemcpyDeviceToHost,
                                 stream)) != cudaSuccess)
            return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        used_i8 = probe.pad[2] == 0;
    }

    // ---- prep, FP16 flavour ----
    auto prep_fp16 = [&]() {
        const int grid_ma = num_sms * 8;
        maxabs_kernel<<<grid_ma, 256, 0, stream>>>(d_base, (size_t)nb * d, &sc->maxabs_bits);
        ++*launches;
        if (!self_join) {
            maxabs_kernel<<<grid_ma, 256, 0, stream>>>(d_queries + (size_t)q_src0 * d,
                                                       (size_t)n_rows * d, &sc->maxabs_bits);
            ++*launches;
        }
        prep_rows_kernel<<<(nb + 7) / 8, 256, 0, stream>>>(d_base, nb, d, lo.d16, xh_b, norm_b,
                                                           fold_b, 1, sc);
        ++*launches;
        if (!self_join) {
            prep_rows_kernel<<<(n_rows + 7) / 8, 256, 0, stream>>>(d_queries + (size_t)q_src0 * d,
                                                                   n_rows, d, lo.d16, xh_q, norm_q,
                                                                   nullptr, 0, sc);
            ++*launches;
        }
        fold_const_kernel<<<(TC_ROWS * FOLD_COLS + 255) / 256, 256, 0, stream>>>(fold_a);
        ++*launches;
    };
    if (!used_i8) prep_fp16();

    // ---- one pass = tensor-core kernel + re-rank over a set of query rows ----
    struct Pass {
        const TcLayout *lo;      // geometry (tiles, splits, list sizes) of this pass
        const __half *xh_q;      // FP16 operand whose row (q_op_begin + r) is query r
        uint64_t a_rows;         // rows of that operand
        int q_op_begin;
        const float *q_mat;      // FP32 matrix whose row (q_row_begin + r) is query r
        int q_row_begin;
        const double *qnorm64;   // indexed like q_mat
        const int *row_ids;      // absolute query ids (second pass) or null
        int n_rows;
        key64_t *cand;
        int *cnt;
        float *tau;
        int *fb_rows;
        uint32_t *fb_count;
        bool i8;                 // integer operands (first pass only)
        const void *b_op;        // base operand rows
        int op_cols;             // operand row length in 2-byte units (the tensor maps' view)
    };
    auto run_pass = [&](const Pass &p) -> cudaError_t {
        const TcLayout &pl = *p.lo;
        TcMaps maps;
        TcArgs args;
        if (!build_plan(p.op_cols, &args.plan) && pl.variant != 4) return cudaErrorInvalidValue;
        const uint32_t boxc[3] = {64, 32, 16};
        const uint32_t a_box_rows = pl.variant == 2 ? 128 : TC_ROWS;
        const uint32_t b_box_rows =
            pl.variant == 3 ? 64 : (pl.variant == 5 ? tc5_tile_rows() : TC_BN);
        for (int c = 0; c < 3; ++c) {
            // classes a small d never uses still need a legal box
            const uint32_t bc = boxc[c] <= (uint32_t)p.op_cols ? boxc[c] : 16;
            if (!make_map(&maps.a[c], p.xh_q, p.a_rows, p.op_cols, bc, a_box_rows) ||
                !make_map(&maps.b[c], p.b_op, nb, p.op_cols, bc, b_box_rows))
                return cudaErrorInvalidValue;
        }
        if (!make_map(&maps.fa, fold_a, TC_ROWS, FOLD_COLS, FOLD_COLS, a_box_rows) ||
            !make_map(&maps.fb, fold_b, nb, FOLD_COLS, FOLD_COLS, b_box_rows))
            return cudaErrorInvalidValue;

        args.n_rows = p.n_rows;
        args.q_begin = q_begin;
        args.q_op_begin = p.q_op_begin;
        args.nb = nb;
        args.self_join = self_join;
        args.row_ids = p.row_ids;
        args.n_qtiles = pl.n_qtiles;
        args.nsplit = pl.nsplit;
        args.q_tail0 = pl.q_tail0;
        args.ns_tail = pl.ns_tail;
        args.tiles_per_split_tail = pl.tiles_per_split_tail;
        args.tiles_per_split = pl.tiles_per_split;
        args.n_btiles = pl.n_btiles;
        args.debug_mode = debug_mode;
        args.nstage = 0;
        args.kcap = kcap;
        args.L_exact = pl.L_exact;
        args.L_inexact = pl.L;
        args.sc = sc;
        args.cand = p.cand;
        args.cand_cnt = p.cnt;
        args.cand_tau = p.tau;

        const int n_items = pl.q_tail0 * pl.nsplit + (pl.n_qtiles - pl.q_tail0) * pl.ns_tail;
        size_t smem;
        cudaError_t pe;
#define LAUNCH_TC(KERNEL, SMEM, GRID)                                                            \
    do {                                                                                         \
        smem = (SMEM);                                                                           \
        if ((pe = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                       (int)smem)) != cudaSuccess)                               \
            return pe;                                                                           \
        KERNEL<<<(GRID), TC_THREADS, smem, stream>>>(maps, args);                                \
    } while (0)
        if (pl.variant == 5) {
            args.nstage = tc5_pick_nstage(&args.plan, pl.d16);
            if (args.nstage < 3) return cudaErrorInvalidValue;
            if ((pe = launch_tc5(&maps, &args, p.xh_q, fold_a, pl.gcap, pl.d16, n_items, num_sms,
                                 stream)) != cudaSuccess)
                return pe;
        } else if (pl.variant == 4) {
            if ((pe = launch_tcs(&maps, &args, pl.gcap, pl.d16, n_items, num_sms, stream)) !=
                cudaSuccess)
                return pe;
        } else if (pl.variant == 3) {
            args.nstage = tc3_pick_nstage(&args.plan);
            if (args.nstage < 2) return cudaErrorInvalidValue;
            if ((pe = launch_tc3(&maps, &args, pl.gcap, n_items, num_sms, stream)) != cudaSuccess)
                return pe;
        } else if (pl.variant == 2) {
            const int clusters = n_items < num_sms / 2 ? n_items : num_sms / 2;
            const int grid = 2 * clusters;
            if (pl.gcap == 256) LAUNCH_TC(knn_tc2_kernel<256>, tc2_smem_bytes<256>(args.plan), grid);
            else LAUNCH_TC(knn_tc2_kernel<512>, tc2_smem_bytes<512>(args.plan), grid);
        } else {
            const int grid = n_items < num_sms ? n_items : num_sms;
            // the PROF instantiations carry the clock64()/counter instrumentation (PQH_TC_DEBUG)
            if (debug_mode != 0) {
                if (pl.gcap == 256)
                    LAUNCH_TC((knn_tc_kernel<256, true, false>), tc_smem_bytes<256>(args.plan), grid);
                else
                    LAUNCH_TC((knn_tc_kernel<512, true, false>), tc_smem_bytes<512>(args.plan), grid);
            } else if (p.i8) {
                if (pl.gcap == 256)
                    LAUNCH_TC((knn_tc_kernel<256, false, true>), tc_smem_bytes<256>(args.plan), grid);
                else
                    LAUNCH_TC((knn_tc_kernel<512, false, true>), tc_smem_bytes<512>(args.plan), grid);
            } else if (pl.gcap == 256) {
                LAUNCH_TC((knn_tc_kernel<256, false, false>), tc_smem_bytes<256>(args.plan), grid);
            } else {
                LAUNCH_TC((knn_tc_kernel<512, false, false>), tc_smem_bytes<512>(args.plan), grid);
            }
        }
#undef LAUNCH_TC
        ++*launches;
        if ((pe = cudaGetLastError()) != cudaSuccess) return pe;
        if (ev && !p.row_ids) cudaEventRecord(ev[2], stream);

        // ---- re-rank + certificate ----
        RerankArgs ra;
        ra.base = d_base;
        ra.queries = p.q_mat;
        ra.q_row_begin = p.q_row_begin;
        ra.q_begin = q_begin;
        ra.row_ids = p.row_ids;
        ra.out_q_begin = out_q_begin;
        ra.n_rows = p.n_rows;
        ra.nb = nb;
        ra.d = d;
        ra.k = k;
        ra.kcap = kcap;
        ra.self_join = self_join;
        ra.nlists = pl.nlists;
        ra.tail_row0 = pl.q_tail0 * pl.qrows;
        ra.ns_tail = pl.ns_tail;
        ra.gcap = pl.gcap;
        ra.cand = p.cand;
        ra.cand_cnt = p.cnt;
        ra.cand_tau = p.tau;
        ra.qnorm64 = p.qnorm64;
        ra.bnorm64 = norm_b;
        ra.check_eps = getenv("PQH_TC_CHECK_EPS") ? 1 : 0;
        ra.sc = sc;
        ra.ksteps = (pl.d16 + FOLD_COLS) / 16;
        ra.i8 = p.i8 ? 1 : 0;
        ra.assign = d_assign;
        ra.dis = d_dis;
        ra.fb_rows = p.fb_rows;
        ra.fb_count = p.fb_count;
        const int nmax = pl.gcap * (pl.nlists > pl.ns_tail ? pl.nlists : pl.ns_tail);
        const int rgrid = (p.n_rows + RR_WARPS - 1) / RR_WARPS;
        const size_t qbytes = (size_t)RR_WARPS * d * sizeof(float);
#define LAUNCH_RR(N)                                                                             \
    do {                                                                                         \
        size_t sm = (size_t)RR_WARPS * N * sizeof(key64_t) + qbytes;                             \
        if ((pe = cudaFuncSetAttribute(knn_rerank_kernel<N>,                                     \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)) != \
            cudaSuccess)                                                                         \
            return pe;                                                                           \
        knn_rerank_kernel<N><<<rgrid, RR_WARPS * 32, sm, stream>>>(ra);                          \
    } while (0)
        if (nmax <= 256) LAUNCH_RR(256);
        else if (nmax <= 512) LAUNCH_RR(512);
        else if (nmax <= 1024) LAUNCH_RR(1024);
        else LAUNCH_RR(2048);
#undef LAUNCH_RR
        ++*launches;
        return cudaGetLastError();
    };

    // ---- first pass: every row of this launch ----
    if (ev) cudaEventRecord(ev[1], stream);
    Pass p1;
    p1.lo = &lo;
    p1.i8 = used_i8;
    p1.b_op = xh_b;
    p1.op_cols = used_i8 ? lo.d32 / 2 : lo.d16;
    p1.xh_q = used_i8 ? reinterpret_cast<const __half *>(a8) : xh_q;
    p1.a_rows = (self_join && !used_i8) ? (uint64_t)nb : (uint64_t)n_rows;
    p1.q_op_begin = (self_join && !used_i8) ? q_begin : 0;
    p1.q_mat = self_join ? d_base : d_queries;
    p1.q_row_begin = q_src0;
    p1.qnorm64 = self_join ? norm_b : norm_q - q_src0;  // indexed by q_row_begin + r
    p1.row_ids = nullptr;
    p1.n_rows = n_rows;
    p1.cand = cand;
    p1.cnt = cnt;
    p1.tau = tau;
    p1.fb_rows = fb_rows;
    p1.fb_count = &sc->fb_count;
    if ((e = run_pass(p1)) != cudaSuccess) return e;

    if (ev) cudaEventRecord(ev[3], stream);
    PrepScalars hsc;
    if ((e = cudaMemcpyAsync(&hsc, sc, sizeof(PrepScalars), cudaMemcpyDeviceToHost, stream)) !=
        cudaSuccess)
        return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    *h_fb_count = hsc.fb_count;

    // ---- second pass: rows whose certificate failed, redone with four times longer lists.
    //      Their cost on the exact kernel would be ~1000x a tensor-core row, so a data set on
    //      which a few per cent of the rows fail must not fall off a cliff.  Skipped when most
    //      rows fail (FP16 cannot resolve this data at all) or when there are only a handful.
    const unsigned n_fb1 = hsc.fb_count;
    if (n_fb1 >= 64 && n_fb1 <= (unsigned)n_rows / 2 && !getenv("PQH_TC_NO_PASS2") &&
        debug_mode == 0) {
        const int L1 = hsc.inexact ? lo.L : lo.L_exact;
        int L2 = 4 * L1;
        if (L2 > 192) L2 = 192;
        if (L2 > L1) {
            if (used_i8) {  // the second pass runs on FP16 operands: build them now
                if ((e = cudaMemsetAsync(&sc->maxabs_bits, 0, 3 * sizeof(uint32_t), stream)) !=
                    cudaSuccess)
                    return e;
                prep_fp16();
            }
            float *q2 = reinterpret_cast<float *>(w + lo.p2_q);
            __half *xh2 = reinterpret_cast<__half *>(w + lo.p2_xh);
            double *norm2 = reinterpret_cast<double *>(w + lo.p2_norm);
            int *fb2 = reinterpret_cast<int *>(w + lo.p2_fb);
            uint32_t *fb2_count = &sc->pad[1];
            const float *src = self_join ? d_base : d_queries;
            const int src_row0 = self_join ? 0 : q_row0;
            for (unsigned done = 0; done < n_fb1; done += (unsigned)lo.p2_rows) {
                const int n2 = (int)(n_fb1 - done < (unsigned)lo.p2_rows ? n_fb1 - done
                                                                          : (unsigned)lo.p2_rows);
                const int *ids = fb_rows + done;
                TcLayout lo2;
                tc_layout(n2, n2, nb, d, k, self_join, num_sms, &lo2, L2);
                if (lo2.nlists > P2_MAX_LISTS || lo2.ns_tail > P2_MAX_LISTS)
                    return cudaErrorInvalidValue;
                gather_rows_kernel<<<(n2 + 7) / 8, 256, 0, stream>>>(src, src_row0, ids, n2, d, q2);
                prep_rows_kernel<<<(n2 + 7) / 8, 256, 0, stream>>>(q2, n2, d, lo.d16, xh2, norm2,
                                                                   nullptr, 0, sc);
                *launches += 2;
                if ((e = cudaMemsetAsync(&sc->tile_cursor, 0, sizeof(uint32_t), stream)) !=
                    cudaSuccess)
                    return e;
                Pass p2;
                p2.lo = &lo2;
                p2.xh_q = xh2;
                p2.a_rows = (uint64_t)n2;
                p2.q_op_begin = 0;
                p2.q_mat = q2;
                p2.q_row_begin = 0;
                p2.qnorm64 = norm2;
                p2.row_ids = ids;
                p2.n_rows = n2;
                p2.cand = reinterpret_cast<key64_t *>(w + lo.p2_cand);
                p2.cnt = reinterpret_cast<int *>(w + lo.p2_cnt);
                p2.tau = reinterpret_cast<float *>(w + lo.p2_tau);
                p2.fb_rows = fb2;
                p2.fb_count = fb2_count;
                p2.i8 = false;
                p2.b_op = xh_b;
                p2.op_cols = lo.d16;
                if ((e = run_pass(p2)) != cudaSuccess) return e;
            }
            if ((e = cudaMemcpyAsync(&hsc, sc, sizeof(PrepScalars), cudaMemcpyDeviceToHost,
                                     stream)) != cudaSuccess)
                return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            *h_fb_count = hsc.pad[1];
            *d_fb_rows_out = fb2;
            if (times) times->rows_pass2 = (long long)n_fb1;
        }
    }

    if (times) {
        times->candidates = 0;
        for (int i = 0; i < 32; ++i) times->candidates += (double)hsc.cand_total[i];
        times->inexact = used_i8 ? 2 : (int)hsc.inexact;
        {
            union { uint32_t u; float f; } cv;
            cv.u = hsc.pad[0];
            times->eps_ratio_max = cv.f;
        }
        if (dbg_env) {
            fprintf(stderr,
                    "tc dbg: slow_tiles=%llu hard_shrinks=%llu idle_shrinks=%llu chunks=%llu | "
                    "epi warps: wait %.3g work %.3g cyc | mma thread: wait_b %.3g wait_acc %.3g "
                    "total %.3g cyc (sums over CTAs) | pass-1 failures %u\n",
                    hsc.dbg[0], hsc.dbg[1], hsc.dbg[2], hsc.dbg[3], (double)hsc.dbg[4],
                    (double)hsc.dbg[5], (double)hsc.dbg[6], (double)hsc.dbg[7],
                    (double)hsc.cand_total[63], n_fb1);
            fprintf(stderr,
                    "tc dbg: filter_tile cycles: load+maxima %.3g release(no hit) %.3g slow %.3g "
                    "hard-shrink %.3g\n",
                    (double)hsc.cand_total[40], (double)hsc.cand_total[41],
                    (double)hsc.cand_total[42], (double)hsc.cand_total[43]);
        }
        if (ev) {
            cudaEventElapsedTime(&times->ms_prep, ev[0], ev[1]);
            cudaEventElapsedTime(&times->ms_gemm, ev[1], ev[2]);
            cudaEventElapsedTime(&times->ms_rerank, ev[2], ev[3]);
        }
    }
    return cudaSuccess;
}

}  // namespace pqh
