// knn_tc5.cu -- distance-GEMM + filter kernel with the query (A) operand in TENSOR MEMORY.
//
// The single-CTA kernel of knn_tc.cu feeds both MMA operands from shared memory: an M=128,
// N=128, K=16 step reads 4 KB of A and 4 KB of B in 64 cycles -- all 128 B/cycle the shared
// memory has -- so the tensor pipe tops out near 58 % (measured, profiles/README.md).  Here the
// query tile lives in TMEM instead (tcgen05.mma with the A operand in tensor memory: lane =
// query row, two consecutive K elements per 32-bit column), written once per work item by the
// epilogue threads with tcgen05.st, and shared memory only carries the streamed base tiles:
// half the shared-memory traffic per MMA, no A region in shared memory (a deeper B ring), and
// the same thread-per-row epilogue.
//
// TMEM budget (512 columns): 2 accumulator stages x 2 query tiles x 80 columns = 320, plus
// 2 x 96 columns for the two 128-row query tiles in FP16 (up to K = 192) -> base tiles of 80 rows.
#include <cuda_fp16.h>

#include <cstdlib>

#include "kernels.h"
#include "knn_tc_common.cuh"

namespace pqh {

namespace {

constexpr int T5_BN = 80;        // base rows per tile (UMMA N)
constexpr int T5_MAXSTAGE = 10;
constexpr int T5_A_COLS = 96;    // TMEM columns reserved per query tile
constexpr int T5_ACC0 = 0;       // accumulators: column (stage * 2 + g) * T5_BN
constexpr int T5_A0 = 320;       // query tiles: column T5_A0 + g * T5_A_COLS

struct T5SmemTail {
    uint64_t b_full[T5_MAXSTAGE], b_empty[T5_MAXSTAGE];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t a_ready, a_free;  // epilogue -> MMA: query tile stored; MMA -> epilogue: item done
    uint64_t t0_ready;
    int item_t0[2];
    uint32_t tmem_base;
};

template <int GCAP>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc5_kernel(const __grid_constant__ TcMaps maps, const TcArgs args, const __half *xh_q,
               const __half *fold_a, int d16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcPlan &plan = args.plan;  // b_off/b_bytes are for 128-row tiles: scale by 80/128
    // stage stride rounded up to 1 KB: SWIZZLE_128B tiles must start 1024-byte aligned
    const int b_bytes = (plan.b_bytes / TC_BN * T5_BN + 1023) & ~1023;
    const int b_tx = plan.b_bytes / TC_BN * T5_BN;  // bytes TMA actually delivers per stage
    const int nstage = args.nstage;
    uint8_t *smB = smem;
    T5SmemTail &tail = *reinterpret_cast<T5SmemTail *>(smB + nstage * b_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 3; ++i) ptx::prefetch_tmap(&maps.b[i]);
        ptx::prefetch_tmap(&maps.fb);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < T5_MAXSTAGE; ++i) {
            ptx::mbar_init(&tail.b_full[i], 1);
            ptx::mbar_init(&tail.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tail.acc_full[i], 1);
            ptx::mbar_init(&tail.acc_empty[i], TC_EPI_WARPS);
        }
        ptx::mbar_init(&tail.a_ready, TC_EPI_WARPS);
        ptx::mbar_init(&tail.a_free, 1);
        ptx::mbar_init(&tail.t0_ready, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(&tail.tmem_base);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = tail.tmem_base;

    const int n_items = args.n_qtiles * args.nsplit;
    const int ksteps = (d16 + FOLD_COLS) / 16;  // UMMA K steps per accumulator

    if (warp == 0) {
        // ===================== TMA producer: base tiles only =====================
        if (ptx::elect_one()) {
            int stage = 0, it_par = 0, my_next = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int sp = item % args.nsplit;
                int t_lo = sp * args.tiles_per_split;
                int t_hi = t_lo + args.tiles_per_split;
                if (t_hi > args.n_btiles) t_hi = args.n_btiles;
                const int t0 = item_start_tile(args.sc, t_lo, t_hi, my_next);
                my_next = t0;
                tail.item_t0[it_par] = t0;
                ptx::mbar_arrive(&tail.t0_ready);
                it_par ^= 1;
                for (int i = 0; i < t_hi - t_lo; ++i) {
                    const int t = item_tile(i, t0, t_lo, t_hi);
                    if ((i & 7) == 0) publish_cursor(args.sc, t);
                    ptx::mbar_wait(&tail.b_empty[stage], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&tail.b_full[stage], b_tx);
                    uint8_t *dst = smB + stage * b_bytes;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        uint8_t *p = dst + sg.b_off / TC_BN * T5_BN;
                        if (sg.fold)
                            ptx::tma_load_2d(p, &maps.fb, &tail.b_full[stage], 0, t * T5_BN);
                        else
                            ptx::tma_load_2d(p, &maps.b[seg_class(sg.bytes)], &tail.b_full[stage],
                                             sg.k_off, t * T5_BN);
                    }
                    if (++stage == nstage) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: A from TMEM, B from shared memory =============
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(128, T5_BN);
            int stage = 0, as = 0;
            uint32_t ph = 0, acc_ph = 0, a_ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int sp = item % args.nsplit;
                int t_lo = sp * args.tiles_per_split;
                int t_hi = t_lo + args.tiles_per_split;
                if (t_hi > args.n_btiles) t_hi = args.n_btiles;
                ptx::mbar_wait(&tail.a_ready, a_ph);  // the epilogue stored this item's queries
                a_ph ^= 1;
                ptx::tc_fence_after_sync();
                for (int i = 0; i < t_hi - t_lo; ++i) {
                    ptx::mbar_wait(&tail.b_full[stage], ph);
                    ptx::mbar_wait(&tail.acc_empty[as], acc_ph ^ 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t b_base = ptx::smem_u32(smB + stage * b_bytes);
                    // alternate between the two (independent) query tiles
                    uint32_t acc = 0, kstep = 0;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        const uint32_t b_addr = b_base + sg.b_off / TC_BN * T5_BN;
                        const uint32_t sbo = 8 * sg.bytes, lay = seg_layout(sg.bytes);
                        for (int kk = 0; kk < sg.bytes / 32; ++kk, ++kstep) {
                            const uint64_t bd = ptx::make_smem_desc(b_addr + kk * 32, sbo, lay);
#pragma unroll
                            for (int g = 0; g < 2; ++g)
                                ptx::umma_f16_ts(
                                    tmem_base + (uint32_t)(T5_ACC0 + (as * 2 + g) * T5_BN),
                                    tmem_base + (uint32_t)(T5_A0 + g * T5_A_COLS) + kstep * 8, bd,
                                    idesc, acc);
                            acc = 1;
                        }
                    }
                    ptx::umma_commit(&tail.b_empty[stage]);
                    ptx::umma_commit(&tail.acc_full[as]);
                    if (++stage == nstage) {
                        stage = 0;
                        ph ^= 1;
                    }
                    as ^= 1;
                    if (as == 0) acc_ph ^= 1;
                }
                ptx::umma_commit(&tail.a_free);  // all MMAs that read this item's A are done
            }
        }
    } else {
        // ===================== epilogue: store A, then threshold filter =====================
        const int e = warp - 2;
        const int g = e >> 2;
        const int quad = warp & 3;
        const int row_in_cta = g * 128 + quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int L = args.sc->inexact ? args.L_inexact : args.L_exact;
        int as = 0, it_par = 0;
        uint32_t acc_ph = 0, free_ph = 0;
        unsigned ctr[4] = {0, 0, 0, 0};
        const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int qt = item / args.nsplit, sp = item % args.nsplit;
            int t_lo = sp * args.tiles_per_split;
            int t_hi = t_lo + args.tiles_per_split;
            if (t_hi > args.n_btiles) t_hi = args.n_btiles;
            const int r = qt * TC_ROWS + row_in_cta;
            const bool valid = r < args.n_rows;
            const int self = (valid && args.self_join)
                                 ? (args.row_ids ? args.row_ids[r] : args.q_begin + r)
                                 : -1;
            const size_t slot = valid ? (size_t)r * args.nsplit + sp : 0;
            key64_t *list = args.cand + slot * GCAP;
            float tau = valid ? NEG_INF : POS_INF;
            int cnt = 0;
            long long b_hi = (long long)t_hi * T5_BN;
            if (b_hi > args.nb) b_hi = args.nb;
            const int ib_hi = (int)b_hi;

            // ---- this thread's query row -> tensor memory (K-major FP16 pairs, then the fold
            //      constants); the previous item's MMAs must have finished reading the old one
            ptx::mbar_wait(&tail.a_free, free_ph ^ 1);
            free_ph ^= 1;
            ptx::tc_fence_after_sync();
            {
                const uint32_t a_addr = tmem_base + lane_addr + (uint32_t)(T5_A0 + g * T5_A_COLS);
                const uint32_t *src = reinterpret_cast<const uint32_t *>(
                    xh_q + (size_t)(args.q_op_begin + r) * d16);
                const uint32_t *fsrc = reinterpret_cast<const uint32_t *>(fold_a);
                const int data_words = d16 / 2;
                for (int c = 0; c < ksteps; ++c) {  // 8 words = one K step of 16 halves
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int word = c * 8 + j;
                        w[j] = word < data_words ? (valid ? __ldg(src + word) : 0u)
                                                 : __ldg(fsrc + (word - data_words));
                    }
                    ptx::tmem_st_32x32b_x8(a_addr + c * 8, w);
                }
                ptx::tmem_wait_st();
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tail.a_ready);

            int t0 = 0;
            for (int i = 0; i < t_hi - t_lo; ++i) {
                wait_acc_with_idle_shrinks<GCAP>(&tail.acc_full[as], acc_ph, L, lane, nullptr, tau,
                                                 cnt, list, ctr);
                ptx::tc_fence_after_sync();
                if (i == 0) {
                    ptx::mbar_wait(&tail.t0_ready, (uint32_t)it_par);
                    t0 = *reinterpret_cast<volatile int *>(&tail.item_t0[it_par]);
                }
                const int t = item_tile(i, t0, t_lo, t_hi);
                const uint32_t tbase =
                    tmem_base + lane_addr + (uint32_t)(T5_ACC0 + (as * 2 + g) * T5_BN);
                const int tile0 = t * T5_BN;
                const bool special = (t == t_hi - 1) || (self >= tile0 && self < tile0 + T5_BN);
                auto release = [&]() {
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tail.acc_empty[as]);
                };
                if (args.debug_mode != 1)
                    filter_tile_n<GCAP, T5_BN>(tbase, tile0, special, ib_hi, self, L, lane,
                                               args.debug_mode, nullptr, tau, cnt, list, ctr,
                                               release);
                else
                    release();
                as ^= 1;
                if (as == 0) acc_ph ^= 1;
            }
            it_par ^= 1;
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
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

size_t tc5_smem_bytes_n(const TcPlan &p, int nstage) {
    return 1024 + (size_t)nstage * (((size_t)p.b_bytes / TC_BN * T5_BN + 1023) & ~(size_t)1023) +
           sizeof(T5SmemTail);
}

}  // namespace

int tc5_tile_rows() { return T5_BN; }

// deepest B ring that fits; 0 if the operand row is too long for the TMEM query-tile region
int tc5_pick_nstage(const void *plan, int d16) {
    const TcPlan &p = *static_cast<const TcPlan *>(plan);
    if ((d16 + FOLD_COLS) / 2 > T5_A_COLS) return 0;
    for (int n = T5_MAXSTAGE; n >= 3; --n)
        if (tc5_smem_bytes_n(p, n) <= 227 * 1024) return n;
    return 0;
}

cudaError_t launch_tc5(const void *maps_v, const void *args_v, const void *xh_q,
                       const void *fold_a, int gcap, int d16, int n_items, int num_sms,
                       cudaStream_t stream) {
    const TcMaps &maps = *static_cast<const TcMaps *>(maps_v);
    const TcArgs &args = *static_cast<const TcArgs *>(args_v);
    const int grid = n_items < num_sms ? n_items : num_sms;
    const size_t smem = tc5_smem_bytes_n(args.plan, args.nstage);
    cudaError_t e;
#define LAUNCH_T5(G)                                                                             \
    do {                                                                                         \
        if ((e = cudaFuncSetAttribute(knn_tc5_kernel<G>,                                         \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != \
            cudaSuccess)                                                                         \
            return e;                                                                            \
        knn_tc5_kernel<G><<<grid, TC_THREADS, smem, stream>>>(                                   \
            maps, args, static_cast<const __half *>(xh_q), static_cast<const __half *>(fold_a),  \
            d16);                                                                                \
    } while (0)
    if (gcap == 256) LAUNCH_T5(256);
    else LAUNCH_T5(512);
#undef LAUNCH_T5
    return cudaGetLastError();
}

}  // namespace pqh
