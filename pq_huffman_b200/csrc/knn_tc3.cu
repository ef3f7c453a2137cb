// knn_tc3.cu -- the CTA-pair distance-GEMM + filter kernel (the default tensor-core path).
//
// Two CTAs of one cluster (one TPC) issue tcgen05.mma.cta_group::2.  The pair owns 512 query
// rows per work item -- each CTA 256 rows = two M halves, A operand resident in its shared
// memory for the whole sweep -- and sweeps the base in tiles of 128 rows, each CTA TMA-loading
// only ITS 64-row half of a tile (the tensor cores of both SMs consume both halves).  Per tile
// the leader issues two M=256,N=128 accumulations (one per M half); 2 accumulator stages x 2
// halves x 128 columns = all 512 TMEM columns, so the MMAs of tile t+1 overlap the epilogue of
// tile t.  Against a single-CTA M=128,N=128 tiling this halves both the L2->SM operand traffic
// and the shared-memory operand reads per FLOP.
//
// The epilogue is split in two so that rare, slow work never sits between the tensor cores and
// their next accumulator stage:
//   * 8 FILTER warps (one thread per query row, the TMEM lane it may read) compare the
//     accumulator against the row's threshold and push the few (row, score, id) hits into a
//     shared-memory ring; their time per tile is short and nearly constant;
//   * 2 COLLECTOR warps (one per M half) drain the rings, append to the rows' candidate lists
//     in global memory, shrink full lists to their L best (shrink_row) and publish the raised
//     threshold in shared memory, asynchronously.
// A filter thread may therefore see a slightly stale (lower) threshold, which only admits
// extra candidates; every rejected element still has a score <= the row's final threshold,
// which is what the certificate in knn_rerank_kernel needs.
//
// Barriers: "full" (TMA -> MMA) live in the leader CTA and count the bytes of both CTAs'
// loads; "empty" and "accumulator full" are per CTA, signalled by multicast tcgen05.commit;
// "accumulator empty" lives in the leader and is arrived on remotely by all 16 filter warps.
#include <cstdlib>

#include "kernels.h"
#include "knn_tc_common.cuh"

namespace pqh {

namespace {

constexpr int T3_MAXSTAGE = 8;  // B-tile ring depth is chosen by the host to fill shared memory
constexpr int T3_THREADS = 384;  // warp 0 TMA, 1 MMA, 2..9 filter, 10..11 collector
constexpr int T3_RING = 1024;    // ring entries per M half (power of two)
constexpr int T3_RING_LOG = 10;

struct RingEntry {
    uint32_t score;  // float bits of S'
    uint32_t id;     // base position
    uint32_t row;    // local row 0..127 inside the M half
    uint32_t seq;    // (position >> T3_RING_LOG) + 1 once the entry is complete
};

struct T3Smem {
    RingEntry ring[2][T3_RING];
    float tau[256];
    int cnt[256];
    uint32_t ring_tail[2], ring_head[2], done_cnt[2], pad_[2];
    uint64_t a_full, a_empty;
    uint64_t b_full[T3_MAXSTAGE], b_empty[T3_MAXSTAGE];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    return *reinterpret_cast<const volatile uint32_t *>(p);
}

// Warp-uniform call (taken when some lane has a hit): reserve ring slots with one atomic per
// warp and publish the entries.
__device__ __noinline__ void ring_push(T3Smem &sm, int g, bool hit, float v, int id, int row,
                                       int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&sm.ring_tail[g], (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (hit) {
        const uint32_t pos = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
        uint32_t spins = 0;
        while (pos - ld_volatile_u32(&sm.ring_head[g]) >= (uint32_t)T3_RING) {
            __nanosleep(64);
            if (++spins > (1u << 26)) {
                printf("libpqh_knn: candidate ring stuck (block %d)\n", (int)blockIdx.x);
                __trap();
            }
        }
        RingEntry *e = &sm.ring[g][pos & (T3_RING - 1)];
        e->score = __float_as_uint(v);
        e->id = (uint32_t)id;
        e->row = (uint32_t)row;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(&e->seq) = (pos >> T3_RING_LOG) + 1u;
    }
}

template <int GCAP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
knn_tc3_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcPlan &plan = args.plan;  // a_off: 256-row A region; b_off is for 128 rows: use half
    const int b_bytes = plan.b_bytes / 2;
    uint8_t *smA = smem;
    uint8_t *smB = smA + plan.a_bytes;
    const int nstage = args.nstage;
    T3Smem &sm = *reinterpret_cast<T3Smem *>(smB + nstage * b_bytes);

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
        ptx::mbar_init(&sm.a_full, 1);
        ptx::mbar_init(&sm.a_empty, 1);
        for (int i = 0; i < T3_MAXSTAGE; ++i) {
            ptx::mbar_init(&sm.b_full[i], 1);
            ptx::mbar_init(&sm.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&sm.acc_full[i], 1);
            ptx::mbar_init(&sm.acc_empty[i], 16);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc_2sm<512>(&sm.tmem_base);
    if (warp == 3) {
        for (int i = lane; i < 2 * T3_RING; i += 32) (&sm.ring[0][0])[i].seq = 0;
        if (lane < 2) {
            sm.ring_tail[lane] = 0;
            sm.ring_head[lane] = 0;
            sm.done_cnt[lane] = 0;
        }
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();  // barriers of both CTAs initialised before any remote signal
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = sm.tmem_base;

    const int n_items = args.n_qtiles * args.nsplit;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs, own halves) =====================
        if (ptx::elect_one()) {
            const uint32_t a_full_l = ptx::mapa_u32(ptx::smem_u32(&sm.a_full), 0);
            int stage = 0;
            uint32_t ph = 0, a_ph = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int qt = item / args.nsplit, sp = item % args.nsplit;
                int t_lo = sp * args.tiles_per_split;
                int t_hi = t_lo + args.tiles_per_split;
                if (t_hi > args.n_btiles) t_hi = args.n_btiles;
                ptx::mbar_wait(&sm.a_empty, a_ph ^ 1);
                if (leader) ptx::mbar_arrive_expect_tx(&sm.a_full, 2 * plan.a_bytes);
                const int arow0 = args.q_op_begin + qt * 512 + (int)rank * 256;
                for (int s = 0; s < plan.nseg; ++s) {
                    const TcSeg &sg = plan.seg[s];
                    if (sg.fold)
                        ptx::tma_load_2d_2sm(smA + sg.a_off, &maps.fa, a_full_l, 0, 0);
                    else
                        ptx::tma_load_2d_2sm(smA + sg.a_off, &maps.a[seg_class(sg.bytes)], a_full_l,
                                             sg.k_off, arow0);
                }
                a_ph ^= 1;
                for (int t = t_lo; t < t_hi; ++t) {
                    ptx::mbar_wait(&sm.b_empty[stage], ph ^ 1);
                    if (leader) ptx::mbar_arrive_expect_tx(&sm.b_full[stage], 2 * b_bytes);
                    const uint32_t b_full_l = ptx::mapa_u32(ptx::smem_u32(&sm.b_full[stage]), 0);
                    uint8_t *dst = smB + stage * b_bytes;
                    const int row0 = t * TC_BN + (int)rank * 64;
                    for (int s = 0; s < plan.nseg; ++s) {
                        const TcSeg &sg = plan.seg[s];
                        if (sg.fold)
                            ptx::tma_load_2d_2sm(dst + sg.b_off / 2, &maps.fb, b_full_l, 0, row0);
                        else
                            ptx::tma_load_2d_2sm(dst + sg.b_off / 2, &maps.b[seg_class(sg.bytes)],
                                                 b_full_l, sg.k_off, row0);
                    }
                    if (++stage == nstage) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(256, TC_BN);
            int stage = 0, as = 0;
            uint32_t ph = 0, acc_ph = 0, a_ph = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int sp = item % args.nsplit;
                int t_lo = sp * args.tiles_per_split;
                int t_hi = t_lo + args.tiles_per_split;
                if (t_hi > args.n_btiles) t_hi = args.n_btiles;
                ptx::mbar_wait(&sm.a_full, a_ph);
                a_ph ^= 1;
                for (int t = t_lo; t < t_hi; ++t) {
                    ptx::mbar_wait(&sm.b_full[stage], ph);
                    ptx::mbar_wait(&sm.acc_empty[as], acc_ph ^ 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t b_base = ptx::smem_u32(smB + stage * b_bytes);
                    const uint32_t a_base = ptx::smem_u32(smA);
#pragma unroll 1
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256 + g * 128);
                        uint32_t acc = 0;
                        for (int s = 0; s < plan.nseg; ++s) {
                            const TcSeg &sg = plan.seg[s];
                            const uint32_t a_addr = a_base + sg.a_off + g * 128 * sg.bytes;
                            const uint32_t b_addr = b_base + sg.b_off / 2;
                            const uint32_t sbo = 8 * sg.bytes, lay = seg_layout(sg.bytes);
                            for (int kk = 0; kk < sg.bytes / 32; ++kk) {
                                ptx::umma_f16_2sm(d_tmem,
                                                  ptx::make_smem_desc(a_addr + kk * 32, sbo, lay),
                                                  ptx::make_smem_desc(b_addr + kk * 32, sbo, lay),
                                                  idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                    ptx::umma_commit_2sm(&sm.b_empty[stage]);
                    ptx::umma_commit_2sm(&sm.acc_full[as]);
                    if (++stage == nstage) {
                        stage = 0;
                        ph ^= 1;
                    }
                    as ^= 1;
                    if (as == 0) acc_ph ^= 1;
                }
                ptx::umma_commit_2sm(&sm.a_empty);
            }
        }
    } else if (warp < 10) {
        // ===================== filter warps =====================
        const int e = warp - 2;
        const int g = e >> 2;       // M half
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        const int row_g = quad * 32 + lane;  // row inside the M half (= TMEM lane)
        const int row_l = g * 128 + row_g;   // row inside this CTA
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        uint32_t acc_empty_l[2];
        acc_empty_l[0] = ptx::mapa_u32(ptx::smem_u32(&sm.acc_empty[0]), 0);
        acc_empty_l[1] = ptx::mapa_u32(ptx::smem_u32(&sm.acc_empty[1]), 0);
        int as = 0;
        uint32_t acc_ph = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int qt = item / args.nsplit, sp = item % args.nsplit;
            int t_lo = sp * args.tiles_per_split;
            int t_hi = t_lo + args.tiles_per_split;
            if (t_hi > args.n_btiles) t_hi = args.n_btiles;
            const int r = qt * 512 + (int)rank * 256 + row_l;
            const bool valid = r < args.n_rows;
            const int self = (valid && args.self_join)
                                 ? (args.row_ids ? args.row_ids[r] : args.q_begin + r)
                                 : -1;
            long long b_hi = (long long)t_hi * TC_BN;
            if (b_hi > args.nb) b_hi = args.nb;
            const int ib_hi = (int)b_hi;
            // the collector of this M half has reset the rows' state for this item
            asm volatile("bar.sync %0, 160;" ::"r"(1 + g) : "memory");

            for (int t = t_lo; t < t_hi; ++t) {
                ptx::mbar_wait(&sm.acc_full[as], acc_ph);
                ptx::tc_fence_after_sync();
                if (args.debug_mode != 1) {
                    const uint32_t tbase = tmem_base + lane_addr + (uint32_t)(as * 256 + g * 128);
                    const int tile0 = t * TC_BN;
                    const bool special = (t == t_hi - 1) || (self >= tile0 && self < tile0 + TC_BN);
                    const float tau = *reinterpret_cast<volatile float *>(&sm.tau[row_l]);
                    uint32_t ra[32], rb[32];
                    ptx::tmem_ld_32x32b_x32(tbase, ra);
                    ptx::tmem_wait_ld();
#pragma unroll 1
                    for (int c2 = 0; c2 < TC_BN / 32; c2 += 2) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = c2 + h;
                            uint32_t(&cur)[32] = h ? rb : ra;
                            uint32_t(&nxt)[32] = h ? ra : rb;
                            // chunk c+1 is in flight while chunk c is filtered
                            if (c + 1 < TC_BN / 32)
                                ptx::tmem_ld_32x32b_x32(tbase + (c + 1) * 32, nxt);
                            float m = ptx::fmax3(__uint_as_float(cur[0]), __uint_as_float(cur[1]),
                                                 __uint_as_float(cur[2]));
#pragma unroll
                            for (int j = 3; j < 31; j += 2)
                                m = ptx::fmax3(m, __uint_as_float(cur[j]),
                                               __uint_as_float(cur[j + 1]));
                            m = fmaxf(m, __uint_as_float(cur[31]));
                            if (__any_sync(0xffffffffu, m > tau) && args.debug_mode != 2) {
                                const int col0 = tile0 + c * 32;
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float v = __uint_as_float(cur[j]);
                                    const int id = col0 + j;
                                    const bool hit =
                                        v > tau && (!special || (id < ib_hi && id != self));
                                    if (__any_sync(0xffffffffu, hit))
                                        ring_push(sm, g, hit, v, id, row_g, lane);
                                }
                            }
                            if (c + 1 < TC_BN / 32) ptx::tmem_wait_ld();
                        }
                    }
                }
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(acc_empty_l[as]);
                as ^= 1;
                if (as == 0) acc_ph ^= 1;
            }
            // all pushes of this warp for this item are complete
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&sm.done_cnt[g], 1u);
            }
        }
    } else {
        // ===================== collector warps =====================
        const int g = warp - 10;
        const int L = args.sc->inexact ? args.L_inexact : args.L_exact;
        const int trigger = (2 * L < GCAP - 32) ? 2 * L : GCAP - 32;
        const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
        uint32_t head = 0, items_done = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int qt = item / args.nsplit, sp = item % args.nsplit;
            const int r0 = qt * 512 + (int)rank * 256 + g * 128;  // first row of this M half
            for (int i = lane; i < 128; i += 32) {
                sm.tau[g * 128 + i] = (r0 + i < args.n_rows) ? NEG_INF : POS_INF;
                sm.cnt[g * 128 + i] = 0;
            }
            __syncwarp();
            asm volatile("bar.sync %0, 160;" ::"r"(1 + g) : "memory");
            ++items_done;
            for (;;) {
                const uint32_t idx = head + (uint32_t)lane;
                const RingEntry *e = &sm.ring[g][idx & (T3_RING - 1)];
                const bool ready =
                    ld_volatile_u32(&e->seq) == (idx >> T3_RING_LOG) + 1u;
                const unsigned rb = __ballot_sync(0xffffffffu, ready);
                const int nready = (rb == 0xffffffffu) ? 32 : (__ffs(~rb) - 1);
                if (nready == 0) {
                    // finished when every filter warp of this half is done and the ring is empty
                    if (ld_volatile_u32(&sm.done_cnt[g]) == 4u * items_done &&
                        ld_volatile_u32(&sm.ring_tail[g]) == head)
                        break;
                    __nanosleep(100);
                    continue;
                }
                __threadfence_block();
                int row = -1, pos = 0;
                if (lane < nready) {
                    const uint32_t sbits = e->score, id = e->id;
                    row = (int)e->row;
                    pos = atomicAdd(&sm.cnt[g * 128 + row], 1);
                    const size_t slot = (size_t)(r0 + row) * args.nsplit + sp;
                    __stcg(args.cand + slot * GCAP + pos, make_key(-__uint_as_float(sbits), id));
                }
                head += (uint32_t)nready;
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&sm.ring_head[g]) = head;
                // shrink every row that crossed its trigger
                unsigned need = __ballot_sync(0xffffffffu, row >= 0 && pos + 1 > trigger);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int srow = __shfl_sync(0xffffffffu, row, src);
                    const int n = sm.cnt[g * 128 + srow];
                    if (n > trigger) {  // not already shrunk for an earlier lane of this batch
                        const size_t slot = (size_t)(r0 + srow) * args.nsplit + sp;
                        key64_t *l = args.cand + slot * GCAP;
                        int kept;
                        const uint32_t x = (GCAP > 128 && n <= 128)
                                               ? shrink_row<128>(l, n, L, lane, kept)
                                               : shrink_row<GCAP>(l, n, L, lane, kept);
                        if (lane == 0) {
                            sm.cnt[g * 128 + srow] = kept;
                            const float t_new = -ordered_to_float(x);
                            const float t_old = sm.tau[g * 128 + srow];
                            *reinterpret_cast<volatile float *>(&sm.tau[g * 128 + srow]) =
                                fmaxf(t_old, t_new);
                        }
                        __syncwarp();
                    }
                }
            }
            // item finished: publish list lengths and final thresholds
            for (int i = lane; i < 128; i += 32) {
                if (r0 + i < args.n_rows) {
                    const size_t slot = (size_t)(r0 + i) * args.nsplit + sp;
                    args.cand_cnt[slot] = sm.cnt[g * 128 + i];
                    args.cand_tau[slot] = sm.tau[g * 128 + i];
                }
            }
            __syncwarp();
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_all();  // the peer may still be signalling this CTA's barriers / TMEM
    if (warp == 2) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

size_t tc3_smem_bytes_n(const TcPlan &p, int nstage) {
    return 1024 + (size_t)p.a_bytes + (size_t)nstage * (p.b_bytes / 2) + sizeof(T3Smem);
}

}  // namespace

// deepest B ring (2..8 stages) that fits in 227 KB of shared memory; 0 if even 2 do not fit
int tc3_pick_nstage(const void *plan) {
    const TcPlan &p = *static_cast<const TcPlan *>(plan);
    for (int n = T3_MAXSTAGE; n >= 2; --n)
        if (tc3_smem_bytes_n(p, n) <= 227 * 1024) return n;
    return 0;
}

cudaError_t launch_tc3(const void *maps_v, const void *args_v, int gcap, int n_items, int num_sms,
                       cudaStream_t stream) {
    const TcMaps &maps = *static_cast<const TcMaps *>(maps_v);
    const TcArgs &args = *static_cast<const TcArgs *>(args_v);
    const int clusters = n_items < num_sms / 2 ? n_items : num_sms / 2;
    const int grid = 2 * clusters;
    const size_t smem = tc3_smem_bytes_n(args.plan, args.nstage);
    cudaError_t e;
#define LAUNCH_T3(G)                                                                             \
    do {                                                                                         \
        if ((e = cudaFuncSetAttribute(knn_tc3_kernel<G>,                                         \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != \
            cudaSuccess)                                                                         \
            return e;                                                                            \
        knn_tc3_kernel<G><<<grid, T3_THREADS, smem, stream>>>(maps, args);                       \
    } while (0)
    if (gcap == 128) LAUNCH_T3(128);
    else if (gcap == 256) LAUNCH_T3(256);
    else LAUNCH_T3(512);
#undef LAUNCH_T3
    return cudaGetLastError();
}

}  // namespace pqh
