// knn_tcs.cu -- the "streaming" distance-GEMM + filter kernel for dimensions whose query tile
// does not fit in shared memory (GIST-shaped d = 960, anything above ~144).
//
// Same geometry and epilogue as knn_tc_kernel (256 query rows per CTA as two M=128 accumulator
// tiles, base tiles of 128 rows, two accumulator stages, eight filter warps), but BOTH operands
// are streamed through one ring of K-blocks: a stage holds the [256 x 64] slice of the query
// tile and the [128 x 64] slice of the base tile for one K-block (48 KB), the MMA thread
// accumulates over all K-blocks of a tile before committing the accumulator.  With d this large
// the kernel is bound by the tensor cores and by L2->SM operand traffic, not by the epilogue:
// each accumulator element is the result of d/128 times more MMA work than at SIFT's d = 128.
#include <cstdlib>

#include "kernels.h"
#include "knn_tc_common.cuh"

namespace pqh {

namespace {

constexpr int TS_NSTAGE = 4;
constexpr int TS_A_BYTES = TC_ROWS * 128;  // one K-block of the query tile (SW128: 64 halves)
constexpr int TS_B_BYTES = TC_BN * 128;
constexpr int TS_STAGE_BYTES = TS_A_BYTES + TS_B_BYTES;

struct TsSmemTail {
    uint64_t full[TS_NSTAGE], empty[TS_NSTAGE];
    EpiBars eb;
};

// K-blocks of a padded operand row of d16 halves followed by the 16 fold columns:
// full 64-half blocks (SW128), then an optional 32-half (SW64) and 16-half (SW32) block, then
// the fold block (SW32, separate tensors).
struct KBlocks {
    int nfull, has32, has16, total;
    __device__ __host__ explicit KBlocks(int d16) {
        nfull = d16 / 64;
        const int rem = d16 - nfull * 64;
        has32 = rem >= 32;
        has16 = (rem - (has32 ? 32 : 0)) >= 16;
        total = nfull + has32 + has16 + 1;
    }
    // -> bytes per row of block i, element offset, fold flag
    __device__ __forceinline__ void get(int i, int &bytes, int &k_off, int &fold) const {
        fold = 0;
        if (i < nfull) {
            bytes = 128;
            k_off = i * 64;
            return;
        }
        int j = i - nfull;
        if (has32) {
            if (j == 0) {
                bytes = 64;
                k_off = nfull * 64;
                return;
            }
            --j;
        }
        if (has16) {
            if (j == 0) {
                bytes = 32;
                k_off = nfull * 64 + (has32 ? 32 : 0);
                return;
            }
            --j;
        }
        bytes = 32;
        k_off = 0;
        fold = 1;
    }
};

template <int GCAP>
__global__ void __launch_bounds__(tc_threads(1), 1)
knn_tcs_kernel(const __grid_constant__ TcMaps maps, const TcArgs args, int d16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    TsSmemTail &tail = *reinterpret_cast<TsSmemTail *>(smem + TS_NSTAGE * TS_STAGE_BYTES);
    const KBlocks kb(d16);

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
        for (int i = 0; i < TS_NSTAGE; ++i) {
            ptx::mbar_init(&tail.full[i], 1);
            ptx::mbar_init(&tail.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tail.eb.acc_full[i], 1);
            ptx::mbar_init(&tail.eb.acc_empty[i], TC_EPI_WARPS);
        }
        ptx::mbar_init(&tail.eb.t0_ready, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(&tail.eb.tmem_base);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = tail.eb.tmem_base;

    const int n_items = total_items(args);

    if (warp == 0) {
        // ===================== TMA producer: one K-block of A and B per stage =============
        if (ptx::elect_one()) {
            int stage = 0, it_par = 0, my_next = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x;; item += gridDim.x) {
                if (item >= n_items) {
                    // consumers saw the flip of the last item long ago (their last accumulator
                    // wait is behind the last loads, which are behind that flip)
                    tail.eb.item_id[it_par] = -1;
                    ptx::mbar_arrive(&tail.eb.t0_ready);
                    break;
                }
                const ItemInfo it = decode_item(args, item);
                const int qt = it.qt, t_lo = it.t_lo, t_hi = it.t_hi;
                const int arow = args.q_op_begin + qt * TC_ROWS;
                // the slot of item i is reused by item i+2; the producer is at most one ring
                // (a fraction of a tile) ahead of the MMA, the epilogue at most two tiles behind
                const ItemOrder ord = item_order(args, qt, t_lo, t_hi, my_next);
                const int wlen = own_window_len(args, qt, t_lo, t_hi);
                tail.eb.item_w0[it_par] = ord.w0;
                tail.eb.item_uc[it_par] = ord.uc;
                tail.eb.item_id[it_par] = item;
                ptx::mbar_arrive(&tail.eb.t0_ready);
                it_par ^= 1;
                for (int ti = 0; ti < t_hi - t_lo; ++ti) {
                    const int t = item_tile(ti, ord, wlen, t_lo, t_hi);
                    if (ti == wlen) my_next = t;
                    if (ti >= wlen) publish_cursor(args, t);
                    for (int i = 0; i < kb.total; ++i) {
                        int bytes, k_off, fold;
                        kb.get(i, bytes, k_off, fold);
                        ptx::mbar_wait(&tail.empty[stage], ph ^ 1);
                        ptx::mbar_arrive_expect_tx(&tail.full[stage], (TC_ROWS + TC_BN) * bytes);
                        uint8_t *sa = smem + stage * TS_STAGE_BYTES;
                        uint8_t *sb = sa + TS_A_BYTES;
                        if (fold) {
                            ptx::tma_load_2d(sa, &maps.fa, &tail.full[stage], 0, 0);
                            ptx::tma_load_2d(sb, &maps.fb, &tail.full[stage], 0, t * TC_BN);
                        } else {
                            const int cls = seg_class(bytes);
                            ptx::tma_load_2d(sa, &maps.a[cls], &tail.full[stage], k_off, arow);
                            ptx::tma_load_2d(sb, &maps.b[cls], &tail.full[stage], k_off, t * TC_BN);
                        }
                        if (++stage == TS_NSTAGE) {
                            stage = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(128, TC_BN);
            int stage = 0, as = 0;
            uint32_t ph = 0, acc_ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const ItemInfo it = decode_item(args, item);
                const int t_lo = it.t_lo, t_hi = it.t_hi;
                for (int ti = 0; ti < t_hi - t_lo; ++ti) {
                    ptx::mbar_wait(&tail.eb.acc_empty[as], acc_ph ^ 1);
                    ptx::tc_fence_after_sync();
                    for (int i = 0; i < kb.total; ++i) {
                        int bytes, k_off, fold;
                        kb.get(i, bytes, k_off, fold);
                        ptx::mbar_wait(&tail.full[stage], ph);
                        ptx::tc_fence_after_sync();
                        const uint32_t a_base = ptx::smem_u32(smem + stage * TS_STAGE_BYTES);
                        const uint32_t b_base = a_base + TS_A_BYTES;
                        const uint32_t sbo = 8 * bytes, lay = seg_layout(bytes);
                        // alternate between the two (independent) query tiles so that
                        // consecutive MMAs never accumulate into the same tile
                        for (int kk = 0; kk < bytes / 32; ++kk) {
                            const uint64_t bd = ptx::make_smem_desc(b_base + kk * 32, sbo, lay);
#pragma unroll
                            for (int g = 0; g < 2; ++g)
                                ptx::umma_f16(tmem_base + (uint32_t)(as * 256 + g * 128),
                                              ptx::make_smem_desc(a_base + g * 128 * bytes + kk * 32,
                                                                  sbo, lay),
                                              bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                        }
                        ptx::umma_commit(&tail.empty[stage]);
                        if (++stage == TS_NSTAGE) {
                            stage = 0;
                            ph ^= 1;
                        }
                    }
                    ptx::umma_commit(&tail.eb.acc_full[as]);
                    as ^= 1;
                    if (as == 0) acc_ph ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: threshold filter (as knn_tc_kernel) =================
        epilogue_warps<GCAP, false, false>(args, warp - 2, warp & 3, lane, &tail.eb);
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

template <int GCAP>
cudaError_t launch_tcs_one(const TcMaps &maps, const TcArgs &args, int d16, int grid,
                           cudaStream_t stream) {
    const size_t smem = 1024 + (size_t)TS_NSTAGE * TS_STAGE_BYTES + sizeof(TsSmemTail);
    cudaError_t e = cudaFuncSetAttribute(knn_tcs_kernel<GCAP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    knn_tcs_kernel<GCAP><<<grid, tc_threads(1), smem, stream>>>(maps, args, d16);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_tcs(const void *maps_v, const void *args_v, int gcap, int d16, int n_items,
                       int num_sms, cudaStream_t stream) {
    const TcMaps &maps = *static_cast<const TcMaps *>(maps_v);
    const TcArgs &args = *static_cast<const TcArgs *>(args_v);
    const int grid = n_items < num_sms ? n_items : num_sms;
    return gcap == 256 ? launch_tcs_one<256>(maps, args, d16, grid, stream)
                       : launch_tcs_one<512>(maps, args, d16, grid, stream);
}

}  // namespace pqh
