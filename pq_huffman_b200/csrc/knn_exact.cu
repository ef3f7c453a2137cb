// knn_exact.cu -- exact FP32 brute-force kNN (SIMT, no tensor cores), sm_100a.
//
// Role: (1) the certificate fallback of the tensor-core path (rows whose
// candidate list cannot be proven complete are redone here), (2) the
// PQH_ALGO_EXACT debug path, (3) small problems.  Distances are computed with
// exactly the scalar sequence of the reference's get_real_dist
// (compute_nn_fast.c:304-311): tmp = a[i]-b[i]; acc = acc + tmp*tmp for
// i = 0..d-1 with separately rounded subtract, multiply and add
// (__fsub_rn/__fmul_rn/__fadd_rn are never contracted into FMA), so every
// distance is bit-identical to the oracle's and the selection below, done on
// (distance, position) keys, reproduces the oracle's rows exactly.
//
// Work split: grid = (query tiles of 64 rows, base splits).  A CTA walks its
// base range in tiles of 64 rows; 256 threads each own a 4x4 block of pairs.
// Passing candidates (dist <= row threshold) are appended to a per-row buffer
// in shared memory; a warp merges a row's buffer into the row's sorted list
// (global memory, L2 resident, touched O(log n) times per row) when the buffer
// could overflow.  knn_finalize_kernel merges the splits and writes the rows.
#include "common.cuh"
#include "kernels.h"

namespace pqh {

namespace {

constexpr int TQ = 64;    // query rows per CTA
constexpr int TB = 64;    // base rows per tile
constexpr int KC = 16;    // dimensions per shared-memory chunk
constexpr int CAP = 128;  // append-buffer slots per row (>= 2*TB)
constexpr int LDS = TQ + 4;

struct ExactSmem {
    float vs[KC][LDS];
    float bs[KC][LDS];
    key64_t buf[TQ][CAP];
    float tau[TQ];
    int cnt[TQ];
    int len[TQ];
    int cur[TQ];
    int qid[TQ];
};

__global__ void __launch_bounds__(256)
knn_exact_kernel(const float *__restrict__ base, const float *__restrict__ queries,
                 const int *__restrict__ row_ids, int q_begin, int q_row0, int n_rows, int nb,
                 int d, int kcap,
                 int self_join, int nsplit, int split_rows, key64_t *__restrict__ lists,
                 int *__restrict__ meta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExactSmem &sm = *reinterpret_cast<ExactSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int ty = tid >> 4, tx = tid & 15;

    if (tid < TQ) {
        int slot = tile * TQ + tid;
        int q = -1;
        if (slot < n_rows) q = row_ids ? row_ids[slot] : q_begin + slot;
        sm.qid[tid] = q;
        sm.tau[tid] = __int_as_float(0x7f800000);  // +inf
        sm.cnt[tid] = 0;
        sm.len[tid] = 0;
        sm.cur[tid] = 0;
    }
    __syncthreads();

    const long long b_lo = (long long)split * split_rows;
    long long b_hi = b_lo + split_rows;
    if (b_hi > nb) b_hi = nb;

    // loader mapping: thread -> (row = tid/4, 4 consecutive dims (tid%4)*4)
    const int lrow = tid >> 2, lcol = (tid & 3) * 4;
    const int lq = sm.qid[lrow];

    for (long long b0 = b_lo; b0 < b_hi; b0 += TB) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

        const long long lb = b0 + lrow;
        for (int k0 = 0; k0 < d; k0 += KC) {
            float qv[4], bv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int col = k0 + lcol + c;
                qv[c] = (lq >= 0 && col < d) ? queries[(size_t)(lq - q_row0) * d + col] : 0.0f;
                bv[c] = (lb < b_hi && col < d) ? base[(size_t)lb * d + col] : 0.0f;
            }
            __syncthreads();  // previous chunk fully consumed
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                sm.vs[lcol + c][lrow] = qv[c];
                sm.bs[lcol + c][lrow] = bv[c];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                float4 a4 = *reinterpret_cast<const float4 *>(&sm.vs[i][ty * 4]);
                float4 b4 = *reinterpret_cast<const float4 *>(&sm.bs[i][tx * 4]);
                float a[4] = {a4.x, a4.y, a4.z, a4.w};
                float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float t = __fsub_rn(a[r], b[c]);
                        acc[r][c] = __fadd_rn(acc[r][c], __fmul_rn(t, t));
                    }
            }
        }

        // candidate filter + append
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = ty * 4 + r;
            const int q = sm.qid[row];
            const float tau = sm.tau[row];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                long long id = b0 + tx * 4 + c;
                if (q >= 0 && id < b_hi && !(self_join && id == q) && acc[r][c] <= tau) {
                    int p = atomicAdd(&sm.cnt[row], 1);
                    sm.buf[row][p] = make_key(acc[r][c], (uint32_t)id);
                }
            }
        }
        __syncthreads();

        const bool last = (b0 + TB >= b_hi);
        for (int j = 0; j < TQ / 8; ++j) {
            const int row = warp * (TQ / 8) + j;
            const int n = sm.cnt[row];
            if (n > 0 && (last || n > CAP - TB)) {
                warp_sort_keys<CAP>(sm.buf[row], n, lane);
                const size_t slot = (size_t)(tile * TQ + row) * nsplit + split;
                key64_t *l0 = lists + slot * 2 * (size_t)kcap;
                const int cur = sm.cur[row];
                const key64_t *src = l0 + (size_t)cur * kcap;
                key64_t *dst = l0 + (size_t)(cur ^ 1) * kcap;
                int m = warp_merge_keys(src, sm.len[row], sm.buf[row], n, dst, kcap, lane);
                __syncwarp();
                if (lane == 0) {
                    sm.cur[row] = cur ^ 1;
                    sm.len[row] = m;
                    sm.cnt[row] = 0;
                    if (m == kcap) sm.tau[row] = key_score(dst[kcap - 1]);
                }
                __syncwarp();
            }
        }
        // the __syncthreads at the top of the next tile's first chunk orders
        // these updates before the next filter pass
    }
    __syncthreads();
    if (tid < TQ) {
        const size_t slot = (size_t)(tile * TQ + tid) * nsplit + split;
        meta[slot] = sm.len[tid] | (sm.cur[tid] << 30);
    }
}

// One warp per query row: merge the row's nsplit lists, write the output row.
__global__ void __launch_bounds__(256)
knn_finalize_kernel(const int *__restrict__ row_ids, int q_begin, int out_q_begin, int n_rows,
                    int k, int kcap, int self_join, int nsplit, const key64_t *__restrict__ lists,
                    const int *__restrict__ meta, key64_t *__restrict__ scratch,
                    int *__restrict__ assign, float *__restrict__ dis) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (slot >= n_rows) return;
    const int q = row_ids ? row_ids[slot] : q_begin + slot;

    const key64_t *acc = nullptr;
    int na = 0;
    key64_t *sc = scratch ? scratch + (size_t)slot * 2 * kcap : nullptr;
    int flip = 0;
    for (int s = 0; s < nsplit; ++s) {
        const size_t ls = (size_t)slot * nsplit + s;
        const int m = meta[ls];
        const int len = m & 0x3fffffff, cur = (m >> 30) & 1;
        const key64_t *l = lists + (ls * 2 + cur) * (size_t)kcap;
        if (s == 0) {
            acc = l;
            na = len;
        } else if (len > 0) {
            key64_t *out = sc + (size_t)flip * kcap;
            na = warp_merge_keys(acc, na, l, len, out, kcap, lane);
            __syncwarp();
            acc = out;
            flip ^= 1;
        }
    }
    const size_t orow = (size_t)(q - out_q_begin) * k;
    int first = 0;
    if (self_join) {
        if (lane == 0) {
            assign[orow] = q;
            dis[orow] = 0.0f;
        }
        first = 1;
    }
    for (int i = lane; i < k - first; i += 32) {
        if (i < na) {
            key64_t x = acc[i];
            assign[orow + first + i] = (int)key_pos(x);
            dis[orow + first + i] = key_score(x);
        } else {
            assign[orow + first + i] = -1;
            dis[orow + first + i] = __int_as_float(0x7f800000);
        }
    }
}

}  // namespace

size_t exact_smem_bytes() { return sizeof(ExactSmem); }

int exact_pick_nsplit(int n_rows, int nb, int num_sms) {
    int tiles = (n_rows + TQ - 1) / TQ;
    if (tiles <= 0) return 1;
    int want = (2 * num_sms + tiles - 1) / tiles;
    int max_split = nb / (TB * 8);
    if (max_split < 1) max_split = 1;
    if (max_split > 64) max_split = 64;
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    return want;
}

size_t exact_workspace_bytes(int n_rows, int kcap, int nsplit) {
    size_t slots = (size_t)((n_rows + TQ - 1) / TQ) * TQ;
    size_t lists = slots * nsplit * 2 * (size_t)kcap * sizeof(key64_t);
    size_t meta = slots * nsplit * sizeof(int);
    size_t scratch = nsplit > 1 ? slots * 2 * (size_t)kcap * sizeof(key64_t) : 0;
    return lists + ((meta + 255) & ~(size_t)255) + scratch + 512;
}

// Upper bound of exact_workspace_bytes over every row count <= n_rows_max (the split
// count grows as the row count shrinks: tiles * nsplit <= 2 * num_sms + tiles).
size_t exact_workspace_bound(int n_rows_max, int kcap, int num_sms) {
    size_t tiles = (size_t)(n_rows_max + TQ - 1) / TQ;
    size_t slot_splits = (size_t)TQ * (2 * (size_t)num_sms + tiles + 1);
    size_t slots = tiles * TQ;
    return slot_splits * (2 * (size_t)kcap * sizeof(key64_t) + sizeof(int)) +
           slots * 2 * (size_t)kcap * sizeof(key64_t) + 4096;
}

// Launches the exact search for n_rows query rows (row_ids == nullptr: rows
// q_begin .. q_begin+n_rows; otherwise the absolute query ids listed).  Query q's
// vector is row (q - q_row0) of d_queries; its output row is (q - out_q_begin).
cudaError_t launch_knn_exact(const float *d_base, const float *d_queries, const int *d_row_ids,
                             int q_begin, int q_row0, int out_q_begin, int n_rows, int nb, int d,
                             int k,
                             int self_join, int *d_assign, float *d_dis, void *d_work,
                             size_t work_bytes, int num_sms, cudaStream_t stream, int *launches) {
    if (n_rows <= 0) return cudaSuccess;
    const int kcap = self_join ? k - 1 : k;
    const int tiles = (n_rows + TQ - 1) / TQ;
    int nsplit = 1;
    key64_t *lists = nullptr, *scratch = nullptr;
    int *meta = nullptr;
    if (kcap > 0) {
        nsplit = exact_pick_nsplit(n_rows, nb, num_sms);
        if (exact_workspace_bytes(n_rows, kcap, nsplit) > work_bytes) return cudaErrorMemoryAllocation;
        size_t slots = (size_t)tiles * TQ;
        unsigned char *p = static_cast<unsigned char *>(d_work);
        p = reinterpret_cast<unsigned char *>(((uintptr_t)p + 255) & ~(uintptr_t)255);
        lists = reinterpret_cast<key64_t *>(p);
        p += slots * nsplit * 2 * (size_t)kcap * sizeof(key64_t);
        meta = reinterpret_cast<int *>(p);
        p += (slots * nsplit * sizeof(int) + 255) & ~(size_t)255;
        scratch = nsplit > 1 ? reinterpret_cast<key64_t *>(p) : nullptr;

        int split_rows = (nb + nsplit - 1) / nsplit;
        split_rows = (split_rows + TB - 1) / TB * TB;
        cudaError_t e = cudaFuncSetAttribute(knn_exact_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(ExactSmem));
        if (e != cudaSuccess) return e;
        dim3 grid(tiles, nsplit);
        knn_exact_kernel<<<grid, 256, sizeof(ExactSmem), stream>>>(
            d_base, d_queries, d_row_ids, q_begin, q_row0, n_rows, nb, d, kcap, self_join, nsplit,
            split_rows, lists, meta);
        if (launches) ++*launches;
    } else {
        // k == 1 self-join: every row is just [self]; finalize writes it.
        nsplit = 0;
    }
    const int warps = 8;
    knn_finalize_kernel<<<(n_rows + warps - 1) / warps, warps * 32, 0, stream>>>(
        d_row_ids, q_begin, out_q_begin, n_rows, k, kcap > 0 ? kcap : 1, self_join, nsplit, lists,
        meta, scratch, d_assign, d_dis);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace pqh
