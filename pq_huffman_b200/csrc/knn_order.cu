// knn_order.cu -- locality order for the tensor-core sweep.
//
// The filter epilogue of knn_tc_kernel keeps, per query row, a threshold tau = "the L-th best
// score seen so far"; a base tile costs that row nothing once no score in it beats tau.  In the
// order the caller's rows come in (compute_nn_fast.c:478-481 hands knn_full_thread whatever
// order the .fvecs file has) near neighbours are spread over the whole base, tau tightens like
// L/n along the sweep, and a third of all (warp, tile) pairs take the slow path.
//
// Here the base rows are grouped into cells of a coarse quantiser (nearest of C pivot rows,
// measured on up to 32 sampled coordinates -- approximate on purpose, it only steers the
// ORDER), the query rows of the launch are grouped the same way, and a query tile starts its
// sweep at the cell its own rows belong to.  A row then meets most of its true neighbours in
// the first few tiles and tau is tight for the rest of the sweep; the result is unchanged --
// every row is still compared with every base row and certified as before -- only the order
// in which the tiles are visited differs.  Candidate ids become positions in the permuted
// base operand; the re-rank kernel maps them back (base_perm) before it orders ties.
//
// Cost: n x C x 32 FMAs + three passes over n integers (1M x 128: 1024 cells, < 2 ms).
#include <cstdlib>

#include "kernels.h"
#include "knn_tc_common.cuh"

namespace pqh {

namespace {

constexpr int ORD_DSUB = 128;      // coordinates used for the cell assignment (sampled if d is larger)
constexpr int ORD_PCHUNK = 64;     // pivots per shared-memory chunk
constexpr int ORD_MAX_CELLS = 4096;

// pivot c = base row (c + 1/2) * nb / C, restricted to the sampled coordinates; pn[c] = |p|^2
__global__ void order_pivots_kernel(const float *__restrict__ base, int nb, int d, int dsub,
                                    int C, float *__restrict__ piv, float *__restrict__ pn) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const size_t row = (size_t)(((long long)c * 2 + 1) * nb / (2LL * C));
    float s = 0.0f;
    for (int j = 0; j < ORD_DSUB; ++j) {
        float v = 0.0f;
        if (j < dsub) v = base[row * d + (size_t)((long long)j * d / dsub)];
        piv[(size_t)c * ORD_DSUB + j] = v;
        s += v * v;
    }
    pn[c] = s;
}

// cell[i] = argmin_c |x_i - p_c|^2 over the sampled coordinates (= argmin |p|^2 - 2 x.p);
// hist[cell]++.  One thread per row, the pivots streamed through shared memory.
__global__ void __launch_bounds__(128)
order_assign_kernel(const float *__restrict__ x, int n, int d, int dsub, int C,
                    const float *__restrict__ piv, const float *__restrict__ pn,
                    int *__restrict__ cell, int *__restrict__ hist) {
    __shared__ __align__(16) float sp[ORD_PCHUNK * ORD_DSUB];
    __shared__ float sn[ORD_PCHUNK];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float xv[ORD_DSUB];
#pragma unroll
    for (int j = 0; j < ORD_DSUB; ++j) {
        xv[j] = 0.0f;
        if (i < n && j < dsub) xv[j] = __ldg(x + (size_t)i * d + (size_t)((long long)j * d / dsub));
    }
    float best = __int_as_float(0x7f800000);
    int bc = 0;
    for (int c0 = 0; c0 < C; c0 += ORD_PCHUNK) {
        __syncthreads();
        const int nc = C - c0 < ORD_PCHUNK ? C - c0 : ORD_PCHUNK;
        for (int t = threadIdx.x; t < nc * ORD_DSUB; t += blockDim.x)
            sp[t] = piv[(size_t)c0 * ORD_DSUB + t];
        for (int t = threadIdx.x; t < nc; t += blockDim.x) sn[t] = pn[c0 + t];
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            const float4 *p4 = reinterpret_cast<const float4 *>(sp + c * ORD_DSUB);
            float dot = 0.0f;
#pragma unroll
            for (int j = 0; j < ORD_DSUB / 4; ++j) {
                const float4 p = p4[j];
                dot = fmaf(xv[4 * j + 0], p.x, dot);
                dot = fmaf(xv[4 * j + 1], p.y, dot);
                dot = fmaf(xv[4 * j + 2], p.z, dot);
                dot = fmaf(xv[4 * j + 3], p.w, dot);
            }
            const float s = fmaf(-2.0f, dot, sn[c]);
            if (s < best) {
                best = s;
                bc = c0 + c;
            }
        }
    }
    if (i < n) {
        cell[i] = bc;
        atomicAdd(hist + bc, 1);
    }
}

// the query rows of a self-join share the base rows' cells: only count them
__global__ void order_count_kernel(const int *__restrict__ cell, int n, int *__restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + cell[i], 1);
}

// exclusive scan of hist[0..C) -> start[0..C] and cursor[0..C) (one block, C <= 4096)
__global__ void __launch_bounds__(1024)
order_scan_kernel(const int *__restrict__ hist, int C, int *__restrict__ start,
                  int *__restrict__ cursor) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = (C + 1023) / 1024;
    int loc[ORD_MAX_CELLS / 1024];
    int s = 0;
    for (int j = 0; j < per; ++j) {
        const int c = t * per + j;
        loc[j] = s;
        s += c < C ? hist[c] : 0;
    }
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const int base = part[t] - s;
    for (int j = 0; j < per; ++j) {
        const int c = t * per + j;
        if (c < C) {
            start[c] = base + loc[j];
            cursor[c] = base + loc[j];
        }
    }
    if (t == 1023) start[C] = part[1023];
}

// counting-sort scatter: perm[pos] = id0 + i for the rows of each cell (order inside a cell is
// whatever the atomics give; nothing depends on it); inv[i] = pos if asked for
__global__ void order_scatter_kernel(const int *__restrict__ cell, int n, int id0,
                                     int *__restrict__ cursor, int *__restrict__ perm,
                                     int *__restrict__ inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = atomicAdd(cursor + cell[i], 1);
    perm[pos] = id0 + i;
    if (inv) inv[i] = pos;
}

// q_self[r] = position of query r's own vector in the permuted base (self-join)
__global__ void order_qself_kernel(const int *__restrict__ q_ids, int n, const int *__restrict__ inv,
                                   int *__restrict__ q_self) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) q_self[r] = inv[q_ids[r]];
}

// qt_start[qt] = first base tile of the cell that query row qt * 256 (permuted order) belongs to
__global__ void order_qtstart_kernel(const int *__restrict__ q_start, const int *__restrict__ b_start,
                                     int C, int n_qtiles, int *__restrict__ qt_start) {
    const int qt = blockIdx.x * blockDim.x + threadIdx.x;
    if (qt >= n_qtiles) return;
    const int r = qt * TC_ROWS;
    int lo = 0, hi = C;  // largest c with q_start[c] <= r
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (q_start[mid] <= r) lo = mid;
        else hi = mid;
    }
    qt_start[qt] = b_start[lo] / TC_BN;
}

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int order_num_cells(int nb) {
    static const int forced = [] {
        const char *e = getenv("PQH_TC_CELLS");  // experiments
        return e ? atoi(e) : 0;
    }();
    if (forced >= 16 && forced <= ORD_MAX_CELLS) return forced;
    int c = 16;
    while (c < ORD_MAX_CELLS && (long long)c * 1024 < nb) c *= 2;
    return c;
}

size_t order_workspace_bytes(int slice_rows, int nb, int slice_qtiles) {
    const int C = order_num_cells(nb);
    size_t b = 0;
    b += up256((size_t)C * ORD_DSUB * sizeof(float)) + up256((size_t)C * sizeof(float));  // pivots
    b += 3 * up256((size_t)nb * sizeof(int));          // cell_b, base_perm, base_inv
    b += 3 * up256((size_t)slice_rows * sizeof(int));  // cell_q, q_perm, q_self
    b += 5 * up256((size_t)(C + 1) * sizeof(int));     // hist_b, start_b, cur_b | hist_q, start_q
    b += up256((size_t)(C + 1) * sizeof(int));         // cur_q
    b += up256((size_t)slice_qtiles * sizeof(int));
    return b + 256;
}

namespace {
struct OrderWork {
    int C, dsub;
    float *piv, *pn;
    int *cell_b, *base_perm, *base_inv, *cell_q, *q_perm, *q_self;
    int *hist_b, *start_b, *cur_b, *hist_q, *start_q, *cur_q, *qt_start;
};
OrderWork carve(void *d_work, int slice_rows, int nb, int d, int slice_qtiles) {
    OrderWork w;
    w.C = order_num_cells(nb);
    static const int forced_dims = [] {
        const char *e = getenv("PQH_TC_ODIMS");  // experiments
        return e ? atoi(e) : 0;
    }();
    w.dsub = d < ORD_DSUB ? d : ORD_DSUB;
    if (forced_dims >= 1 && forced_dims < w.dsub) w.dsub = forced_dims;
    unsigned char *p = reinterpret_cast<unsigned char *>(((uintptr_t)d_work + 255) & ~(uintptr_t)255);
    auto take = [&](size_t bytes) {
        void *r = p;
        p += up256(bytes);
        return r;
    };
    w.piv = static_cast<float *>(take((size_t)w.C * ORD_DSUB * sizeof(float)));
    w.pn = static_cast<float *>(take((size_t)w.C * sizeof(float)));
    w.cell_b = static_cast<int *>(take((size_t)nb * sizeof(int)));
    w.base_perm = static_cast<int *>(take((size_t)nb * sizeof(int)));
    w.base_inv = static_cast<int *>(take((size_t)nb * sizeof(int)));
    w.cell_q = static_cast<int *>(take((size_t)slice_rows * sizeof(int)));
    w.q_perm = static_cast<int *>(take((size_t)slice_rows * sizeof(int)));
    w.q_self = static_cast<int *>(take((size_t)slice_rows * sizeof(int)));
    w.hist_b = static_cast<int *>(take((size_t)(w.C + 1) * sizeof(int)));
    w.start_b = static_cast<int *>(take((size_t)(w.C + 1) * sizeof(int)));
    w.cur_b = static_cast<int *>(take((size_t)(w.C + 1) * sizeof(int)));
    w.hist_q = static_cast<int *>(take((size_t)(w.C + 1) * sizeof(int)));
    w.start_q = static_cast<int *>(take((size_t)(w.C + 1) * sizeof(int)));
    w.cur_q = static_cast<int *>(take((size_t)(w.C + 1) * sizeof(int)));
    w.qt_start = static_cast<int *>(take((size_t)slice_qtiles * sizeof(int)));
    return w;
}
}  // namespace

// Once per search: cells of the base rows, the permuted base order and its inverse.
cudaError_t launch_order_base(const float *d_base, int nb, int d, int slice_rows_max,
                              int slice_qtiles_max, void *d_work, cudaStream_t stream,
                              int *launches, OrderOut *out) {
    const OrderWork w = carve(d_work, slice_rows_max, nb, d, slice_qtiles_max);
    cudaError_t e;
    // hist_b, start_b, cur_b are contiguous: one memset
    if ((e = cudaMemsetAsync(w.hist_b, 0, 3 * up256((size_t)(w.C + 1) * sizeof(int)), stream)) !=
        cudaSuccess)
        return e;
    order_pivots_kernel<<<(w.C + 127) / 128, 128, 0, stream>>>(d_base, nb, d, w.dsub, w.C, w.piv,
                                                               w.pn);
    order_assign_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(d_base, nb, d, w.dsub, w.C, w.piv,
                                                              w.pn, w.cell_b, w.hist_b);
    order_scan_kernel<<<1, 1024, 0, stream>>>(w.hist_b, w.C, w.start_b, w.cur_b);
    order_scatter_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(w.cell_b, nb, 0, w.cur_b,
                                                               w.base_perm, w.base_inv);
    *launches += 4;
    out->base_perm = w.base_perm;
    out->base_inv = w.base_inv;
    out->q_perm = nullptr;
    out->q_self = nullptr;
    out->qt_start = nullptr;
    return cudaGetLastError();
}

// Once per row slice: the slice's query rows [q_begin, q_begin + n_rows) grouped by cell.
// d_queries: the FP32 rows of the slice (row r is query q_begin + r) when they are not base
// rows; null for a self-join (query q is base row q).
cudaError_t launch_order_queries(const float *d_queries, int q_begin, int n_rows, int nb, int d,
                                 int self_join, int n_qtiles, int slice_rows_max,
                                 int slice_qtiles_max, void *d_work, cudaStream_t stream,
                                 int *launches, OrderOut *out) {
    const OrderWork w = carve(d_work, slice_rows_max, nb, d, slice_qtiles_max);
    cudaError_t e;
    if ((e = cudaMemsetAsync(w.hist_q, 0, 3 * up256((size_t)(w.C + 1) * sizeof(int)), stream)) !=
        cudaSuccess)
        return e;
    const int *cq = w.cell_q;
    if (self_join) {
        cq = w.cell_b + q_begin;
        order_count_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(cq, n_rows, w.hist_q);
    } else {
        order_assign_kernel<<<(n_rows + 127) / 128, 128, 0, stream>>>(
            d_queries, n_rows, d, w.dsub, w.C, w.piv, w.pn, w.cell_q, w.hist_q);
    }
    order_scan_kernel<<<1, 1024, 0, stream>>>(w.hist_q, w.C, w.start_q, w.cur_q);
    order_scatter_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(cq, n_rows, q_begin, w.cur_q,
                                                                   w.q_perm, nullptr);
    *launches += 3;
    if (self_join) {
        order_qself_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(w.q_perm, n_rows, w.base_inv,
                                                                     w.q_self);
        ++*launches;
    }
    order_qtstart_kernel<<<(n_qtiles + 255) / 256, 256, 0, stream>>>(w.start_q, w.start_b, w.C,
                                                                     n_qtiles, w.qt_start);
    ++*launches;
    out->base_perm = w.base_perm;
    out->base_inv = w.base_inv;
    out->q_perm = w.q_perm;
    out->q_self = self_join ? w.q_self : nullptr;
    out->qt_start = w.qt_start;
    return cudaGetLastError();
}

// second pass: own-column positions of rows named by absolute id
cudaError_t launch_order_self_of(const int *d_ids, int n, const int *base_inv, int *d_q_self,
                                 cudaStream_t stream, int *launches) {
    order_qself_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_ids, n, base_inv, d_q_self);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pqh
