// knn_order.cu -- locality order for the tensor-core sweep.
//
// The filter epilogue of knn_tc_kernel keeps, per query row, a threshold tau = "the L-th best
// score seen so far"; a base tile costs that row nothing once no score in it beats tau.  In the
// order the caller's rows come in (compute_nn_fast.c:478-481 hands knn_full_thread whatever
// order the .fvecs file has) near neighbours are spread over the whole base, tau tightens like
// L/n along the sweep, and a third of all (warp, tile) pairs take the slow path -- and because
// an accumulator stage is only free again when all eight filter warps are done with it, one
// slow warp stalls the whole tile.
//
// Here the base rows are grouped into the cells of a coarse quantiser (nearest of C pivot rows),
// the query rows of a launch are grouped the same way, and a query tile starts its sweep at the
// cell its own rows belong to.  A row then meets most of its true neighbours in the first few
// tiles and tau is tight for the rest of the sweep (measured on SIFT1M-shaped data, 4096 cells:
// slow (warp, tile) pairs 33 % -> 0.6 %).  The result is unchanged -- every row is still
// compared with every base row and certified as before -- only the order in which the tiles
// are visited differs.  Candidate ids become positions in the permuted base operand; the
// re-rank kernel maps them back (base_perm) before it orders ties.
//
// The nearest pivot of a row is itself a k = 1 search of the row against the C pivots, so it
// runs on the tensor cores through the same GEMM kernel with an arg-max epilogue (launch_knn_tc
// drives it: pivot operand, one short sweep per row slice; the streaming kernel keeps candidate
// lists and order_pick_kernel takes the best entry of every row's); what is here are the small
// kernels around it: pivot choice and numbering (spatially close cells get neighbouring numbers),
// counting sort by cell, inverse permutation, start tile of every query tile.
#include <cstdlib>

#include "kernels.h"
#include "knn_tc_common.cuh"

namespace pqh {

namespace {

constexpr int ORD_MAX_CELLS = 4096;
constexpr int ORD_MIN_CELLS = 512;  // four base tiles: the shortest sweep the filter kernel takes

// pivot c = base row (c + 1/2) * nb / C
__global__ void order_pivot_ids_kernel(int *__restrict__ ids, int C, int nb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) ids[c] = (int)(((long long)c * 2 + 1) * nb / (2LL * C));
}

// Numbering of the cells.  A sweep starts with a WINDOW of consecutive base tiles at the query
// tile's own cell, so what lies next to a cell in the permuted base should be near it in space
// too.  The pivots are therefore renumbered along a three-level tree: every ORD_GROUP-th pivot
// stands for a group, every pivot joins the group whose representative is nearest to it, every
// group likewise joins one of ORD_TOPS top-level representatives, and the cells are numbered by
// (top, group, original number).  Small, off the tensor cores: C x C/8 x d multiply-adds.
constexpr int ORD_GROUP = 8, ORD_TOPS = 16;

// out[i] = j minimising |x[ids[i * i_stride]] - x[ids[j * ref_stride]]|^2 over j < n_ref (ties:
// lowest j).  One warp per i.
__global__ void __launch_bounds__(256)
order_nearest_kernel(const float *__restrict__ x, int d, const int *__restrict__ ids, int n,
                     int i_stride, int ref_stride, int n_ref, int *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const float *a = x + (size_t)ids[i * i_stride] * d;
    float best = __int_as_float(0x7f800000);
    int bj = 0;
    for (int j0 = 0; j0 < n_ref; j0 += 8) {  // eight candidates at a time: independent chains
        const float *b[8];
        float s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u < n_ref ? j0 + u : n_ref - 1;
            b[u] = x + (size_t)ids[j * ref_stride] * d;
            s[u] = 0.0f;
        }
        for (int t = lane; t < d; t += 32) {
            const float av = a[t];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float df = av - __ldg(b[u] + t);
                s[u] = fmaf(df, df, s[u]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (j0 + u < n_ref && s[u] < best) {
                best = s[u];
                bj = j0 + u;
            }
    }
    if (lane == 0) out[i] = bj;
}

// ids_out = ids_in sorted by (top of the pivot's group, group, pivot number); one block, C <= 4096
__global__ void __launch_bounds__(1024)
order_rank_cells_kernel(const int *__restrict__ group_of, const int *__restrict__ top_of, int C,
                        const int *__restrict__ ids_in, int *__restrict__ ids_out) {
    __shared__ uint32_t key[4096];
    for (int c = threadIdx.x; c < 4096; c += 1024) {
        uint32_t k = 0xFFFFFFFFu;
        if (c < C) {
            const int g = group_of[c];
            k = ((uint32_t)top_of[g] << 24) | ((uint32_t)g << 12) | (uint32_t)c;
        }
        key[c] = k;
    }
    __syncthreads();
    for (int len = 2; len <= 4096; len <<= 1)
        for (int inc = len >> 1; inc > 0; inc >>= 1) {
            for (int t = threadIdx.x; t < 2048; t += 1024) {
                const int lo = ((t / inc) * (inc << 1)) + (t % inc), hi = lo + inc;
                const bool up = (lo & len) == 0;
                const uint32_t a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
            __syncthreads();
        }
    for (int r = threadIdx.x; r < C; r += 1024) ids_out[r] = ids_in[key[r] & 0xFFFu];
}

// cell[r] = position (= pivot number) of the best-scoring entry of row r's candidate lists from
// the sweep over the pivots; hist[cell]++.  Keys order by descending score, so the smallest key
// of a row is its nearest pivot.
__global__ void order_pick_kernel(const key64_t *__restrict__ cand, const int *__restrict__ cnt,
                                  int n_rows, int gcap, int nlists, int tail_row0, int ns_tail,
                                  int *__restrict__ cell, int *__restrict__ hist) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int my_lists = r < tail_row0 ? nlists : ns_tail;
    const size_t slot0 = list_slot(r, 0, tail_row0, nlists, ns_tail);
    key64_t best = KEY_INF;
    for (int s = 0; s < my_lists; ++s) {
        const key64_t *l = cand + (slot0 + s) * gcap;
        const int c = cnt[slot0 + s];
        for (int i = 0; i < c; ++i) {
            const key64_t k = __ldcg(l + i);
            best = k < best ? k : best;
        }
    }
    const int c = best == KEY_INF ? 0 : (int)(uint32_t)best;
    cell[r] = c;
    atomicAdd(hist + c, 1);
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

// qt_start[qt] = window of base tiles at which query tile qt starts its sweep: first tile in the
// low 24 bits, length in the bits above (own_window_len / item_order in knn_tc_common.cuh).  The
// window runs from the first tile of the cell of the tile's first row to the last tile of the
// cell of its last row -- its rows are sorted by cell, so that is where their cell-mates are --
// but is at least min_tiles and at most ORD_MAX_WINDOW tiles long.  (With all base rows as
// queries a tile spans a cell or two; a slice of the rows -- one GPU's share -- spreads a tile
// over many cells: 125k of 1M rows, 8 cells.  A fixed 8-tile window left those rows without their
// neighbours: 9x the slow (warp, tile) pairs.)
constexpr int ORD_MAX_WINDOW = 64;
__global__ void order_qtstart_kernel(const int *__restrict__ q_start, const int *__restrict__ b_start,
                                     int C, int n_qtiles, int n_rows, int min_tiles,
                                     int *__restrict__ qt_start) {
    const int qt = blockIdx.x * blockDim.x + threadIdx.x;
    if (qt >= n_qtiles) return;
    auto cell_of = [&](int r) {  // largest c with q_start[c] <= r
        int lo = 0, hi = C;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (q_start[mid] <= r) lo = mid;
            else hi = mid;
        }
        return lo;
    };
    const int r0 = qt * TC_ROWS;
    const int r1 = r0 + TC_ROWS - 1 < n_rows - 1 ? r0 + TC_ROWS - 1 : n_rows - 1;
    const int c0 = cell_of(r0), c1 = cell_of(r1);
    const int t0 = b_start[c0] / TC_BN;
    const int last_row = b_start[c1 + 1] > b_start[c0] ? b_start[c1 + 1] - 1 : b_start[c0];
    int wlen = last_row / TC_BN - t0 + 1;
    wlen = wlen < min_tiles ? min_tiles : wlen;
    wlen = wlen > ORD_MAX_WINDOW ? ORD_MAX_WINDOW : wlen;
    qt_start[qt] = t0 | (wlen << 24);
}

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// Number of cells: about 32 base rows per cell, between 512 and 4096 (the cell assignment is a
// sweep over the pivots, and 4096 cells are what pays at 1M rows: 244 rows per cell, slow (warp,
// tile) pairs 33 % -> 0.6 %).  Short sweeps want the small cells: a 100k-row base ran its
// GEMM + filter in 3.6 ms with 512 cells and in 2.6 ms with 4096, because only cells this small
// put a row's neighbours into the few tiles its sweep starts with.
// 0: no locality order for a base this small (the sweep is too short to matter).
int order_num_cells(int nb) {
    static const int forced = [] {
        const char *e = getenv("PQH_TC_CELLS");  // experiments
        return e ? atoi(e) : 0;
    }();
    if (nb < 32768) return 0;
    if (forced >= ORD_MIN_CELLS && forced <= ORD_MAX_CELLS) return forced / TC_BN * TC_BN;
    int c = ORD_MIN_CELLS;
    while (c < ORD_MAX_CELLS && (long long)c * 32 < nb) c *= 2;
    return c;
}

size_t order_workspace_bytes(int slice_rows, int nb, int slice_qtiles) {
    const int C = order_num_cells(nb);
    size_t b = 0;
    b += 4 * up256((size_t)C * sizeof(int));           // pivot ids (raw, renumbered), groups, tops
    b += 3 * up256((size_t)nb * sizeof(int));          // cell_b, base_perm, base_inv
    b += 3 * up256((size_t)slice_rows * sizeof(int));  // cell_q, q_perm, q_self
    b += 6 * up256((size_t)(C + 1) * sizeof(int));     // hist / start / cursor, base and queries
    b += up256((size_t)slice_qtiles * sizeof(int));
    return b + 256;
}

void order_carve(void *d_work, int slice_rows, int nb, int slice_qtiles, OrderBufs *w) {
    w->C = order_num_cells(nb);
    unsigned char *p = reinterpret_cast<unsigned char *>(((uintptr_t)d_work + 255) & ~(uintptr_t)255);
    auto take = [&](size_t bytes) {
        void *r = p;
        p += up256(bytes);
        return r;
    };
    w->piv_ids = static_cast<int *>(take((size_t)w->C * sizeof(int)));
    w->piv_raw = static_cast<int *>(take((size_t)w->C * sizeof(int)));
    w->piv_group = static_cast<int *>(take((size_t)w->C * sizeof(int)));
    w->group_top = static_cast<int *>(take((size_t)w->C * sizeof(int)));
    w->cell_b = static_cast<int *>(take((size_t)nb * sizeof(int)));
    w->base_perm = static_cast<int *>(take((size_t)nb * sizeof(int)));
    w->base_inv = static_cast<int *>(take((size_t)nb * sizeof(int)));
    w->cell_q = static_cast<int *>(take((size_t)slice_rows * sizeof(int)));
    w->q_perm = static_cast<int *>(take((size_t)slice_rows * sizeof(int)));
    w->q_self = static_cast<int *>(take((size_t)slice_rows * sizeof(int)));
    w->hist_b = static_cast<int *>(take((size_t)(w->C + 1) * sizeof(int)));
    w->start_b = static_cast<int *>(take((size_t)(w->C + 1) * sizeof(int)));
    w->cur_b = static_cast<int *>(take((size_t)(w->C + 1) * sizeof(int)));
    w->hist_q = static_cast<int *>(take((size_t)(w->C + 1) * sizeof(int)));
    w->start_q = static_cast<int *>(take((size_t)(w->C + 1) * sizeof(int)));
    w->cur_q = static_cast<int *>(take((size_t)(w->C + 1) * sizeof(int)));
    w->qt_start = static_cast<int *>(take((size_t)slice_qtiles * sizeof(int)));
}

// pivot rows, numbered along the group tree; clears the base histogram
cudaError_t launch_order_pivots(const OrderBufs &w, const float *d_base, int nb, int d,
                                cudaStream_t stream, int *launches) {
    static const bool flat = [] {
        const char *e = getenv("PQH_TC_FLAT_CELLS");  // experiments: cells in pivot-row order
        return e && atoi(e) != 0;
    }();
    cudaError_t e = cudaMemsetAsync(w.hist_b, 0, 3 * up256((size_t)(w.C + 1) * sizeof(int)), stream);
    if (e != cudaSuccess) return e;
    const int groups = w.C / ORD_GROUP;
    if (flat || groups < 2 * ORD_TOPS) {
        order_pivot_ids_kernel<<<(w.C + 127) / 128, 128, 0, stream>>>(w.piv_ids, w.C, nb);
        ++*launches;
        return cudaGetLastError();
    }
    order_pivot_ids_kernel<<<(w.C + 127) / 128, 128, 0, stream>>>(w.piv_raw, w.C, nb);
    order_nearest_kernel<<<(w.C + 7) / 8, 256, 0, stream>>>(d_base, d, w.piv_raw, w.C, 1, ORD_GROUP,
                                                            groups, w.piv_group);
    order_nearest_kernel<<<(groups + 7) / 8, 256, 0, stream>>>(
        d_base, d, w.piv_raw, groups, ORD_GROUP, ORD_GROUP * (groups / ORD_TOPS), ORD_TOPS,
        w.group_top);
    order_rank_cells_kernel<<<1, 1024, 0, stream>>>(w.piv_group, w.group_top, w.C, w.piv_raw,
                                                    w.piv_ids);
    *launches += 4;
    return cudaGetLastError();
}

// cells of the rows of one assignment sweep, from its candidate lists
cudaError_t launch_order_pick(const void *cand, const int *cnt, int n_rows, int gcap, int nlists,
                              int tail_row0, int ns_tail, int *cell, int *hist,
                              cudaStream_t stream, int *launches) {
    order_pick_kernel<<<(n_rows + 127) / 128, 128, 0, stream>>>(
        static_cast<const key64_t *>(cand), cnt, n_rows, gcap, nlists, tail_row0, ns_tail, cell,
        hist);
    ++*launches;
    return cudaGetLastError();
}

// Once per search, after every base row has its cell: the permuted base order and its inverse.
cudaError_t launch_order_sort_base(const OrderBufs &w, int nb, cudaStream_t stream, int *launches) {
    order_scan_kernel<<<1, 1024, 0, stream>>>(w.hist_b, w.C, w.start_b, w.cur_b);
    order_scatter_kernel<<<(nb + 255) / 256, 256, 0, stream>>>(w.cell_b, nb, 0, w.cur_b,
                                                               w.base_perm, w.base_inv);
    *launches += 2;
    return cudaGetLastError();
}

// clears the query histogram of a slice (before launch_order_pick fills it for queries that are
// not base rows)
cudaError_t launch_order_clear_queries(const OrderBufs &w, cudaStream_t stream) {
    return cudaMemsetAsync(w.hist_q, 0, 3 * up256((size_t)(w.C + 1) * sizeof(int)), stream);
}

// Once per row slice: the slice's query rows [q_begin, q_begin + n_rows) grouped by cell.
// Self-join: query q is base row q and shares its cell; otherwise cell_q / hist_q were filled by
// launch_order_pick.
cudaError_t launch_order_sort_queries(const OrderBufs &w, int q_begin, int n_rows, int self_join,
                                      int n_qtiles, int own_tiles, cudaStream_t stream,
                                      int *launches, OrderOut *out) {
    const int *cq = w.cell_q;
    if (self_join) {
        cudaError_t e = launch_order_clear_queries(w, stream);
        if (e != cudaSuccess) return e;
        cq = w.cell_b + q_begin;
        order_count_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(cq, n_rows, w.hist_q);
        ++*launches;
    }
    order_scan_kernel<<<1, 1024, 0, stream>>>(w.hist_q, w.C, w.start_q, w.cur_q);
    order_scatter_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(cq, n_rows, q_begin, w.cur_q,
                                                                   w.q_perm, nullptr);
    *launches += 2;
    if (self_join) {
        order_qself_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(w.q_perm, n_rows, w.base_inv,
                                                                     w.q_self);
        ++*launches;
    }
    order_qtstart_kernel<<<(n_qtiles + 255) / 256, 256, 0, stream>>>(
        w.start_q, w.start_b, w.C, n_qtiles, n_rows, own_tiles, w.qt_start);
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
