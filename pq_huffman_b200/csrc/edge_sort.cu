// edge_sort.cu -- stable sort of the kNN graph's edges for the first consumer of the graph.
//
// The reference's mst_builder turns nn_indices / nn_dist into N*k records
// {int source, int target, float dist} (mst.h:12-16) and sorts them twice with libc qsort: by
// distance before Kruskal's pass (mst.c:175, mst_dist_comparator, mst.c:15-26) and, for the edges
// that made it into the forest, by source (mst.c:214, mst_source_comparator, mst.c:28-32).  With
// the kNN search itself down to a fraction of a second, the first of these is the longest step
// of the pipeline (16M records at 1M x 16, 320M at Deep10M x 32).
//
// glibc's qsort is a merge sort (2.39: stdlib/qsort.c, stable whenever its scratch buffer can be
// allocated), so for a comparator that is a strict weak order its result is "ascending, equal
// keys in their original order".  That is what runs here: a least-significant-digit radix sort
// (four 8-bit passes) of (key, original position) pairs, each pass a counting sort whose scatter
// keeps the order of equal digits, then one gather of the 12-byte records.
//
// mst_dist_comparator calls two distances equal when |a - b| < 1e-9 (in float arithmetic), which
// is NOT a strict weak order in general: the sort is only taken when, in the sorted sequence,
// every two neighbouring DIFFERENT keys are at least 1e-9 apart (float subtraction is monotone,
// so all other pairs then are too) and no key is NaN -- then "equal" means "same value" and the
// comparator is the plain order of the floats.  Otherwise the flags word says so and the
// caller keeps its own qsort (squared distances of real data are far above that granularity;
// exact duplicates, distance 0, are equal under both readings).
#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace pqh {

namespace {

constexpr int ES_THREADS = 256;
constexpr int ES_WARPS = ES_THREADS / 32;

struct Edge {
    int source, target;
    float dist;
};

// key of record i: mode 0 = distance (ascending floats as unsigned integers, -0 == +0),
// mode 1 = source (signed).  flags: bit 0 = a NaN distance was seen.
__global__ void __launch_bounds__(ES_THREADS)
es_keys_kernel(const Edge *__restrict__ e, size_t n, int mode, uint32_t *__restrict__ key,
               uint32_t *__restrict__ idx, uint32_t *__restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k;
    if (mode == 0) {
        float d = e[i].dist;
        if (d != d) atomicOr(flags, 1u);
        if (d == 0.0f) d = 0.0f;  // -0 -> +0: the comparator cannot tell them apart
        k = float_to_ordered(d);
    } else {
        k = (uint32_t)e[i].source ^ 0x80000000u;
    }
    key[i] = k;
    idx[i] = (uint32_t)i;
}

// counting-sort pass, step 1: digit counts of every block's contiguous chunk.
// hist[digit * gridDim.x + block]
__global__ void __launch_bounds__(ES_THREADS)
es_hist_kernel(const uint32_t *__restrict__ key, size_t n, size_t chunk, int shift,
               uint32_t *__restrict__ hist) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t b0 = (size_t)blockIdx.x * chunk;
    const size_t b1 = b0 + chunk < n ? b0 + chunk : n;
    for (size_t i = b0 + threadIdx.x; i < b1; i += ES_THREADS)
        atomicAdd(&cnt[(key[i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = cnt[threadIdx.x];
}

// step 2: exclusive scan of hist (digit-major, block-minor) in two small kernels: block d scans
// the counts of digit d over the chunks (one per thread, G <= 1024) and leaves the digit's total
// in tot[d]; then one block scans the 256 totals.  (A single block walking all 256 * G counts
// took 0.45 ms per pass -- more than the scatter itself.)
__global__ void __launch_bounds__(1024)
es_scan_digit_kernel(uint32_t *__restrict__ hist, int G, uint32_t *__restrict__ tot) {
    __shared__ uint32_t wsum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t *h = hist + (size_t)blockIdx.x * G;
    const uint32_t v = t < G ? h[t] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    const uint32_t before = warp ? wsum[warp - 1] : 0u;
    if (t < G) h[t] = before + inc - v;
    if (t == 1023) tot[blockIdx.x] = wsum[31];
}
__global__ void __launch_bounds__(256) es_scan_tot_kernel(uint32_t *__restrict__ tot) {
    __shared__ uint32_t part[256];
    const int t = threadIdx.x;
    const uint32_t v = tot[t];
    part[t] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const uint32_t u = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += u;
        __syncthreads();
    }
    tot[t] = part[t] - v;
}

// step 3: stable scatter.  A block walks its chunk tile by tile (256 records, in order); within a
// tile a record's place among the records with the same digit is (same digit in earlier warps) +
// (same digit in earlier lanes of its warp), found with match_any; base[] carries each digit's
// next free position from tile to tile.
__global__ void __launch_bounds__(ES_THREADS)
es_scatter_kernel(const uint32_t *__restrict__ key_in, const uint32_t *__restrict__ idx_in, size_t n,
                  size_t chunk, int shift, const uint32_t *__restrict__ offs,
                  const uint32_t *__restrict__ tot, uint32_t *__restrict__ key_out,
                  uint32_t *__restrict__ idx_out) {
    __shared__ uint32_t base[256];
    __shared__ uint16_t wcnt[ES_WARPS][256];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    base[t] = tot[t] + offs[(size_t)t * gridDim.x + blockIdx.x];
    for (int w = 0; w < ES_WARPS; ++w) wcnt[w][t] = 0;
    __syncthreads();
    const size_t b0 = (size_t)blockIdx.x * chunk;
    const size_t b1 = b0 + chunk < n ? b0 + chunk : n;
    for (size_t tile = b0; tile < b1; tile += ES_THREADS) {
        const size_t i = tile + t;
        const bool active = i < b1;
        uint32_t k = 0, v = 0, d = 0, rank = 0;
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            k = key_in[i];
            v = idx_in[i];
            d = (k >> shift) & 255u;
            const unsigned peers = __match_any_sync(amask, d);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0) wcnt[warp][d] = (uint16_t)__popc(peers);
        }
        __syncthreads();
        if (active) {
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += wcnt[w][d];
            const uint32_t pos = base[d] + before + rank;
            key_out[pos] = k;
            idx_out[pos] = v;
        }
        __syncthreads();
        uint32_t placed = 0;
        for (int w = 0; w < ES_WARPS; ++w) {
            placed += wcnt[w][t];
            wcnt[w][t] = 0;
        }
        base[t] += placed;
        __syncthreads();
    }
}

// Is mst_dist_comparator the plain order of the floats on this data?  flags bit 1: two different
// neighbouring keys of the sorted sequence are closer than 1e-9 (as the comparator computes it:
// float difference, |.| compared in double).
__global__ void __launch_bounds__(ES_THREADS)
es_gap_kernel(const uint32_t *__restrict__ key, size_t n, uint32_t *__restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= n) return;
    const uint32_t k0 = key[i - 1], k1 = key[i];
    if (k0 == k1) return;
    const float diff = ordered_to_float(k1) - ordered_to_float(k0);
    if (fabs((double)diff) < 1e-9) atomicOr(flags, 2u);
}

__global__ void __launch_bounds__(ES_THREADS)
es_gather_kernel(const Edge *__restrict__ in, const uint32_t *__restrict__ idx, size_t n,
                 Edge *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Edge e = in[idx[i]];
    out[i] = e;
}

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

int es_blocks(size_t n) {
    // chunks of at least 4096 records, at most 1024 blocks (the scan takes one count per thread)
    size_t b = (n + 4095) / 4096;
    if (b > 1024) b = 1024;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

size_t edge_sort_workspace_bytes(size_t n) {
    return 4 * up256(n * sizeof(uint32_t)) + up256((size_t)(256 * 1024 + 256) * sizeof(uint32_t)) + 1024;
}

// Sorts the n records at d_in (12 bytes each) into d_out; *d_flags (device word, zeroed here)
// receives bit 0 = NaN key, bit 1 = keys closer than the comparator's tolerance: with either set
// the result must not be used (mode 0).  Everything is enqueued on `stream`.
cudaError_t launch_edge_sort(const void *d_in, void *d_out, size_t n, int mode, void *d_work,
                             uint32_t *d_flags, cudaStream_t stream, int *launches) {
    unsigned char *w = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(d_work) + 255) & ~(uintptr_t)255);
    uint32_t *key[2], *idx[2];
    key[0] = reinterpret_cast<uint32_t *>(w);
    key[1] = reinterpret_cast<uint32_t *>(w + up256(n * 4));
    idx[0] = reinterpret_cast<uint32_t *>(w + 2 * up256(n * 4));
    idx[1] = reinterpret_cast<uint32_t *>(w + 3 * up256(n * 4));
    uint32_t *hist = reinterpret_cast<uint32_t *>(w + 4 * up256(n * 4));
    uint32_t *tot = hist + (size_t)256 * 1024;
    cudaError_t e = cudaMemsetAsync(d_flags, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess || n == 0) return e;
    const Edge *in = static_cast<const Edge *>(d_in);
    const int eb = (int)((n + ES_THREADS - 1) / ES_THREADS);
    es_keys_kernel<<<eb, ES_THREADS, 0, stream>>>(in, n, mode, key[0], idx[0], d_flags);
    ++*launches;
    const int G = es_blocks(n);
    const size_t chunk = ((n + G - 1) / G + ES_THREADS - 1) / ES_THREADS * ES_THREADS;
    int cur = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        es_hist_kernel<<<G, ES_THREADS, 0, stream>>>(key[cur], n, chunk, shift, hist);
        es_scan_digit_kernel<<<256, 1024, 0, stream>>>(hist, G, tot);
        es_scan_tot_kernel<<<1, 256, 0, stream>>>(tot);
        es_scatter_kernel<<<G, ES_THREADS, 0, stream>>>(key[cur], idx[cur], n, chunk, shift, hist,
                                                        tot, key[cur ^ 1], idx[cur ^ 1]);
        cur ^= 1;
        *launches += 4;
    }
    if (mode == 0) {
        es_gap_kernel<<<eb, ES_THREADS, 0, stream>>>(key[cur], n, d_flags);
        ++*launches;
    }
    es_gather_kernel<<<eb, ES_THREADS, 0, stream>>>(in, idx[cur], n, static_cast<Edge *>(d_out));
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pqh
