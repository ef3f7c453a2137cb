// pqh_knn.cu -- context management and the C ABI of libpqh_knn.so (include/pqh_knn.h).
//
// The exported knn_full_thread replaces Yael's symbol of the same name at the
// reference's only call site (compute_nn_fast.c:478-481).  The host side of the
// reference calls it once per block, serially, from its main thread, with
// pageable malloc'ed buffers it owns; the context below is
// created on first use and keeps streams, device buffers and workspace alive
// across those calls.  There is no CPU fallback: without a usable sm_100 device
// every entry point fails (knn_full_thread: message + exit(1), the reference's
// error convention, compute_nn_fast.c:154-156).
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pqh_knn.h"
#include "kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                       \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(e__ == cudaErrorMemoryAllocation ? PQH_ENOMEM : PQH_ECUDA,         \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                         \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return PQH_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(PQH_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes,
                        cudaGetErrorString(e));
        }
        cap = want;
        return PQH_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceCtx {
    int dev = -1;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t kev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf base, queries, assign, dis, work;
    // resident data set of the block searches (pqh_knn_dataset_*)
    DevBuf dataset, block_ids;
    int ds_n = 0, ds_d = 0;
};

struct Context {
    std::vector<DeviceCtx> devs;
    int algo = PQH_ALGO_AUTO;
    pqh_knn_stats_t stats;
};

std::mutex g_mu;
// one lock per device slot: block searches on different devices run concurrently, everything
// that touches one device's buffers is serialised (lock order: g_mu, then g_dev_mu[slot])
constexpr int MAX_DEVS = 16;
std::mutex g_dev_mu[MAX_DEVS];
Context *g_ctx = nullptr;
int g_algo = PQH_ALGO_AUTO;

int parse_env_devices(std::vector<int> &out) {
    const char *s = getenv("PQH_DEVICES");
    if (!s || !*s) {
        out.push_back(0);
        return PQH_OK;
    }
    const char *p = s;
    while (*p) {
        char *end = nullptr;
        long v = strtol(p, &end, 10);
        if (end == p) return fail(PQH_EINVAL, "PQH_DEVICES: cannot parse '%s'", s);
        out.push_back((int)v);
        p = end;
        while (*p == ',' || *p == ' ') ++p;
    }
    return PQH_OK;
}

void destroy_ctx(Context *c) {
    if (!c) return;
    for (auto &d : c->devs) {
        if (cudaSetDevice(d.dev) != cudaSuccess) continue;
        if (d.stream) cudaStreamSynchronize(d.stream);
        d.base.release();
        d.queries.release();
        d.assign.release();
        d.dis.release();
        d.work.release();
        d.dataset.release();
        d.block_ids.release();
        for (auto &e : d.ev)
            if (e) cudaEventDestroy(e);
        for (auto &e : d.kev)
            if (e) cudaEventDestroy(e);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete c;
}

int create_ctx(const int *devices, int n, Context **out) {
    std::vector<int> ids;
    if (devices && n > 0) {
        ids.assign(devices, devices + n);
    } else {
        int rc = parse_env_devices(ids);
        if (rc) return rc;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(PQH_ENODEV, "no CUDA device available (%s); libpqh_knn has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if ((int)ids.size() > MAX_DEVS)
        return fail(PQH_EINVAL, "at most %d devices per context", MAX_DEVS);
    Context *c = new Context();
    memset(&c->stats, 0, sizeof(c->stats));
    for (int id : ids) {
        if (id < 0 || id >= count) {
            destroy_ctx(c);
            return fail(PQH_ENODEV, "device %d requested but only %d visible", id, count);
        }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, id) != cudaSuccess || prop.major != 10) {
            destroy_ctx(c);
            return fail(PQH_ENODEV,
                        "device %d is not an sm_100 (Blackwell B200) GPU; this library ships "
                        "sm_100a code only and has no fallback",
                        id);
        }
        DeviceCtx d;
        d.dev = id;
        d.num_sms = prop.multiProcessorCount;
        if (cudaSetDevice(id) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) {
            destroy_ctx(c);
            return fail(PQH_ECUDA, "cannot create a stream on device %d: %s", id,
                        cudaGetErrorString(cudaGetLastError()));
        }
        for (auto &ev : d.ev) cudaEventCreate(&ev);
        for (auto &ev : d.kev) cudaEventCreate(&ev);
        c->devs.push_back(d);
    }
    *out = c;
    return PQH_OK;
}

int get_ctx(Context **out) {
    if (!g_ctx) {
        int rc = create_ctx(nullptr, 0, &g_ctx);
        if (rc) return rc;
    }
    *out = g_ctx;
    return PQH_OK;
}

int check_args(int nq_total, int nb, int d, int k, const void *b, const void *v, int q_begin,
               int q_end, const void *assign, const void *dis) {
    if (nq_total < 0 || nb < 0 || d <= 0 || k < 0)
        return fail(PQH_EINVAL, "bad sizes nq=%d nb=%d d=%d k=%d", nq_total, nb, d, k);
    if (q_begin < 0 || q_end > nq_total || q_begin > q_end)
        return fail(PQH_EINVAL, "bad row range [%d,%d) of %d", q_begin, q_end, nq_total);
    if (q_end > q_begin && k > 0 && (!assign || !dis))
        return fail(PQH_EINVAL, "null output pointer");
    if (q_end > q_begin && nb > 0 && (!b || !v)) return fail(PQH_EINVAL, "null input pointer");
    return PQH_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

size_t exact_ws(int n_rows, int nb, int k, int num_sms) {
    (void)nb;
    return pqh::exact_workspace_bound(n_rows, k > 1 ? k : 1, num_sms);
}

size_t align1k(size_t x) { return (x + 1023) & ~(size_t)1023; }

size_t workspace_bytes(int n_rows, int nb, int d, int k, int self_join, int num_sms) {
    size_t ex = exact_ws(n_rows, nb, k, num_sms);
    if (!pqh::tc_supported(n_rows, nb, d, k, self_join)) return ex + 1024;
    // tensor-core scratch, then room for the exact kernel on the rows it hands back
    return align1k(pqh::tc_workspace_bytes(n_rows, nb, d, k, self_join, num_sms)) + ex + 2048;
}

// Search rows [q_begin,q_end) with all pointers already on device `dc`.
// Query q's vector is row (q - q_row0) of d_v; its output row is (q - q_begin).
int search_on_device(DeviceCtx &dc, int algo, int nb, int d, int k, const float *d_b,
                     const float *d_v, int q_row0, int self_join, int q_begin, int q_end,
                     int *d_assign, float *d_dis, void *d_work, size_t work_bytes,
                     cudaStream_t stream, pqh_knn_stats_t *st, bool sync_for_stats) {
    const int n_rows = q_end - q_begin;
    if (n_rows <= 0 || k <= 0) return PQH_OK;
    int launches = 0;
    long long rows_tc = 0, rows_ex = 0;
    bool exact_timed = false;
    if (algo == PQH_ALGO_AUTO && pqh::tc_supported(n_rows, nb, d, k, self_join)) {
        const size_t tcb = align1k(pqh::tc_workspace_bytes(n_rows, nb, d, k, self_join, dc.num_sms));
        if (tcb > work_bytes) return fail(PQH_EINVAL, "workspace too small");
        int *d_fb_rows = nullptr;
        unsigned n_fb = 0;
        pqh::TcTimes tt;
        CU_TRY(pqh::launch_knn_tc(d_b, d_v, q_row0, q_begin, q_begin, n_rows, nb, d, k, self_join,
                                  d_assign, d_dis, d_work, tcb, dc.num_sms, stream, &launches,
                                  &d_fb_rows, &n_fb, dc.kev, &tt));
        if (st) {
            st->ms_prep += tt.ms_prep;
            st->ms_gemm += tt.ms_gemm;
            st->ms_rerank += tt.ms_rerank;
            st->candidates += tt.candidates;
            if (tt.eps_ratio_max > st->eps_ratio_max) st->eps_ratio_max = tt.eps_ratio_max;
            st->inexact_operands = tt.inexact;
            st->rows_second_pass += tt.rows_pass2;
        }
        rows_tc = n_rows - (long long)n_fb;
        if (n_fb > 0) {
            rows_ex = n_fb;
            CU_TRY(cudaEventRecord(dc.kev[4], stream));
            CU_TRY(pqh::launch_knn_exact(d_b, d_v, d_fb_rows, 0, q_row0, q_begin, (int)n_fb, nb, d,
                                         k, self_join, d_assign, d_dis,
                                         static_cast<unsigned char *>(d_work) + tcb,
                                         work_bytes - tcb, dc.num_sms, stream, &launches));
            CU_TRY(cudaEventRecord(dc.kev[5], stream));
            exact_timed = true;
        }
    } else {
        rows_ex = n_rows;
        CU_TRY(cudaEventRecord(dc.kev[4], stream));
        CU_TRY(pqh::launch_knn_exact(d_b, d_v, nullptr, q_begin, q_row0, q_begin, n_rows, nb, d, k,
                                     self_join, d_assign, d_dis, d_work, work_bytes, dc.num_sms,
                                     stream, &launches));
        CU_TRY(cudaEventRecord(dc.kev[5], stream));
        exact_timed = true;
    }
    if (exact_timed && st && sync_for_stats) {
        CU_TRY(cudaStreamSynchronize(stream));
        st->ms_exact += ev_ms(dc.kev[4], dc.kev[5]);
    }
    if (st) {
        st->rows += n_rows;
        st->rows_tensor += rows_tc;
        st->rows_exact += rows_ex;
        st->launches += launches;
    }
    return PQH_OK;
}

// Host-pointer search on one device.
int search_rows_host(Context *c, DeviceCtx &dc, int nq_total, int nb, int d, int k, const float *b,
                     size_t ldb, const float *v, size_t ldv, int self_join, int q_begin, int q_end,
                     int *assign, float *dis, pqh_knn_stats_t *st) {
    const int n_rows = q_end - q_begin;
    if (n_rows <= 0 || k <= 0) return PQH_OK;
    std::lock_guard<std::mutex> dev_lock(g_dev_mu[&dc - c->devs.data()]);
    CU_TRY(cudaSetDevice(dc.dev));
    const size_t base_bytes = (size_t)nb * d * sizeof(float);
    const size_t out_elems = (size_t)n_rows * k;
    int rc;
    if ((rc = dc.base.ensure(base_bytes ? base_bytes : 4))) return rc;
    if ((rc = dc.assign.ensure(out_elems * sizeof(int)))) return rc;
    if ((rc = dc.dis.ensure(out_elems * sizeof(float)))) return rc;
    const size_t wbytes = workspace_bytes(n_rows, nb, d, k, self_join, dc.num_sms);
    if ((rc = dc.work.ensure(wbytes))) return rc;

    CU_TRY(cudaEventRecord(dc.ev[0], dc.stream));
    if (base_bytes) {
        if (ldb == (size_t)d)
            CU_TRY(cudaMemcpyAsync(dc.base.p, b, base_bytes, cudaMemcpyHostToDevice, dc.stream));
        else  // strided rows (e.g. an .fvecs image): the copy engine drops the padding
            CU_TRY(cudaMemcpy2DAsync(dc.base.p, (size_t)d * sizeof(float), b, ldb * sizeof(float),
                                     (size_t)d * sizeof(float), (size_t)nb,
                                     cudaMemcpyHostToDevice, dc.stream));
    }
    const float *d_b = static_cast<const float *>(dc.base.p);
    const float *d_v = d_b;
    int q_row0 = 0;
    if (!self_join) {
        // only the rows this call searches travel: row (q - q_begin) of the device copy
        const size_t qbytes = (size_t)n_rows * d * sizeof(float);
        if ((rc = dc.queries.ensure(qbytes))) return rc;
        if (ldv == (size_t)d)
            CU_TRY(cudaMemcpyAsync(dc.queries.p, v + (size_t)q_begin * d, qbytes,
                                   cudaMemcpyHostToDevice, dc.stream));
        else
            CU_TRY(cudaMemcpy2DAsync(dc.queries.p, (size_t)d * sizeof(float),
                                     v + (size_t)q_begin * ldv, ldv * sizeof(float),
                                     (size_t)d * sizeof(float), (size_t)n_rows,
                                     cudaMemcpyHostToDevice, dc.stream));
        d_v = static_cast<const float *>(dc.queries.p);
        q_row0 = q_begin;
    }
    CU_TRY(cudaEventRecord(dc.ev[1], dc.stream));
    rc = search_on_device(dc, c->algo, nb, d, k, d_b, d_v, q_row0, self_join, q_begin, q_end,
                          static_cast<int *>(dc.assign.p), static_cast<float *>(dc.dis.p),
                          dc.work.p, dc.work.cap, dc.stream, st, false);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(dc.ev[2], dc.stream));
    CU_TRY(cudaMemcpyAsync(assign, dc.assign.p, out_elems * sizeof(int), cudaMemcpyDeviceToHost,
                           dc.stream));
    CU_TRY(cudaMemcpyAsync(dis, dc.dis.p, out_elems * sizeof(float), cudaMemcpyDeviceToHost,
                           dc.stream));
    CU_TRY(cudaEventRecord(dc.ev[3], dc.stream));
    CU_TRY(cudaStreamSynchronize(dc.stream));
    if (st) {
        st->ms_h2d += ev_ms(dc.ev[0], dc.ev[1]);
        st->ms_kernels += ev_ms(dc.ev[1], dc.ev[2]);
        st->ms_d2h += ev_ms(dc.ev[2], dc.ev[3]);
        if (st->rows_exact > 0) st->ms_exact += ev_ms(dc.kev[4], dc.kev[5]);
    }
    return PQH_OK;
}

// One warp per block row: out[r] = data[ids[r]] (what block_push's memcpy does on the host,
// fast_nn_block.c:41-54).
__global__ void gather_block_rows_kernel(const float *__restrict__ data, const int *__restrict__ ids,
                                         int m, int d, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int n_warps = (int)((gridDim.x * (size_t)blockDim.x) >> 5);
    for (int r = warp; r < m; r += n_warps) {
        const float *src = data + (size_t)ids[r] * d;
        float *dst = out + (size_t)r * d;
        if ((d & 3) == 0) {
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            for (int c = lane; c < d / 4; c += 32) d4[c] = __ldg(s4 + c);
        } else {
            for (int c = lane; c < d; c += 32) dst[c] = __ldg(src + c);
        }
    }
}

// Self-join among rows `row_ids` of the data set resident on `dc` (host pointers in and out).
int dataset_search_block(Context *c, DeviceCtx &dc, int m, const int *row_ids, int k, int *assign,
                         float *dis, pqh_knn_stats_t *st) {
    if (m <= 0 || k <= 0) return PQH_OK;
    const int d = dc.ds_d;
    CU_TRY(cudaSetDevice(dc.dev));
    const size_t out_elems = (size_t)m * k;
    int rc;
    if ((rc = dc.queries.ensure((size_t)m * d * sizeof(float)))) return rc;
    if ((rc = dc.block_ids.ensure((size_t)m * sizeof(int)))) return rc;
    if ((rc = dc.assign.ensure(out_elems * sizeof(int)))) return rc;
    if ((rc = dc.dis.ensure(out_elems * sizeof(float)))) return rc;
    if ((rc = dc.work.ensure(workspace_bytes(m, m, d, k, 1, dc.num_sms)))) return rc;

    CU_TRY(cudaEventRecord(dc.ev[0], dc.stream));
    CU_TRY(cudaMemcpyAsync(dc.block_ids.p, row_ids, (size_t)m * sizeof(int),
                           cudaMemcpyHostToDevice, dc.stream));
    float *blk = static_cast<float *>(dc.queries.p);
    const int warps_per_cta = 8;
    int grid = (m + warps_per_cta - 1) / warps_per_cta;
    if (grid > 8 * dc.num_sms) grid = 8 * dc.num_sms;
    gather_block_rows_kernel<<<grid, warps_per_cta * 32, 0, dc.stream>>>(
        static_cast<const float *>(dc.dataset.p), static_cast<const int *>(dc.block_ids.p), m, d,
        blk);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(dc.ev[1], dc.stream));
    rc = search_on_device(dc, c->algo, m, d, k, blk, blk, 0, 1, 0, m,
                          static_cast<int *>(dc.assign.p), static_cast<float *>(dc.dis.p),
                          dc.work.p, dc.work.cap, dc.stream, st, false);
    if (rc) return rc;
    if (st) ++st->launches;
    CU_TRY(cudaEventRecord(dc.ev[2], dc.stream));
    CU_TRY(cudaMemcpyAsync(assign, dc.assign.p, out_elems * sizeof(int), cudaMemcpyDeviceToHost,
                           dc.stream));
    CU_TRY(cudaMemcpyAsync(dis, dc.dis.p, out_elems * sizeof(float), cudaMemcpyDeviceToHost,
                           dc.stream));
    CU_TRY(cudaEventRecord(dc.ev[3], dc.stream));
    CU_TRY(cudaStreamSynchronize(dc.stream));
    if (st) {
        st->ms_h2d += ev_ms(dc.ev[0], dc.ev[1]);
        st->ms_kernels += ev_ms(dc.ev[1], dc.ev[2]);
        st->ms_d2h += ev_ms(dc.ev[2], dc.ev[3]);
        if (st->rows_exact > 0) st->ms_exact += ev_ms(dc.kev[4], dc.kev[5]);
    }
    return PQH_OK;
}

int search_rows_impl(int nq_total, int nb, int d, int k, const float *b, size_t ldb, const float *v,
                     size_t ldv, int self_join, int q_begin, int q_end, int *assign, float *dis) {
    int rc = check_args(nq_total, nb, d, k, b, v, q_begin, q_end, assign, dis);
    if (rc) return rc;
    if (self_join && nq_total != nb)
        return fail(PQH_EINVAL, "self-join needs nq == nb (got %d, %d)", nq_total, nb);
    std::lock_guard<std::mutex> lock(g_mu);
    Context *c = nullptr;
    if ((rc = get_ctx(&c))) return rc;
    c->algo = g_algo;
    auto t0 = std::chrono::steady_clock::now();
    pqh_knn_stats_t st;
    memset(&st, 0, sizeof(st));
    const int n_rows = q_end - q_begin;
    const int G = (int)c->devs.size();
    if (G == 1 || n_rows < 2 * G) {
        rc = search_rows_host(c, c->devs[0], nq_total, nb, d, k, b, ldb, v, ldv, self_join, q_begin,
                              q_end, assign, dis, &st);
    } else {
        // query rows split evenly over the devices; each device
        // receives the base over its own PCIe link and returns its slice.
        std::vector<std::thread> th;
        std::vector<int> rcs(G, 0);
        std::vector<std::string> errs(G);
        std::vector<pqh_knn_stats_t> sts(G);
        const int per = (n_rows + G - 1) / G;
        for (int g = 0; g < G; ++g) {
            memset(&sts[g], 0, sizeof(pqh_knn_stats_t));
            th.emplace_back([&, g]() {
                int lo = q_begin + g * per, hi = lo + per;
                if (hi > q_end) hi = q_end;
                if (lo >= hi) return;
                rcs[g] = search_rows_host(c, c->devs[g], nq_total, nb, d, k, b, ldb, v, ldv,
                                          self_join, lo, hi, assign + (size_t)(lo - q_begin) * k,
                                          dis + (size_t)(lo - q_begin) * k, &sts[g]);
                if (rcs[g]) errs[g] = g_err;
            });
        }
        for (auto &t : th) t.join();
        for (int g = 0; g < G; ++g) {
            if (rcs[g] && !rc) {
                rc = rcs[g];
                g_err = errs[g];
            }
            st.rows += sts[g].rows;
            st.rows_tensor += sts[g].rows_tensor;
            st.rows_exact += sts[g].rows_exact;
            st.rows_second_pass += sts[g].rows_second_pass;
            st.launches += sts[g].launches;
            if (sts[g].ms_h2d > st.ms_h2d) st.ms_h2d = sts[g].ms_h2d;
            if (sts[g].ms_kernels > st.ms_kernels) st.ms_kernels = sts[g].ms_kernels;
            if (sts[g].ms_d2h > st.ms_d2h) st.ms_d2h = sts[g].ms_d2h;
            if (sts[g].ms_prep > st.ms_prep) st.ms_prep = sts[g].ms_prep;
            if (sts[g].ms_gemm > st.ms_gemm) st.ms_gemm = sts[g].ms_gemm;
            if (sts[g].ms_rerank > st.ms_rerank) st.ms_rerank = sts[g].ms_rerank;
            if (sts[g].ms_exact > st.ms_exact) st.ms_exact = sts[g].ms_exact;
            st.candidates += sts[g].candidates;
            if (sts[g].eps_ratio_max > st.eps_ratio_max) st.eps_ratio_max = sts[g].eps_ratio_max;
            st.inexact_operands |= sts[g].inexact_operands;
        }
    }
    st.ms_total =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->stats = st;
    return rc;
}

}  // namespace

extern "C" {

const char *pqh_last_error(void) { return g_err.c_str(); }

int pqh_knn_init(const int *devices, int n_devices) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_ctx) {
        destroy_ctx(g_ctx);
        g_ctx = nullptr;
    }
    return create_ctx(devices, n_devices, &g_ctx);
}

void pqh_knn_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    destroy_ctx(g_ctx);
    g_ctx = nullptr;
}

int pqh_knn_device_count(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    return g_ctx ? (int)g_ctx->devs.size() : 0;
}

int pqh_knn_set_algo(int algo) {
    if (algo != PQH_ALGO_AUTO && algo != PQH_ALGO_EXACT)
        return fail(PQH_EINVAL, "unknown algo %d", algo);
    g_algo = algo;
    return PQH_OK;
}

int pqh_knn_search_rows(int nq_total, int nb, int d, int k, const float *b, const float *v,
                        int self_join, int q_begin, int q_end, int *assign, float *dis) {
    return search_rows_impl(nq_total, nb, d, k, b, (size_t)d, v, (size_t)d, self_join, q_begin,
                            q_end, assign, dis);
}

int pqh_knn_search_rows_ld(int nq_total, int nb, int d, int k, const float *b, long long ldb,
                           const float *v, long long ldv, int self_join, int q_begin, int q_end,
                           int *assign, float *dis) {
    if (ldb < d || ldv < d) return fail(PQH_EINVAL, "leading dimension smaller than d");
    return search_rows_impl(nq_total, nb, d, k, b, (size_t)ldb, v, (size_t)ldv, self_join, q_begin,
                            q_end, assign, dis);
}

long long pqh_knn_workspace_bytes(int nq, int nb, int d, int k) {
    size_t a = workspace_bytes(nq, nb, d, k, 0, 148), b = workspace_bytes(nq, nb, d, k, 1, 148);
    return (long long)(a > b ? a : b);
}

int pqh_knn_search_device(int device, int nq_total, int nb, int d, int k, const float *d_b,
                          const float *d_v, int self_join, int q_begin, int q_end, int *d_assign,
                          float *d_dis, void *d_work, long long work_bytes, void *stream) {
    int rc = check_args(nq_total, nb, d, k, d_b, d_v, q_begin, q_end, d_assign, d_dis);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Context *c = nullptr;
    if ((rc = get_ctx(&c))) return rc;
    c->algo = g_algo;
    DeviceCtx *dc = nullptr;
    for (auto &x : c->devs)
        if (x.dev == device) dc = &x;
    if (!dc) return fail(PQH_EINVAL, "device %d is not part of the context", device);
    CU_TRY(cudaSetDevice(device));
    const int n_rows = q_end - q_begin;
    if (n_rows > 0 &&
        (long long)workspace_bytes(n_rows, nb, d, k, self_join, dc->num_sms) > work_bytes)
        return fail(PQH_EINVAL, "workspace too small: need %zu bytes",
                    workspace_bytes(n_rows, nb, d, k, self_join, dc->num_sms));
    pqh_knn_stats_t st;
    memset(&st, 0, sizeof(st));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : dc->stream;
    rc = search_on_device(*dc, c->algo, nb, d, k, d_b, d_v, 0, self_join, q_begin, q_end,
                          d_assign, d_dis, d_work, (size_t)work_bytes, s, &st, stream == nullptr);
    if (rc) return rc;
    if (!stream) CU_TRY(cudaStreamSynchronize(s));
    c->stats = st;
    return PQH_OK;
}

int pqh_knn_dataset_load(int n, int d, const float *rows, long long ld) {
    if (n < 0 || d <= 0 || ld < d || (n > 0 && !rows))
        return fail(PQH_EINVAL, "bad data set n=%d d=%d ld=%lld", n, d, ld);
    std::lock_guard<std::mutex> lock(g_mu);
    Context *c = nullptr;
    int rc;
    if ((rc = get_ctx(&c))) return rc;
    memset(&c->stats, 0, sizeof(c->stats));
    for (size_t g = 0; g < c->devs.size(); ++g) {
        DeviceCtx &dc = c->devs[g];
        std::lock_guard<std::mutex> dev_lock(g_dev_mu[g]);
        CU_TRY(cudaSetDevice(dc.dev));
        const size_t bytes = (size_t)n * d * sizeof(float);
        if ((rc = dc.dataset.ensure(bytes ? bytes : 4))) return rc;
        if (bytes)
            CU_TRY(cudaMemcpy2DAsync(dc.dataset.p, (size_t)d * sizeof(float), rows,
                                     (size_t)ld * sizeof(float), (size_t)d * sizeof(float),
                                     (size_t)n, cudaMemcpyHostToDevice, dc.stream));
        dc.ds_n = n;
        dc.ds_d = d;
    }
    for (auto &dc : c->devs) {
        CU_TRY(cudaSetDevice(dc.dev));
        CU_TRY(cudaStreamSynchronize(dc.stream));
    }
    return PQH_OK;
}

int pqh_knn_dataset_search_block(int slot, int m, const int *row_ids, int k, int *assign,
                                 float *dis) {
    Context *c = nullptr;
    int algo;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        c = g_ctx;
        algo = g_algo;
    }
    if (!c) return fail(PQH_EINVAL, "no data set loaded");
    if (slot < 0 || slot >= (int)c->devs.size())
        return fail(PQH_EINVAL, "device slot %d out of range (context has %zu)", slot,
                    c->devs.size());
    DeviceCtx &dc = c->devs[slot];
    if (m < 0 || k < 0 || (m > 0 && (!row_ids || (k > 0 && (!assign || !dis)))))
        return fail(PQH_EINVAL, "bad block m=%d k=%d", m, k);
    pqh_knn_stats_t st;
    memset(&st, 0, sizeof(st));
    auto t0 = std::chrono::steady_clock::now();
    int rc;
    {
        std::lock_guard<std::mutex> dev_lock(g_dev_mu[slot]);
        if (dc.ds_d <= 0) return fail(PQH_EINVAL, "no data set loaded");
        for (int i = 0; i < m; ++i)
            if (row_ids[i] < 0 || row_ids[i] >= dc.ds_n)
                return fail(PQH_EINVAL, "row id %d at position %d outside the data set (%d rows)",
                            row_ids[i], i, dc.ds_n);
        c->algo = algo;
        rc = dataset_search_block(c, dc, m, row_ids, k, assign, dis, &st);
    }
    const double ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> lock(g_mu);
    pqh_knn_stats_t &a = c->stats;
    a.rows += st.rows;
    a.rows_tensor += st.rows_tensor;
    a.rows_exact += st.rows_exact;
    a.rows_second_pass += st.rows_second_pass;
    a.launches += st.launches;
    a.ms_total += ms;
    a.ms_h2d += st.ms_h2d;
    a.ms_kernels += st.ms_kernels;
    a.ms_d2h += st.ms_d2h;
    a.ms_prep += st.ms_prep;
    a.ms_gemm += st.ms_gemm;
    a.ms_rerank += st.ms_rerank;
    a.ms_exact += st.ms_exact;
    a.candidates += st.candidates;
    if (st.eps_ratio_max > a.eps_ratio_max) a.eps_ratio_max = st.eps_ratio_max;
    a.inexact_operands |= st.inexact_operands;
    return rc;
}

void pqh_knn_dataset_release(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_ctx) return;
    for (size_t g = 0; g < g_ctx->devs.size(); ++g) {
        std::lock_guard<std::mutex> dev_lock(g_dev_mu[g]);
        DeviceCtx &dc = g_ctx->devs[g];
        if (cudaSetDevice(dc.dev) != cudaSuccess) continue;
        cudaStreamSynchronize(dc.stream);
        dc.dataset.release();
        dc.block_ids.release();
        dc.ds_n = dc.ds_d = 0;
    }
}

void *pqh_knn_host_alloc(long long bytes) {
    void *p = nullptr;
    if (bytes <= 0) return nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        Context *c = nullptr;
        if (get_ctx(&c)) return nullptr;
        cudaSetDevice(c->devs[0].dev);
    }
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        fail(PQH_ENOMEM, "cudaHostAlloc(%lld bytes) failed", bytes);
        return nullptr;
    }
    return p;
}

void pqh_knn_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int pqh_knn_last_stats(pqh_knn_stats_t *out) {
    if (!out) return fail(PQH_EINVAL, "null stats pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_ctx) return fail(PQH_EINVAL, "no context");
    *out = g_ctx->stats;
    return PQH_OK;
}

void knn_full_thread(int distance_type, int nq, int nb, int d, int k, const float *b,
                     const float *v, const float *b_weights, int *assign, float *dis,
                     int n_thread) {
    (void)n_thread;
    if (distance_type != 2 || b_weights != NULL) {
        fprintf(stderr,
                "libpqh_knn: knn_full_thread supports distance_type=2 (squared L2) with "
                "b_weights=NULL only, as compute_nn_fast.c:478-481 calls it\n");
        exit(1);
    }
    const int self_join = (b == v && nq == nb) ? 1 : 0;
    int rc = search_rows_impl(nq, nb, d, k, b, (size_t)d, v, (size_t)d, self_join, 0, nq, assign,
                              dis);
    if (rc) {
        fprintf(stderr, "libpqh_knn: knn_full_thread failed: %s\n", g_err.c_str());
        exit(1);
    }
}

}  // extern "C"
