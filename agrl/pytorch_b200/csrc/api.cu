// api.cu -- ABI bookkeeping and the host-buffer entry points of libagrl_b200.
#include "common.cuh"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace agrl {

static thread_local char     tl_error[512] = "";
static thread_local uint64_t tl_launches = 0;

void set_cuda_error(cudaError_t e, const char *what, const char *file, int line) {
    snprintf(tl_error, sizeof(tl_error), "%s (%s) at %s:%d: %s", cudaGetErrorName(e), what, file, line,
             cudaGetErrorString(e));
    (void)cudaGetLastError();          // clear the sticky-less error so later calls start clean
}
// Optional per-thread kernel timeline: one CUDA event after every launch, on the launching stream.
struct Profiler {
    static constexpr int kMax = 4096;
    bool on = false;
    int n = 0;
    cudaEvent_t ev[kMax + 1];
    cudaEvent_t begin[kMax + 1];       // optional explicit start of launch i (before_launch), else ev[i - 1]
    bool has_begin[kMax + 1];
    const char *name[kMax + 1];
};
static thread_local Profiler tl_prof;

bool profiling_active() { return tl_prof.on; }

void before_launch(cudaStream_t st) {
    Profiler &p = tl_prof;
    if (!p.on || p.n >= Profiler::kMax || p.has_begin[p.n + 1]) return;
    if (cudaEventCreate(&p.begin[p.n + 1]) != cudaSuccess) { (void)cudaGetLastError(); return; }
    if (cudaEventRecord(p.begin[p.n + 1], st) != cudaSuccess) {
        (void)cudaGetLastError();
        cudaEventDestroy(p.begin[p.n + 1]);
        return;
    }
    p.has_begin[p.n + 1] = true;
}

void after_launch(cudaStream_t st, const char *name) {
    ++tl_launches;
    Profiler &p = tl_prof;
    if (p.on && p.n < Profiler::kMax) {
        ++p.n;
        if (cudaEventCreate(&p.ev[p.n]) != cudaSuccess || cudaEventRecord(p.ev[p.n], st) != cudaSuccess) {
            (void)cudaGetLastError();
            if (p.has_begin[p.n]) { cudaEventDestroy(p.begin[p.n]); p.has_begin[p.n] = false; }
            --p.n;
            return;
        }
        p.name[p.n] = name;
        if (p.n < Profiler::kMax) p.has_begin[p.n + 1] = false;
    }
}

// per-thread stream + stream-ordered scratch for the *_host entry points
struct HostCtx {
    cudaStream_t stream = nullptr;
    int device = -1;
    int ensure() {
        int dev = -1;
        AGRL_CUDA_TRY(cudaGetDevice(&dev));
        if (stream && dev == device) return AGRL_OK;
        if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
        AGRL_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        device = dev;
        cudaMemPool_t pool;
        AGRL_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = ~0ull;          // keep freed blocks cached: repeated calls reuse them
        AGRL_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        return AGRL_OK;
    }
};
static thread_local HostCtx tl_ctx;

// RAII list of stream-ordered allocations
struct Scratch {
    cudaStream_t st;
    void *ptrs[16];
    int n = 0;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st); }
    int alloc(void **p, size_t bytes) {
        if (n >= 16) return AGRL_E_INVALID;
        AGRL_CUDA_TRY(cudaMallocAsync(p, bytes ? bytes : 16, st));
        ptrs[n++] = *p;
        return AGRL_OK;
    }
};

#define AGRL_TRY(expr) do { int _rc = (expr); if (_rc != AGRL_OK) return _rc; } while (0)

static int status_to_code(uint32_t st) {
    if (st & AGRL_ST_LABEL_RANGE) return AGRL_E_LABEL_RANGE;
    if (st & AGRL_ST_NO_VALID_QUERY) return AGRL_E_NO_VALID_QUERY;
    if (st & AGRL_ST_ZERO_DIVISION) return AGRL_E_ZERO_DIVISION;
    return AGRL_OK;
}

}  // namespace agrl

using namespace agrl;

extern "C" int agrl_abi_version(void) { return AGRL_B200_ABI_VERSION; }

extern "C" const char *agrl_status_string(int code) {
    switch (code) {
        case AGRL_OK: return "ok";
        case AGRL_E_INVALID: return "invalid argument";
        case AGRL_E_NO_DEVICE: return "no CUDA device of compute capability 10.x (this library has no CPU fallback)";
        case AGRL_E_CUDA: return "CUDA runtime error (see agrl_last_cuda_error)";
        case AGRL_E_WORKSPACE: return "workspace missing or too small";
        case AGRL_E_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case AGRL_E_NO_VALID_QUERY: return "Error: all query identities do not appear in gallery";
        case AGRL_E_ZERO_DIVISION: return "division by zero: a query has no cross-camera match (MARS metric)";
        case AGRL_E_LABEL_RANGE: return "pid / camid outside the int32 range";
        default: return "unknown status";
    }
}

extern "C" const char *agrl_last_cuda_error(void) { return tl_error; }

extern "C" uint64_t agrl_launch_count(void) { return tl_launches; }

extern "C" int agrl_device_ok(void) {
    static std::atomic<int> cache[64];          // 0 unknown, 1 ok, 2 not ok
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) { (void)cudaGetLastError(); return AGRL_E_NO_DEVICE; }
    if (dev < 64) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c) return c == 1 ? AGRL_OK : AGRL_E_NO_DEVICE;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return AGRL_E_NO_DEVICE;
    }
    const bool ok = (major == 10);
    if (dev < 64) cache[dev].store(ok ? 1 : 2, std::memory_order_relaxed);
    return ok ? AGRL_OK : AGRL_E_NO_DEVICE;
}

// ------------------------------------------------------------------------------------------------
// ranking, host buffers
// ------------------------------------------------------------------------------------------------
static int upload_rank_inputs(Scratch &sc, cudaStream_t st, const float *dist, const int64_t *qp,
                              const int64_t *gp, const int64_t *qc, const int64_t *gc, int64_t nq, int64_t ng,
                              float **d_dist, int64_t **d_lab) {
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(d_dist), sizeof(float) * nq * ng));
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(d_lab), sizeof(int64_t) * 2 * (nq + ng)));
    AGRL_CUDA_TRY(cudaMemcpyAsync(*d_dist, dist, sizeof(float) * nq * ng, cudaMemcpyHostToDevice, st));
    int64_t *l = *d_lab;
    AGRL_CUDA_TRY(cudaMemcpyAsync(l, qp, sizeof(int64_t) * nq, cudaMemcpyHostToDevice, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(l + nq, qc, sizeof(int64_t) * nq, cudaMemcpyHostToDevice, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(l + 2 * nq, gp, sizeof(int64_t) * ng, cudaMemcpyHostToDevice, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(l + 2 * nq + ng, gc, sizeof(int64_t) * ng, cudaMemcpyHostToDevice, st));
    return AGRL_OK;
}

extern "C" int agrl_rank_market1501_host(const float *distmat, const int64_t *q_pids, const int64_t *g_pids,
                                         const int64_t *q_camids, const int64_t *g_camids,
                                         int64_t num_q, int64_t num_g, int64_t max_rank,
                                         float *cmc, float *map, float *all_ap,
                                         int64_t *rank_len_out, int64_t *num_valid_out) {
    if (!distmat || !q_pids || !g_pids || !q_camids || !g_camids || !cmc || !map) return AGRL_E_INVALID;
    if (num_q < 0 || num_g < 1 || max_rank < 1) return AGRL_E_INVALID;
    AGRL_TRY(agrl_device_ok());
    AGRL_TRY(tl_ctx.ensure());
    cudaStream_t st = tl_ctx.stream;
    const int64_t rank_len = max_rank < num_g ? max_rank : num_g;
    if (rank_len_out) *rank_len_out = rank_len;
    if (num_q == 0) return AGRL_E_NO_VALID_QUERY;
    Scratch sc(st);
    float *d_dist; int64_t *d_lab; void *ws; char *d_out;
    AGRL_TRY(upload_rank_inputs(sc, st, distmat, q_pids, g_pids, q_camids, g_camids, num_q, num_g, &d_dist, &d_lab));
    const size_t wsb = agrl_rank_workspace_bytes(num_q, num_g, max_rank);
    AGRL_TRY(sc.alloc(&ws, wsb));
    // outputs packed: [cmc rank_len f32][map f32][status u32][num_valid i64][all_ap nq f32]
    const size_t off_map = sizeof(float) * rank_len, off_st = off_map + 4;
    const size_t off_nv = align_up(off_st + 4, 8), off_ap = off_nv + 8;
    const size_t out_bytes = off_ap + sizeof(float) * num_q;
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(&d_out), out_bytes));
    AGRL_TRY(agrl_rank_market1501_dev(d_dist, num_g, d_lab, d_lab + 2 * num_q, d_lab + num_q,
                                      d_lab + 2 * num_q + num_g, num_q, num_g, max_rank,
                                      reinterpret_cast<float *>(d_out), reinterpret_cast<float *>(d_out + off_map),
                                      reinterpret_cast<float *>(d_out + off_ap),
                                      reinterpret_cast<int64_t *>(d_out + off_nv),
                                      reinterpret_cast<uint32_t *>(d_out + off_st), ws, wsb, st));
    // [map f32][status u32][pad to 8][num_valid i64]: 16 bytes when off_map is a multiple of 8, 20 when rank_len is odd
    char small[24];
    if (off_ap - off_map > sizeof(small)) return AGRL_E_INVALID;
    AGRL_CUDA_TRY(cudaMemcpyAsync(cmc, d_out, sizeof(float) * rank_len, cudaMemcpyDeviceToHost, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(small, d_out + off_map, off_ap - off_map, cudaMemcpyDeviceToHost, st));
    if (all_ap) AGRL_CUDA_TRY(cudaMemcpyAsync(all_ap, d_out + off_ap, sizeof(float) * num_q, cudaMemcpyDeviceToHost, st));
    AGRL_CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t status; int64_t nv;
    memcpy(map, small, 4);
    memcpy(&status, small + (off_st - off_map), 4);
    memcpy(&nv, small + (off_nv - off_map), 8);
    if (num_valid_out) *num_valid_out = nv;
    return status_to_code(status);
}

extern "C" int agrl_rank_mars_host(const float *distmat, const int64_t *q_pids, const int64_t *g_pids,
                                   const int64_t *q_camids, const int64_t *g_camids,
                                   int64_t num_q, int64_t num_g, int64_t max_rank,
                                   double *cmc, double *map, double *all_ap) {
    if (!distmat || !q_pids || !g_pids || !q_camids || !g_camids || !cmc || !map) return AGRL_E_INVALID;
    if (num_q < 1 || num_g < 1 || max_rank < 1) return AGRL_E_INVALID;
    if (max_rank > num_g || max_rank > 8192) return AGRL_E_UNSUPPORTED;
    AGRL_TRY(agrl_device_ok());
    AGRL_TRY(tl_ctx.ensure());
    cudaStream_t st = tl_ctx.stream;
    Scratch sc(st);
    float *d_dist; int64_t *d_lab; void *ws; char *d_out;
    AGRL_TRY(upload_rank_inputs(sc, st, distmat, q_pids, g_pids, q_camids, g_camids, num_q, num_g, &d_dist, &d_lab));
    const size_t wsb = agrl_rank_workspace_bytes(num_q, num_g, max_rank);
    AGRL_TRY(sc.alloc(&ws, wsb));
    // outputs packed: [cmc max_rank f64][map f64][status u32 + pad][all_ap nq f64]
    const size_t off_map = sizeof(double) * max_rank, off_st = off_map + 8, off_ap = off_st + 8;
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(&d_out), off_ap + sizeof(double) * num_q));
    AGRL_TRY(agrl_rank_mars_dev(d_dist, num_g, d_lab, d_lab + 2 * num_q, d_lab + num_q, d_lab + 2 * num_q + num_g,
                                num_q, num_g, max_rank, reinterpret_cast<double *>(d_out),
                                reinterpret_cast<double *>(d_out + off_map), reinterpret_cast<double *>(d_out + off_ap),
                                reinterpret_cast<uint32_t *>(d_out + off_st), ws, wsb, st));
    char small[16];
    AGRL_CUDA_TRY(cudaMemcpyAsync(cmc, d_out, sizeof(double) * max_rank, cudaMemcpyDeviceToHost, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(small, d_out + off_map, 16, cudaMemcpyDeviceToHost, st));
    if (all_ap) AGRL_CUDA_TRY(cudaMemcpyAsync(all_ap, d_out + off_ap, sizeof(double) * num_q, cudaMemcpyDeviceToHost, st));
    AGRL_CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t status;
    memcpy(map, small, 8);
    memcpy(&status, small + 8, 4);
    return status_to_code(status);
}

// ------------------------------------------------------------------------------------------------
// distance matrix, host buffers
// ------------------------------------------------------------------------------------------------
extern "C" int agrl_distance_host(const float *q_host, const float *g_host, float *out_host,
                                  int64_t num_q, int64_t num_g, int64_t dim, int metric, int split) {
    if (!q_host || !g_host || !out_host) return AGRL_E_INVALID;
    if (num_q < 0 || num_g < 0 || dim < 1) return AGRL_E_INVALID;
    AGRL_TRY(agrl_device_ok());
    if (num_q == 0 || num_g == 0) return AGRL_OK;
    AGRL_TRY(tl_ctx.ensure());
    cudaStream_t st = tl_ctx.stream;
    Scratch sc(st);
    float *d_q, *d_g, *d_out; void *ws;
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(&d_q), sizeof(float) * num_q * dim));
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(&d_g), sizeof(float) * num_g * dim));
    AGRL_TRY(sc.alloc(reinterpret_cast<void **>(&d_out), sizeof(float) * num_q * num_g));
    const size_t wsb = agrl_distance_workspace_bytes(num_q, num_g, dim, split);
    if (wsb == 0) return AGRL_E_INVALID;
    AGRL_TRY(sc.alloc(&ws, wsb));
    AGRL_CUDA_TRY(cudaMemcpyAsync(d_q, q_host, sizeof(float) * num_q * dim, cudaMemcpyHostToDevice, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(d_g, g_host, sizeof(float) * num_g * dim, cudaMemcpyHostToDevice, st));
    AGRL_TRY(agrl_distance_dev(d_q, dim, d_g, dim, d_out, num_g, num_q, num_g, dim, metric, split, ws, wsb, st));
    AGRL_CUDA_TRY(cudaMemcpyAsync(out_host, d_out, sizeof(float) * num_q * num_g, cudaMemcpyDeviceToHost, st));
    AGRL_CUDA_TRY(cudaStreamSynchronize(st));
    return AGRL_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel timeline (bench.py's roofline numbers come from here: CUDA events on the launching stream)
// ------------------------------------------------------------------------------------------------
extern "C" int agrl_profile_begin(void *stream) {
    Profiler &p = tl_prof;
    if (p.on) return AGRL_E_INVALID;
    p.n = 0;
    p.has_begin[0] = p.has_begin[1] = false;
    AGRL_CUDA_TRY(cudaEventCreate(&p.ev[0]));
    AGRL_CUDA_TRY(cudaEventRecord(p.ev[0], static_cast<cudaStream_t>(stream)));
    p.name[0] = "begin";
    p.on = true;
    return AGRL_OK;
}

extern "C" int agrl_profile_end(char *text, size_t cap) {
    Profiler &p = tl_prof;
    if (!p.on) return AGRL_E_INVALID;
    p.on = false;
    size_t off = 0;
    if (text && cap) text[0] = 0;
    int rc = AGRL_OK;
    for (int i = 1; i <= p.n && rc == AGRL_OK; ++i)          // launches of one call may sit on several streams
        if (cudaEventSynchronize(p.ev[i]) != cudaSuccess) rc = AGRL_E_CUDA;
    for (int i = 1; i <= p.n && rc == AGRL_OK; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.has_begin[i] ? p.begin[i] : p.ev[i - 1], p.ev[i]) != cudaSuccess) { rc = AGRL_E_CUDA; break; }
        if (text && off + 64 < cap) off += snprintf(text + off, cap - off, "%s:%.6f;", p.name[i], ms);
    }
    for (int i = 0; i <= p.n; ++i) {
        cudaEventDestroy(p.ev[i]);
        if (p.has_begin[i]) { cudaEventDestroy(p.begin[i]); p.has_begin[i] = false; }
    }
    if (p.n < Profiler::kMax && p.has_begin[p.n + 1]) { cudaEventDestroy(p.begin[p.n + 1]); p.has_begin[p.n + 1] = false; }
    p.n = 0;
    if (rc != AGRL_OK) (void)cudaGetLastError();
    return rc;
}
