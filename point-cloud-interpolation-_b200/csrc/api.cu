// api.cu -- ABI plumbing of libb200pc.so: error reporting, device queries, the FP32 peak
// micro-benchmark and the host-buffer convenience wrappers.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace b200pc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return B200PC_ECUDA;
}

// SM count of the current device; 148 (B200) when no device is visible so that the planning
// helpers (workspace sizing) still work on a build box.  Compute entry points fail on their
// first CUDA call in that case: there is no CPU fallback.
int sm_count() {
    static thread_local int cached_dev = -2, cached = 148;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        else { cudaGetLastError(); cached = 148; }
        cached_dev = dev;
    }
    return cached;
}

static int env_int(const char *name, int unset) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : unset;
}

static Tuning g_tuning;
static bool g_tuning_loaded = false;

static void load_tuning() {
    Tuning t;
    t.force_q = env_int("B200PC_FORCE_Q", 0);
    t.force_warps = env_int("B200PC_FORCE_WARPS", 0);
    t.force_split = env_int("B200PC_FORCE_SPLIT", 0);
    t.natural_order = env_int("B200PC_NATURAL_ORDER", -1);
    t.nodrain = getenv("B200PC_DEBUG_NODRAIN") != nullptr;
    t.filter = env_int("B200PC_FILTER", -1);
    t.small_path = env_int("B200PC_SMALL_PATH", -1);
    t.gather_rows = env_int("B200PC_GATHER_ROWS", -1);
    t.gather_flat = env_int("B200PC_GATHER_FLAT", -1);
    t.interp_rows = env_int("B200PC_INTERP_ROWS", -1);
    t.interp_flat = env_int("B200PC_INTERP_FLAT", -1);
    t.bulk = env_int("B200PC_BULK", -1);
    t.fps_cluster = env_int("B200PC_FPS_CLUSTER", 0);
    t.fps_flat = env_int("B200PC_FPS_FLAT", -1);
    t.interleave = env_int("B200PC_INTERLEAVE", 0);
    t.grid = env_int("B200PC_GRID", 1);
    t.seed = env_int("B200PC_SEED", 0);
    t.debug_plan = env_int("B200PC_DEBUG_PLAN", 0);
    t.bounds_trip = env_int("B200PC_BOUNDS_TRIP", 0);
    g_tuning = t;
    g_tuning_loaded = true;
}

const Tuning &tuning() {
    if (!g_tuning_loaded) load_tuning();
    return g_tuning;
}

// 8 independent packed-FMA chains per thread, fully register resident
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float *sink) {
    // self-test of the bounds build: a negative iteration count (B200PC_BOUNDS_TRIP=1) must trap; compiled out of the shipped library
    B200PC_DEV_ASSERT(iters >= 0);
    f32x2 a[8];
    const f32x2 m = splat2(1.0000001f), c = splat2(seed);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = splat2(seed + (float)(threadIdx.x + i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], m, c);
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; unpack2(a[i], lo, hi); acc += lo + hi; }
    if (acc == 123.456f) sink[0] = acc;  // never true; keeps the chains alive
}

}  // namespace b200pc

using namespace b200pc;

extern "C" const char *b200pc_last_error(void) { return g_err; }
extern "C" int b200pc_version(void) { return 200; }
extern "C" void b200pc_tuning_reload(void) { load_tuning(); }

extern "C" int b200pc_device_sm_count(void) {
    int dev = -1, n = 0;
    B200PC_CUDA(cudaGetDevice(&dev));
    B200PC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
};
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};
}  // namespace

extern "C" int b200pc_fma_peak(int iters, double *tflops, double *ms_out, b200pc_stream_t stream) {
    B200PC_REQUIRE(iters > 0 && tflops, "fma_peak: bad arguments");
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    const int blocks = sms * 8, threads = 256;
    DevBuf sink;                        // RAII: nothing leaks on an early error return
    EventPair ev;
    B200PC_CUDA(sink.alloc(sizeof(float)));
    B200PC_CUDA(cudaEventCreate(&ev.a));
    B200PC_CUDA(cudaEventCreate(&ev.b));
    fma_peak_kernel<<<blocks, threads, 0, st>>>(tuning().bounds_trip ? -1 : iters / 8 + 1, 0.5f, static_cast<float *>(sink.p));  // warm-up
    B200PC_LAUNCH_CHECK();
    B200PC_CUDA(cudaEventRecord(ev.a, st));
    fma_peak_kernel<<<blocks, threads, 0, st>>>(iters, 0.5f, static_cast<float *>(sink.p));
    B200PC_LAUNCH_CHECK();
    B200PC_CUDA(cudaEventRecord(ev.b, st));
    B200PC_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0.f;
    B200PC_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    const double flop = (double)blocks * threads * (double)iters * 8.0 /*chains*/ * 2.0 /*lanes*/ * 2.0 /*mul+add*/;
    *tflops = flop / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return B200PC_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer wrappers: for callers without a device allocator of their own
// ------------------------------------------------------------------------------------------------
extern "C" int b200pc_knn_host(const float *ref, const float *qry, int B, int N, int S, int k, int form, int64_t *idx,
                               float *dist) {
    B200PC_REQUIRE(ref && qry && idx, "knn_host: null pointer");
    B200PC_REQUIRE(B >= 1 && N >= 1 && S >= 1 && k >= 1 && k <= N, "knn_host: bad sizes");
    DevBuf dref, dq, didx, ddist, ws;
    const size_t nr = (size_t)B * N * 3 * 4, nq = (size_t)B * S * 3 * 4, no = (size_t)B * S * k;
    const size_t wsb = b200pc_search_workspace_bytes(B, N, S, k);
    B200PC_CUDA(dref.alloc(nr)); B200PC_CUDA(dq.alloc(nq)); B200PC_CUDA(didx.alloc(no * 8));
    B200PC_CUDA(ddist.alloc(no * 4)); B200PC_CUDA(ws.alloc(wsb));
    B200PC_CUDA(cudaMemcpy(dref.p, ref, nr, cudaMemcpyHostToDevice));
    B200PC_CUDA(cudaMemcpy(dq.p, qry, nq, cudaMemcpyHostToDevice));
    int rc = b200pc_knn((const float *)dref.p, (const float *)dq.p, B, N, S, k, form, (int64_t *)didx.p,
                        dist ? (float *)ddist.p : nullptr, ws.p, wsb, nullptr);
    if (rc != B200PC_OK) return rc;
    B200PC_CUDA(cudaMemcpy(idx, didx.p, no * 8, cudaMemcpyDeviceToHost));
    if (dist) B200PC_CUDA(cudaMemcpy(dist, ddist.p, no * 4, cudaMemcpyDeviceToHost));
    return B200PC_OK;
}

// Asynchronous host-buffer kNN: everything is enqueued on `stream` and nothing is allocated, freed or synchronised, so a
// plain C caller can keep two calls in flight on two streams (two arenas) and overlap the read-back of one with the upload
// and search of the next, like b200pc.hostio.KnnHostPipeline does above torch.  Host buffers should be pinned.
static size_t knn_async_layout(int B, int N, int S, int k, size_t off[5]) {
    const size_t sz[5] = {(size_t)B * N * 12, (size_t)B * S * 12, (size_t)B * S * k * 4, (size_t)B * S * k * 4,
                          b200pc_search_workspace_bytes(B, N, S, k)};
    size_t o = 0;
    for (int i = 0; i < 5; ++i) { off[i] = o; o += align_up(sz[i], 256); }
    return o;
}

extern "C" size_t b200pc_knn_async_host_workspace_bytes(int B, int N, int S, int k) {
    if (B <= 0 || N <= 0 || S <= 0 || k <= 0) return 256;
    size_t off[5];
    return knn_async_layout(B, N, S, k, off);
}

extern "C" int b200pc_knn_async_host(const float *ref, const float *qry, int B, int N, int S, int k, int form, int32_t *idx,
                                     float *dist, void *arena, size_t arena_bytes, b200pc_stream_t stream) {
    B200PC_REQUIRE(ref && qry && idx && arena, "knn_async_host: null pointer");
    B200PC_REQUIRE(B >= 1 && N >= 1 && S >= 1 && k >= 1 && k <= N, "knn_async_host: bad sizes");
    size_t off[5];
    const size_t need = knn_async_layout(B, N, S, k, off);
    if (arena_bytes < need) {
        set_error("knn_async_host: arena too small (%zu < %zu bytes)", arena_bytes, need);
        return B200PC_EWORKSPACE;
    }
    B200PC_REQUIRE((reinterpret_cast<uintptr_t>(arena) & 255) == 0, "knn_async_host: the arena must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    char *a = static_cast<char *>(arena);
    float *dref = reinterpret_cast<float *>(a + off[0]), *dq = reinterpret_cast<float *>(a + off[1]);
    int32_t *didx = reinterpret_cast<int32_t *>(a + off[2]);
    float *ddist = dist ? reinterpret_cast<float *>(a + off[3]) : nullptr;
    const size_t no = (size_t)B * S * k;
    B200PC_CUDA(cudaMemcpyAsync(dref, ref, (size_t)B * N * 12, cudaMemcpyHostToDevice, st));
    B200PC_CUDA(cudaMemcpyAsync(dq, qry, (size_t)B * S * 12, cudaMemcpyHostToDevice, st));
    const int rc = b200pc_knn_i32(dref, dq, B, N, S, k, form, didx, ddist, a + off[4], need - off[4], stream);
    if (rc != B200PC_OK) return rc;
    B200PC_CUDA(cudaMemcpyAsync(idx, didx, no * 4, cudaMemcpyDeviceToHost, st));
    if (dist) B200PC_CUDA(cudaMemcpyAsync(dist, ddist, no * 4, cudaMemcpyDeviceToHost, st));
    return B200PC_OK;
}

extern "C" int b200pc_ball_query_host(const float *xyz, const float *new_xyz, int B, int N, int S, float r2, int nsample,
                                      int64_t *idx) {
    B200PC_REQUIRE(xyz && new_xyz && idx, "ball_query_host: null pointer");
    B200PC_REQUIRE(B >= 1 && N >= 1 && S >= 1 && nsample >= 1, "ball_query_host: bad sizes");
    DevBuf dref, dq, didx, ws;
    const size_t nr = (size_t)B * N * 3 * 4, nq = (size_t)B * S * 3 * 4, no = (size_t)B * S * nsample;
    const size_t wsb = b200pc_search_workspace_bytes(B, N, S, nsample);
    B200PC_CUDA(dref.alloc(nr)); B200PC_CUDA(dq.alloc(nq)); B200PC_CUDA(didx.alloc(no * 8)); B200PC_CUDA(ws.alloc(wsb));
    B200PC_CUDA(cudaMemcpy(dref.p, xyz, nr, cudaMemcpyHostToDevice));
    B200PC_CUDA(cudaMemcpy(dq.p, new_xyz, nq, cudaMemcpyHostToDevice));
    int rc = b200pc_ball_query((const float *)dref.p, (const float *)dq.p, B, N, S, r2, nsample, (int64_t *)didx.p, ws.p,
                               wsb, nullptr);
    if (rc != B200PC_OK) return rc;
    B200PC_CUDA(cudaMemcpy(idx, didx.p, no * 8, cudaMemcpyDeviceToHost));
    return B200PC_OK;
}

extern "C" int b200pc_fps_host(const float *xyz, int B, int N, int npoint, const int64_t *start, int64_t *idx) {
    B200PC_REQUIRE(xyz && start && idx, "fps_host: null pointer");
    B200PC_REQUIRE(B >= 1 && N >= 1 && npoint >= 1, "fps_host: bad sizes");
    DevBuf dx, ds, di;
    const size_t nx = (size_t)B * N * 3 * 4;
    B200PC_CUDA(dx.alloc(nx)); B200PC_CUDA(ds.alloc((size_t)B * 8)); B200PC_CUDA(di.alloc((size_t)B * npoint * 8));
    B200PC_CUDA(cudaMemcpy(dx.p, xyz, nx, cudaMemcpyHostToDevice));
    B200PC_CUDA(cudaMemcpy(ds.p, start, (size_t)B * 8, cudaMemcpyHostToDevice));
    int rc = b200pc_fps((const float *)dx.p, B, N, npoint, (const int64_t *)ds.p, (int64_t *)di.p, nullptr, 0, nullptr);
    if (rc != B200PC_OK) return rc;
    B200PC_CUDA(cudaMemcpy(idx, di.p, (size_t)B * npoint * 8, cudaMemcpyDeviceToHost));
    return B200PC_OK;
}
