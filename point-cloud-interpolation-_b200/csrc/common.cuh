// common.cuh -- shared plumbing for libb200pc.so (sm_100a only).
//   * error handling behind the C ABI (thread-local message, no exceptions cross the ABI)
//   * PTX wrappers: packed fp32x2 math (FFMA2/FMUL2/FADD2), 3-input min (FMNMX3),
//     mbarrier, 1-D bulk TMA copy (cp.async.bulk -> UBLKCP), vector reductions
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200pc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200pc is written for sm_100a (B200) only"
#endif

namespace b200pc {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define B200PC_CUDA(call)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) return ::b200pc::cuda_fail(_e, #call, __FILE__, __LINE__);     \
    } while (0)

#define B200PC_REQUIRE(cond, ...)                                                             \
    do {                                                                                      \
        if (!(cond)) { ::b200pc::set_error(__VA_ARGS__); return B200PC_EINVAL; }              \
    } while (0)

#define B200PC_LAUNCH_CHECK() B200PC_CUDA(cudaGetLastError())

// Device-side index checks of the bounds build (make bounds, -DB200PC_BOUNDS): a violated condition prints where and traps
// the launch, so the caller sees a CUDA error.  Compiled out of the shipped library.
#ifdef B200PC_BOUNDS
#include <stdio.h>
#define B200PC_DEV_ASSERT(cond)                                                                                          \
    do {                                                                                                                 \
        if (!(cond)) {                                                                                                   \
            printf("b200pc device assert failed: %s (%s:%d) block (%d,%d,%d) thread %d\n", #cond, __FILE__, __LINE__,     \
                   (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)threadIdx.x);                                  \
            __trap();                                                                                                    \
        }                                                                                                                \
    } while (0)
#else
#define B200PC_DEV_ASSERT(cond) ((void)0)
#endif

static inline cudaStream_t as_stream(b200pc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();

// Tuning / debugging knobs (B200PC_* environment variables; NOT part of the ABI).  They are read ONCE, at the first
// launch, and cached -- a launch never calls getenv.  b200pc_tuning_reload() re-reads them (tests and A/B probes that
// flip a knob between calls).  -1 = not set (the launcher's own default applies).
struct Tuning {
    int force_q, force_warps, force_split;   // search planner overrides (0 = none)
    int natural_order;                       // top-k refs kept in index order (1) / dealt out strided (0)
    int nodrain;                             // measurement only: the search starts with tau = -inf
    int filter;                              // 0: broadcast filter on every tile (the v6 hot loop)
    int small_path;                          // 0 / 1: forbid / force the warp-per-query path
    int gather_rows, gather_flat, interp_rows, interp_flat;
    int bulk;                                // 0: register-path row movers (gather.cu / group.cu) instead of rowmove.cu
    int fps_cluster;
    int fps_flat;                            // 0: two-level arg-max (block, then cluster records); 1: flat exchange of warp keys; -1: by cluster size
    int interleave;                          // 1: cell-ordered queries are dealt out to the CTAs warp by warp instead of in contiguous blocks
    int debug_plan;                          // print the plan of every streaming search to stderr
    int bounds_trip;                         // self-test of the bounds build: b200pc_fma_peak launches with an index its check rejects
    int seed;                                // 0 (default): starting thresholds = corner bound of the query's cell box; n > 0: k-th distance inside boxes up to level n-1
    int grid;                                // 0: top-k searches start from tau = +inf; 1: default (warm start when worth it); 2: always
};
const Tuning &tuning();

// ---------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic.  A "pair" is a 64-bit register holding two IEEE fp32 lanes; each op
// rounds both lanes to nearest-even independently, exactly like the scalar instruction would.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 splat2(float v) { return pack2(v, v); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D).  Addresses are 32-bit shared-window addresses.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"   // suspend-time hint: park the warp instead of spinning
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(20000u)
        : "memory");
}
// single non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- distributed shared memory (thread-block clusters) ----
// shared::cta address -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16-byte asynchronous store into a peer CTA's shared memory; the bytes are counted on the peer's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst_cluster),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar_cluster)
                 : "memory");
}
// 8-byte asynchronous store into a peer CTA's shared memory (same completion mechanism)
__device__ __forceinline__ void st_async_b64(uint32_t dst_cluster, unsigned long long v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst_cluster), "l"(v),
                 "r"(bar_cluster)
                 : "memory");
}
// wait with cluster-scope acquire: data written by peers with st.async is visible afterwards
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "DONE_C:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// streaming (no L1 allocate) 128-bit global accesses for the HBM-bound kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, const float4 &v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace b200pc
