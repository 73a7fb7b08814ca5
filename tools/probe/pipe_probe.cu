// pipe_probe.cu -- throughput of the packed fp32x2 instruction mix used by the search hot loop.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run on B200
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b){ f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b){ f32x2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
__device__ __forceinline__ float min3(float a, float b, float c){ float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}

template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, float seed, float *sink) {
    f32x2 a[8];
    const f32x2 m = pk(1.0000001f, 1.0000001f), c = pk(seed, seed), one = pk(1.f, 1.f);
    float mn = 1e30f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = pk(seed + threadIdx.x + i, seed + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = fma2(a[i], m, c); }                       // FFMA2 only
            if (MODE == 1) { a[i] = add2(a[i], c); }                          // FADD2 only
            if (MODE == 2) { a[i] = mul2(a[i], m); }                          // FMUL2 only
            if (MODE == 3) { f32x2 t = mul2(a[i], m); t = fma2(a[i], c, t); t = fma2(a[i], m, t); t = add2(t, c); a[i] = add2(t, m); }   // hot-loop mix
            if (MODE == 4) { f32x2 t = fma2(a[i], m, c); t = fma2(a[i], c, t); t = fma2(a[i], m, t); t = fma2(t, one, c); a[i] = fma2(t, one, m); } // all-FFMA2 equivalent
            if (MODE == 5) { f32x2 t = mul2(a[i], m); t = fma2(a[i], c, t); t = fma2(a[i], m, t); t = add2(t, c); a[i] = add2(t, m);
                             float lo = __uint_as_float((unsigned)a[i]), hi = __uint_as_float((unsigned)(a[i] >> 32)); mn = min3(mn, lo, hi); }  // mix + FMNMX3 per pair
        }
    }
    float acc = mn;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
    if (acc == 123.456f) sink[0] = acc;
}

template <int MODE> void run(const char *name, int per_iter_packed) {
    float *sink; cudaMalloc(&sink, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 1 << 14;
    probe<MODE><<<blocks, 256>>>(iters / 8, 0.5f, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); probe<MODE><<<blocks, 256>>>(iters, 0.5f, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * 8 /*warps*/ * iters * 8.0 * per_iter_packed;
    double per_smsp_per_clk = warp_instr / (sms * 4.0) / (ms * 1e-3 * 1.965e9);
    printf("%-28s %.3f ms  packed instr/clk/SMSP (at 1.965 GHz) = %.3f\n", name, ms, per_smsp_per_clk);
    cudaFree(sink);
}
int main() {
    run<0>("FFMA2 only", 1); run<1>("FADD2 only", 1); run<2>("FMUL2 only", 1);
    run<3>("mix MUL,FMA,FMA,ADD,ADD", 5); run<4>("same as 5 FFMA2", 5); run<5>("mix + FMNMX3", 5);
    return 0;
}
