// hotloop_probe.cu -- the search hot loop in isolation (no TMA, no drain): what is the practical ceiling?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b){ f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
__device__ __forceinline__ void upk(f32x2 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b){ f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
__device__ __forceinline__ float min3(float a, float b, float c){ float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}

// VAR 0: form0 (5 packed) + min3 tree + FSETP/mask   (the shipped hot loop)
// VAR 1: same but t-space (4 packed: no query-norm add)
// VAR 2: form0, NO min/compare at all (just xor-accumulate the results)  -> pure math+LDS ceiling
// VAR 3: form0 with min over 16-ref chunks
template <int VAR, int Q>
__global__ void __launch_bounds__(256) hot(const float4 *__restrict__ g, int iters, unsigned *out) {
    __shared__ float4 tile[1024];   // 512 refs as pairs
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = g[i];
    __syncthreads();
    f32x2 qa[Q], qb[Q], qcz[Q], qn[Q]; float tau[Q]; unsigned mask[Q]; f32x2 acc[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) { float v = 0.001f * (threadIdx.x + 7 * j); qa[j] = pk(v, v); qb[j] = pk(-v, -v); qcz[j] = pk(2*v, 2*v); qn[j] = pk(v*v, v*v); tau[j] = -1.0f + v; mask[j] = 0; acc[j] = 0; }
    for (int it = 0; it < iters; ++it) {
        unsigned bit = 1u;
        const float4 *cp = tile;
#pragma unroll 2
        for (int c = 0; c < 64; ++c, cp += 8, bit = (bit << 1) | (bit >> 31)) {
            float4 A[4], B[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) { A[p] = cp[2*p]; B[p] = cp[2*p+1]; }
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                float d[8];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    f32x2 X = pk(A[p].x, A[p].y), Y = pk(A[p].z, A[p].w), Z = pk(B[p].x, B[p].y), W = pk(B[p].z, B[p].w);
                    f32x2 t = mul2(X, qa[j]); t = fma2(Y, qb[j], t); t = fma2(Z, qcz[j], t); t = add2(t, W);
                    if (VAR != 1) t = add2(t, qn[j]);
                    if (VAR == 2) acc[j] ^= t;
                    upk(t, d[2*p], d[2*p+1]);
                }
                if (VAR != 2) {
                    float m = fminf(min3(d[0], d[1], d[2]), min3(d[6], d[7], min3(d[3], d[4], d[5])));
                    if (m < tau[j]) mask[j] |= bit;
                }
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < Q; ++j) r ^= mask[j] ^ (unsigned)acc[j] ^ (unsigned)(acc[j] >> 32);
    if (r == 0x12345u) out[0] = r;
}

template <int VAR, int Q> void run(const char *name, const float4 *g, int threads, int blocks_per_sm) {
    unsigned *out; cudaMalloc(&out, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm, iters = 200;
    hot<VAR, Q><<<blocks, threads>>>(g, 10, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); hot<VAR, Q><<<blocks, threads>>>(g, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double pairs = (double)blocks * threads * Q * 512.0 * iters;
    printf("%-34s thr=%d bps=%d  %.3f ms  %.2f Tpairs/s  %.1f TFLOP/s (8 flop/pair) = %.1f%% of 74.4\n", name, threads, blocks_per_sm, ms, pairs / ms / 1e9, pairs * 8 / ms / 1e9, pairs * 8 / ms / 1e9 / 74.4 * 100);
    cudaFree(out);
}
int main() {
    float4 *g; cudaMalloc(&g, 1024 * 16); cudaMemset(g, 0, 1024 * 16);
    run<0,1>("form0 Q=1", g, 256, 4); run<0,2>("form0 Q=2", g, 256, 4); run<0,4>("form0 Q=4", g, 256, 2);
    run<0,2>("form0 Q=2 (8 warps/SM)", g, 256, 1); run<0,2>("form0 Q=2 (16 warps/SM)", g, 256, 2); run<0,2>("form0 Q=2 (48 warps/SM)", g, 256, 6);
    run<1,2>("t-space Q=2", g, 256, 4); run<1,4>("t-space Q=4", g, 256, 2);
    run<2,2>("form0 no-min Q=2", g, 256, 4); run<2,4>("form0 no-min Q=4", g, 256, 2);
    return 0;
}
