// lane_filter_probe.cu -- the lane filter's inner loop in isolation: how many cycles per (query x 8 refs) step per SMSP?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lane_filter_probe lane_filter_probe.cu ; run on B200
//   MODE 0: 16-byte query records, splat folded into the FFMA2 operand (.F32)       -- what search.cu does
//   MODE 1: 32-byte pre-splatted records, two LDS.128, true 64-bit operands
//   MODE 2: MODE 0 without the min/compare tail (FFMA2 + LDS only)
//   MODE 3: MODE 0 with 16 refs (two chunks) per lane
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b){ f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
__device__ __forceinline__ void upk(f32x2 v, float &a, float &b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ float min3(float a, float b, float c){ float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(448, 2) probe(int rounds, int nq, const float4 *__restrict__ refs, float *sink) {
    extern __shared__ float4 q[];     // nq records (16 or 32 bytes each)
    for (int i = threadIdx.x; i < nq * 2; i += blockDim.x) q[i] = make_float4(0.001f * i, -0.002f * i, 0.003f * i, -1.0f + 1e-3f * (i & 7));
    __syncthreads();
    constexpr int NR = MODE == 3 ? 16 : 8;
    float4 R[NR];
#pragma unroll
    for (int p = 0; p < NR; ++p) R[p] = refs[(threadIdx.x * NR + p) & 1023];
    unsigned acc = 0;
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < rounds; ++r) {
        const float4 *qr = q + ((r * 7 + (threadIdx.x >> 5)) & 7) * (MODE == 1 ? 64 : 32);
        unsigned mine = 0;
#pragma unroll UNROLL
        for (int l = 0; l < 32; ++l) {
            f32x2 b0, b1, b2; float thr;
            if (MODE == 1) { const float4 A = qr[2 * l], B = qr[2 * l + 1]; b0 = pk(A.x, A.y); b1 = pk(A.z, A.w); b2 = pk(B.x, B.y); thr = B.z; }
            else { const float4 A = qr[l]; b0 = pk(A.x, A.x); b1 = pk(A.y, A.y); b2 = pk(A.z, A.z); thr = A.w; }
            float d[NR];
#pragma unroll
            for (int p = 0; p < NR / 2; ++p) {
                f32x2 t = fma2(pk(R[2 * p].x, R[2 * p].y), b0, pk(R[2 * p + 1].z, R[2 * p + 1].w));
                t = fma2(pk(R[2 * p].z, R[2 * p].w), b1, t);
                t = fma2(pk(R[2 * p + 1].x, R[2 * p + 1].y), b2, t);
                upk(t, d[2 * p], d[2 * p + 1]);
            }
            if (MODE == 2) { float s = 0; for (int i = 0; i < NR; ++i) s += 0.f * d[i]; if (s == 1.f) mine |= 1u << l; }
            else {
                float m = fminf(min3(d[0], d[1], d[2]), min3(d[3], d[4], d[5]));
                m = min3(m, d[6], d[7]);
                if (NR == 16) { m = min3(m, d[8], d[9]); m = min3(m, d[10], d[11]); m = min3(m, d[12], d[13]); m = min3(m, d[14], d[15]); }
                if (m < thr) mine |= 1u << l;
            }
        }
        acc ^= mine + lane;
    }
    if (acc == 0x12345u) sink[0] = (float)acc;
}

template <int MODE, int UNROLL> void run(const char *name, int warps, const float4 *refs, float *sink) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int rounds = 4096, smem = 64 * 1024;
    auto k = probe<MODE, UNROLL>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<sms * 2, warps * 32, smem>>>(64, 512, refs, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<sms * 2, warps * 32, smem>>>(rounds, 512, refs, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double steps = (double)sms * 2 * warps * rounds * 32;            // warp-level (query x chunk-row) steps
    const double cyc = ms * 1e-3 * 1.965e9 * sms * 4 / steps;             // SMSP cycles per step
    const int refs_per_lane = MODE == 3 ? 16 : 8;
    printf("%-44s warps/CTA %2d  %.3f ms  %.1f cycles/step/SMSP  -> %.1f%% of the 8-flop FP32 peak (%s)\n", name, warps, ms, cyc,
           100.0 * (refs_per_lane * 4.0) / cyc, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float4 *refs; float *sink; cudaMalloc(&refs, 1024 * 16); cudaMemset(refs, 0x3c, 1024 * 16); cudaMalloc(&sink, 4);
    for (int w : {14, 8, 4}) {
        run<0, 4>("16 B records, .F32 splat, unroll 4", w, refs, sink);
        run<0, 8>("16 B records, .F32 splat, unroll 8", w, refs, sink);
        run<0, 32>("16 B records, .F32 splat, unroll 32", w, refs, sink);
        run<1, 8>("32 B pre-splatted records, unroll 8", w, refs, sink);
        run<2, 8>("FFMA2 + LDS only, unroll 8", w, refs, sink);
        run<3, 4>("16 refs per lane, unroll 4", w, refs, sink);
    }
    return 0;
}
