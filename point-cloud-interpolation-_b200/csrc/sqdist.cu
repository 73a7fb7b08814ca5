// sqdist.cu -- square_distance(src, dst) as a standalone op (Utils/Pointnet2Utils.py:20-41).
//
// The searches never materialise this matrix; the op exists for API parity and for tests.  It is
// bound by the HBM WRITE of the [B,N,M] result (1 GiB at 16384^2), so the kernel is organised
// around 16-byte streaming stores: a thread owns 4 consecutive dst columns (their coordinates and
// norms live in registers) and walks down ROWS src rows held in shared memory.
// Rounding order is torch-CPU's: dot = fma(z,z', fma(y,y', x*x')); d = ((-2*dot) + |s|^2) + |d|^2.
#include "common.cuh"

namespace b200pc {

constexpr int SQD_ROWS = 16;      // src rows per CTA
constexpr int SQD_THREADS = 256;  // 4 columns each -> 1024 columns per CTA

__device__ __forceinline__ float sq_norm_rn(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__global__ void __launch_bounds__(SQD_THREADS) sqdist_kernel(const float *__restrict__ src, const float *__restrict__ dst,
                                                             int N, int M, float *__restrict__ out, int vec_ok) {
    __shared__ float4 srow[SQD_ROWS];  // {-2x, -2y, -2z, |s|^2}
    const int b = blockIdx.z;
    const int n0 = blockIdx.y * SQD_ROWS;
    const int m0 = (blockIdx.x * SQD_THREADS + threadIdx.x) * 4;
    if (threadIdx.x < SQD_ROWS) {
        const int n = n0 + threadIdx.x;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) {
            const float *p = src + ((size_t)b * N + n) * 3;
            v = make_float4(-2.f * p[0], -2.f * p[1], -2.f * p[2], sq_norm_rn(p[0], p[1], p[2]));
        }
        srow[threadIdx.x] = v;
    }
    float dx[4], dy[4], dz[4], dn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = m0 + j;
        dx[j] = dy[j] = dz[j] = dn[j] = 0.f;
        if (m < M) {
            const float *p = dst + ((size_t)b * M + m) * 3;
            dx[j] = p[0]; dy[j] = p[1]; dz[j] = p[2];
            dn[j] = sq_norm_rn(dx[j], dy[j], dz[j]);
        }
    }
    __syncthreads();
    if (m0 >= M) return;
#pragma unroll 4
    for (int r = 0; r < SQD_ROWS; ++r) {
        const int n = n0 + r;
        if (n >= N) break;
        const float4 s = srow[r];
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = __fmul_rn(s.x, dx[j]);
            t = __fmaf_rn(s.y, dy[j], t);
            t = __fmaf_rn(s.z, dz[j], t);
            d[j] = __fadd_rn(__fadd_rn(t, s.w), dn[j]);
        }
        float *o = out + ((size_t)b * N + n) * M + m0;
        if (vec_ok && m0 + 3 < M) {
            stg_stream(reinterpret_cast<float4 *>(o), make_float4(d[0], d[1], d[2], d[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (m0 + j < M) o[j] = d[j];
        }
    }
}

}  // namespace b200pc

using namespace b200pc;

extern "C" int b200pc_square_distance(const float *src, const float *dst, int B, int N, int M, float *out,
                                      b200pc_stream_t stream) {
    B200PC_REQUIRE(B >= 0 && N >= 0 && M >= 0, "square_distance: bad sizes");
    if (B == 0 || N == 0 || M == 0) return B200PC_OK;
    B200PC_REQUIRE(src && dst && out, "square_distance: null pointer");
    const int vec_ok = (M % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    dim3 grid((M + SQD_THREADS * 4 - 1) / (SQD_THREADS * 4), (N + SQD_ROWS - 1) / SQD_ROWS, B);
    B200PC_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "square_distance: problem too large for one launch");
    sqdist_kernel<<<grid, SQD_THREADS, 0, as_stream(stream)>>>(src, dst, N, M, out, vec_ok);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}
