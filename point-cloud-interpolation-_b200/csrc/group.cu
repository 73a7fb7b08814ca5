// group.cu -- fused neighbourhood grouping: gather + centre subtraction + concat + layout change in ONE pass.
//
// Replaces the tail of the reference's grouping layers, which materialise four intermediates per call:
//   * Utils/Layers.py:57-66 (Group.forward):  index_points(points, ind) - new_points.view(B,S,1,C),
//     index_points(features, ind), torch.cat([...], -1), .permute(0,3,2,1).contiguous()     -> [B, 3+D, K, S]
//   * Utils/Pointnet2Utils.py:243-253 (PointNetSetAbstractionMsg.forward): the same with the feature channels
//     FIRST (torch.cat([grouped_points, grouped_xyz], -1)) and the permute left to the Conv2d              -> [B, D+3, K, S]
// Pure data movement plus one fp32 subtraction per xyz element, so the result is bit-identical to the reference's.
//
// HBM-bound: algorithmic bytes = B*S*K*(8 [idx] + 4*(3+D) [write]) + the gathered rows (re-read from L2: every ref row
// is hit ~K*S/N times).
#include "common.cuh"

namespace b200pc {

int group_bulk(const float *xyz, const float *new_xyz, const float *feat, const int64_t *idx, int B, int N, int S, int K, int D,
               int xyz_first, float *out, int force, cudaStream_t st);   // rowmove.cu; -100 = shape not served

constexpr int GROUP_S = 32;       // centres per block tile of the backward kernel

__device__ __forceinline__ long wrap_index(long i, int N) { return i < 0 ? i + N : i; }   // torch advanced indexing

// Forward: a lane owns one (centre, slot) pair -- threads are numbered like the output's two innermost axes, s fastest.
// It reads its own gathered row front to back (the row's 128-byte lines are re-used through L1) and writes channel
// after channel; every store instruction of a warp is 128 contiguous bytes of the Conv2d layout.  Streaming stores keep
// the output from evicting the feature table, which is re-read ~K*S/N times, from L2.  (A version that staged 32 x 64
// tiles through shared memory to make the row reads coalesced was 1.6x slower: profiles/r01_notes.md.)
// Scalar row reads: any D (SetConv 1 has D = 3).
__global__ void __launch_bounds__(256) group_points_scalar_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                                 const float *__restrict__ feat, const int64_t *__restrict__ idx,
                                                                 int N, int S, int K, int D, int xyz_first, long total,
                                                                 float *__restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;     // e = (b*K + k)*S + s : s fastest, like the output
    if (e >= total) return;
    const int s = (int)(e % S), k = (int)((e / S) % K), b = (int)(e / S / K);
    const int C = D + 3, xoff = xyz_first ? 0 : D, foff = xyz_first ? 3 : 0;
    const long r = wrap_index(idx[((size_t)b * S + s) * K + k], N);
    const bool ok = r >= 0 && r < N;
    const size_t row = (size_t)b * N + (ok ? r : 0), cstride = (size_t)K * S;
    float *o = out + ((size_t)b * C * K + k) * S + s;
    const float *cr = new_xyz + ((size_t)b * S + s) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) __stcs(o + (xoff + c) * cstride, __fsub_rn(ok ? __ldg(xyz + row * 3 + c) : 0.0f, cr[c]));
    for (int d = 0; d < D; ++d) __stcs(o + (foff + d) * cstride, ok ? __ldg(feat + row * D + d) : 0.0f);
}

// Same lane-per-(centre, slot) scheme with 16-byte row reads (D % 4 == 0, 16-byte aligned features).
__global__ void __launch_bounds__(256) group_points_direct4_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                                   const float4 *__restrict__ feat4, const int64_t *__restrict__ idx,
                                                                   int N, int S, int K, int D, int xyz_first, long total,
                                                                   float *__restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int s = (int)(e % S), k = (int)((e / S) % K), b = (int)(e / S / K);
    const int C = D + 3, xoff = xyz_first ? 0 : D, foff = xyz_first ? 3 : 0, D4 = D >> 2;
    const long r = wrap_index(idx[((size_t)b * S + s) * K + k], N);
    const bool ok = r >= 0 && r < N;
    const size_t row = (size_t)b * N + (ok ? r : 0), cstride = (size_t)K * S;
    float *o = out + ((size_t)b * C * K + k) * S + s;
    const float *cr = new_xyz + ((size_t)b * S + s) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) __stcs(o + (xoff + c) * cstride, __fsub_rn(ok ? __ldg(xyz + row * 3 + c) : 0.0f, cr[c]));
    const float4 *fr = feat4 + row * D4;
    float *of = o + (size_t)foff * cstride;
#pragma unroll 4
    for (int q = 0; q < D4; ++q) {
        const float4 v = ok ? __ldg(fr + q) : make_float4(0.f, 0.f, 0.f, 0.f);      // the row's 128-byte lines are re-used through L1
        __stcs(of + (4 * q + 0) * cstride, v.x); __stcs(of + (4 * q + 1) * cstride, v.y);       // streaming stores: the output must not
        __stcs(of + (4 * q + 2) * cstride, v.z); __stcs(of + (4 * q + 3) * cstride, v.w);       // evict the feature table from L2
    }
}

constexpr int GROUP_BWD_CB = 256;
constexpr int GROUP_BWD_STRIDE = GROUP_BWD_CB + 1;

// backward w.r.t. the features: grad_feat[b, idx[b,s,k], d] += grad_out[b, foff+d, k, s]
__global__ void __launch_bounds__(256) group_points_bwd_kernel(const float *__restrict__ gout, const int64_t *__restrict__ idx,
                                                               int N, int S, int K, int D, int xyz_first, float *__restrict__ gfeat) {
    __shared__ float tile[GROUP_S * GROUP_BWD_STRIDE];
    __shared__ long rows[GROUP_S];
    const int s0 = blockIdx.x * GROUP_S, k = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = D + 3, foff = xyz_first ? 3 : 0;
    if (threadIdx.x < GROUP_S) {
        const int s = s0 + threadIdx.x;
        rows[threadIdx.x] = s < S ? wrap_index(idx[((size_t)b * S + s) * K + k], N) : 0;
    }
    __syncthreads();
    for (int d0 = 0; d0 < D; d0 += GROUP_BWD_CB) {
        const int dn = min(GROUP_BWD_CB, D - d0);
        if (s0 + lane < S)
            for (int dl = warp; dl < dn; dl += 8)
                tile[lane * GROUP_BWD_STRIDE + dl] = gout[(((size_t)b * C + foff + d0 + dl) * K + k) * S + s0 + lane];
        __syncthreads();
        for (int sl = warp; sl < GROUP_S; sl += 8) {
            if (s0 + sl >= S || rows[sl] < 0 || rows[sl] >= N) continue;
            B200PC_DEV_ASSERT(rows[sl] >= 0 && rows[sl] < N && d0 + dn <= D);
            float *gr = gfeat + ((size_t)b * N + rows[sl]) * D + d0;
            for (int dl = lane; dl < dn; dl += 32) atomicAdd(gr + dl, tile[sl * GROUP_BWD_STRIDE + dl]);   // RED.ADD.F32, coalesced
        }
        __syncthreads();
    }
}

}  // namespace b200pc

using namespace b200pc;

extern "C" int b200pc_group_points(const float *xyz, const float *new_xyz, const float *feat, const int64_t *idx, int B,
                                   int N, int S, int K, int D, int xyz_first, float *out, b200pc_stream_t stream) {
    B200PC_REQUIRE(D == 0 || feat, "group_points: D=%d feature channels but no feature pointer", D);
    B200PC_REQUIRE(B >= 0 && N >= 1 && S >= 0 && K >= 0 && D >= 0, "group_points: bad sizes B=%d N=%d S=%d K=%d D=%d", B, N, S, K, D);
    if (B == 0 || S == 0 || K == 0) return B200PC_OK;
    B200PC_REQUIRE(xyz && new_xyz && idx && out, "group_points: null pointer");
    const long total = (long)B * K * S;
    B200PC_REQUIRE((total + 255) / 256 < (1L << 31), "group_points: problem too large for one launch");
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (tuning().bulk != 0 && D > 0) {                      // the asynchronous-copy path (rowmove.cu) where it is the faster one
        const int rc = group_bulk(xyz, new_xyz, feat, idx, B, N, S, K, D, xyz_first != 0, out, tuning().bulk > 0, as_stream(stream));
        if (rc != -100) return rc;
    }
    if (D > 0 && D % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0)
        group_points_direct4_kernel<<<blocks, 256, 0, as_stream(stream)>>>(xyz, new_xyz, reinterpret_cast<const float4 *>(feat), idx, N, S,
                                                                          K, D, xyz_first != 0, total, out);
    else
        group_points_scalar_kernel<<<blocks, 256, 0, as_stream(stream)>>>(xyz, new_xyz, D ? feat : nullptr, idx, N, S, K, D,
                                                                        xyz_first != 0, total, out);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}

extern "C" int b200pc_group_points_bwd(const float *grad_out, const int64_t *idx, int B, int N, int S, int K, int D,
                                       int xyz_first, float *grad_feat, b200pc_stream_t stream) {
    B200PC_REQUIRE(B >= 0 && N >= 1 && S >= 0 && K >= 0 && D >= 1, "group_points_bwd: bad sizes");
    B200PC_REQUIRE(K <= 65535 && B <= 65535, "group_points_bwd: K=%d / B=%d exceed the grid limits", K, B);
    if (B == 0 || S == 0 || K == 0) return B200PC_OK;
    B200PC_REQUIRE(grad_out && idx && grad_feat, "group_points_bwd: null pointer");
    dim3 grid((S + GROUP_S - 1) / GROUP_S, K, B);
    group_points_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, N, S, K, D, xyz_first != 0, grad_feat);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}
