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
// is hit ~K*S/N times).  A block owns 32 consecutive centres of one (batch, neighbour-slot) pair: rows are read
// coalesced along the channel axis (lane <-> channel), staged in a padded shared-memory tile, and written coalesced
// along the centre axis (lane <-> s), which is the innermost axis of the Conv2d layout.
#include "common.cuh"

namespace b200pc {

constexpr int GROUP_S = 32;       // centres per block
constexpr int GROUP_CB = 256;     // channels per pass through the shared-memory tile
constexpr int GROUP_STRIDE = GROUP_CB + 1;   // odd: lanes reading one column hit 32 different banks

__device__ __forceinline__ long wrap_index(long i, int N) { return i < 0 ? i + N : i; }   // torch advanced indexing

__global__ void __launch_bounds__(256) group_points_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                           const float *__restrict__ feat, const int64_t *__restrict__ idx,
                                                           int N, int S, int K, int D, int xyz_first, float *__restrict__ out) {
    __shared__ float tile[GROUP_S * GROUP_STRIDE];
    __shared__ long rows[GROUP_S];
    const int s0 = blockIdx.x * GROUP_S, k = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = D + 3;
    const int xoff = xyz_first ? 0 : D, foff = xyz_first ? 3 : 0;     // channel offsets of the two groups
    if (threadIdx.x < GROUP_S) {
        const int s = s0 + threadIdx.x;
        rows[threadIdx.x] = s < S ? wrap_index(idx[((size_t)b * S + s) * K + k], N) : 0;    // out of range (e.g. the ball
    }                                                                                        // query's empty-ball sentinel N):
    __syncthreads();                                                                         // the row reads as zeros, like b200pc_gather
    for (int c0 = 0; c0 < C; c0 += GROUP_CB) {
        const int cn = min(GROUP_CB, C - c0);
        // phase 1: warp w stages centres w, w+8, w+16, w+24; lanes run along the channels of the gathered row
        for (int sl = warp; sl < GROUP_S; sl += 8) {
            const int s = s0 + sl;
            if (s >= S) continue;
            const long r = rows[sl];
            const bool ok = r >= 0 && r < N;
            const float *xr = xyz + ((size_t)b * N + r) * 3, *cr = new_xyz + ((size_t)b * S + s) * 3;
            const float *fr = feat ? feat + ((size_t)b * N + r) * D : nullptr;
            for (int cl = lane; cl < cn; cl += 32) {
                const int c = c0 + cl;
                float v;
                if (c >= xoff && c < xoff + 3) v = __fsub_rn(ok ? xr[c - xoff] : 0.0f, cr[c - xoff]);
                else v = ok ? fr[c - foff] : 0.0f;
                tile[sl * GROUP_STRIDE + cl] = v;
            }
        }
        __syncthreads();
        // phase 2: one channel per warp pass, lanes run along the centres (the innermost axis of the output)
        if (s0 + lane < S)
            for (int cl = warp; cl < cn; cl += 8)
                out[(((size_t)b * C + c0 + cl) * K + k) * S + s0 + lane] = tile[lane * GROUP_STRIDE + cl];
        __syncthreads();
    }
}

// backward w.r.t. the features: grad_feat[b, idx[b,s,k], d] += grad_out[b, foff+d, k, s]
__global__ void __launch_bounds__(256) group_points_bwd_kernel(const float *__restrict__ gout, const int64_t *__restrict__ idx,
                                                               int N, int S, int K, int D, int xyz_first, float *__restrict__ gfeat) {
    __shared__ float tile[GROUP_S * GROUP_STRIDE];
    __shared__ long rows[GROUP_S];
    const int s0 = blockIdx.x * GROUP_S, k = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = D + 3, foff = xyz_first ? 3 : 0;
    if (threadIdx.x < GROUP_S) {
        const int s = s0 + threadIdx.x;
        rows[threadIdx.x] = s < S ? wrap_index(idx[((size_t)b * S + s) * K + k], N) : 0;
    }
    __syncthreads();
    for (int d0 = 0; d0 < D; d0 += GROUP_CB) {
        const int dn = min(GROUP_CB, D - d0);
        if (s0 + lane < S)
            for (int dl = warp; dl < dn; dl += 8)
                tile[lane * GROUP_STRIDE + dl] = gout[(((size_t)b * C + foff + d0 + dl) * K + k) * S + s0 + lane];
        __syncthreads();
        for (int sl = warp; sl < GROUP_S; sl += 8) {
            if (s0 + sl >= S || rows[sl] < 0 || rows[sl] >= N) continue;
            float *gr = gfeat + ((size_t)b * N + rows[sl]) * D + d0;
            for (int dl = lane; dl < dn; dl += 32) atomicAdd(gr + dl, tile[sl * GROUP_STRIDE + dl]);   // RED.ADD.F32, coalesced
        }
        __syncthreads();
    }
}

}  // namespace b200pc

using namespace b200pc;

extern "C" int b200pc_group_points(const float *xyz, const float *new_xyz, const float *feat, const int64_t *idx, int B,
                                   int N, int S, int K, int D, int xyz_first, float *out, b200pc_stream_t stream) {
    B200PC_REQUIRE(xyz && new_xyz && idx && out, "group_points: null pointer");
    B200PC_REQUIRE(D == 0 || feat, "group_points: D=%d feature channels but no feature pointer", D);
    B200PC_REQUIRE(B >= 0 && N >= 1 && S >= 0 && K >= 0 && D >= 0, "group_points: bad sizes B=%d N=%d S=%d K=%d D=%d", B, N, S, K, D);
    B200PC_REQUIRE(K <= 65535 && B <= 65535, "group_points: K=%d / B=%d exceed the grid limits", K, B);
    if (B == 0 || S == 0 || K == 0) return B200PC_OK;
    dim3 grid((S + GROUP_S - 1) / GROUP_S, K, B);
    group_points_kernel<<<grid, 256, 0, as_stream(stream)>>>(xyz, new_xyz, D ? feat : nullptr, idx, N, S, K, D, xyz_first != 0, out);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}

extern "C" int b200pc_group_points_bwd(const float *grad_out, const int64_t *idx, int B, int N, int S, int K, int D,
                                       int xyz_first, float *grad_feat, b200pc_stream_t stream) {
    B200PC_REQUIRE(grad_out && idx && grad_feat, "group_points_bwd: null pointer");
    B200PC_REQUIRE(B >= 0 && N >= 1 && S >= 0 && K >= 0 && D >= 1, "group_points_bwd: bad sizes");
    B200PC_REQUIRE(K <= 65535 && B <= 65535, "group_points_bwd: K=%d / B=%d exceed the grid limits", K, B);
    if (B == 0 || S == 0 || K == 0) return B200PC_OK;
    dim3 grid((S + GROUP_S - 1) / GROUP_S, K, B);
    group_points_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, N, S, K, D, xyz_first != 0, grad_feat);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}
