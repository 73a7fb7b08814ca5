// rowmove.cu -- the HBM-bound row movers on the TMA path (sm_100a): whole feature rows travel global -> shared with one
// cp.async.bulk per row (SASS UBLKCP), completion is counted on per-warp mbarriers, and rows leave either as ONE bulk
// store per group of consecutive output rows (index_points) or after a register pass (three_interpolate, group_points).
//
//   gather_bulk_kernel   index_points            Utils/Pointnet2Utils.py:44-61 (= pytorch3d knn_gather, Utils/Layers.py:396,434)
//   interp_bulk_kernel   three_interpolate       Utils/Layers.py:187-188, Utils/Pointnet2Utils.py:304
//   group_bulk_kernel    Group.forward tail      Utils/Layers.py:57-66, SA-MSG grouping Utils/Pointnet2Utils.py:243-253
//
// Why: the register-path kernels of gather.cu / group.cu buy memory-level parallelism with registers (8 rows in flight =
// 73 registers, 28 % of the warp slots; 4 rows of three_interpolate = 121 registers) and pay one L1 wavefront per 32-byte
// sector of every gathered row (group_points: 580 wavefront-cycles per 8 KB of output -- the measured 3.5 TB/s is exactly
// that bound).  A bulk copy costs one instruction per ROW, holds no register while in flight, and writes shared memory
// without passing the LSU, so a CTA keeps 100-190 KB in flight with 4 warps.  Every warp runs its own ring of stages
// (no block barrier after set-up); work is dealt to the warps in groups small enough that the last wave is > 95 % full.
#include <type_traits>

#include "common.cuh"

namespace b200pc {

constexpr int RM_WARPS = 4;                       // warps per CTA, one CTA per SM
constexpr int RM_BAR_AREA = 256;                  // mbarriers: RM_WARPS x up to 8 stages x 8 bytes
constexpr size_t RM_SMEM_BUDGET = 200 * 1024;     // per CTA, leaves room for the driver's reservation

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_zero16(uint32_t a) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// per-warp ring set-up: NST mbarriers (one arrival each: the lane that posts the byte count)
template <int NST>
__device__ __forceinline__ uint32_t ring_init(unsigned char *smem, int warp, int lane) {
    const uint32_t bar0 = smem_u32(smem) + (uint32_t)(warp * NST) * 8u;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(bar0 + 8u * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    return bar0;
}

// ---------------------------------------------------------------------------------------------
// index_points: out[row, :] = points[b(row), idx[row], :]
// A group = RS consecutive output rows = one contiguous block of the output: RS bulk loads in, ONE bulk store out.
// ---------------------------------------------------------------------------------------------
template <int NST>
__global__ void __launch_bounds__(RM_WARPS * 32, 1) gather_bulk_kernel(const char *__restrict__ points, const int64_t *__restrict__ idx,
                                                                       int N, uint32_t row_bytes, int RS, long R, long rows_total,
                                                                       char *__restrict__ out, int *__restrict__ oob) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stage_bytes = (uint32_t)RS * row_bytes;
    const uint32_t bar0 = ring_init<NST>(smem, warp, lane);
    const uint32_t data0 = smem_u32(smem) + RM_BAR_AREA + (uint32_t)(warp * NST) * stage_bytes;
    const long groups = (rows_total + RS - 1) / RS;
    const long gw = (long)blockIdx.x * RM_WARPS + warp, nw = (long)gridDim.x * RM_WARPS;
    const long n_my = groups > gw ? (groups - gw + nw - 1) / nw : 0;

    // source row (b*N + i) of this lane's row in the warp's j-th group; -1: index out of range (row of zeros), -2: no row
    auto load_src = [&](long j) -> long {
        const long row = (gw + j * nw) * RS + lane;
        if (j >= n_my || lane >= RS || row >= rows_total) return -2;
        long i = idx[row];
        if (i < 0) i += N;
        if (i < 0 || i >= N) return -1;
        return (row / R) * N + i;
    };
    auto issue = [&](long j, long src) {
        const int st = (int)(j % NST);
        const uint32_t bar = bar0 + 8u * st, dst = data0 + (uint32_t)st * stage_bytes + (uint32_t)lane * row_bytes;
        const unsigned valid = __ballot_sync(FULL, src >= 0);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(valid) * row_bytes);
        __syncwarp();
        if (src >= 0) bulk_g2s(dst, points + (size_t)src * row_bytes, row_bytes, bar);
        else if (src == -1) {
            for (uint32_t o = 0; o < row_bytes; o += 16) sts_zero16(dst + o);
            if (oob) *oob = 1;
        }
    };

    long pre[NST - 1];
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) pre[p] = load_src(p);          // all index loads of the prologue are in flight together
#pragma unroll
    for (int p = 0; p < NST - 1; ++p)
        if (p < n_my) issue(p, pre[p]);
    long nxt = load_src(NST - 1);
    for (long i = 0; i < n_my; ++i) {
        const int st = (int)(i % NST);
        mbar_wait(bar0 + 8u * st, (uint32_t)((i / NST) & 1));
        fence_proxy_async();                                          // rows zero-filled by lanes are visible to the bulk store
        __syncwarp();
        if (lane == 0) {
            const long row0 = (gw + i * nw) * RS;
            const long nrows = rows_total - row0 < RS ? rows_total - row0 : RS;
            bulk_s2g(out + (size_t)row0 * row_bytes, data0 + (uint32_t)st * stage_bytes, (uint32_t)nrows * row_bytes);
            bulk_commit();
        }
        const long j = i + NST - 1;                                   // refill the stage whose store was committed one iteration ago
        if (j < n_my) {
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            issue(j, nxt);
            nxt = load_src(j + 1);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------
// three_interpolate: out[row, :] = (f[i0]*w0 + f[i1]*w1) + f[i2]*w2
// A group = RS dense rows: 3*RS bulk loads, the (cleaned) weights ride in the stage header, the warp mixes one row
// per iteration from shared memory and stores it with coalesced 16-byte stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rm_mix3(float a, float wa, float b, float wb, float c, float wc) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb)), __fmul_rn(c, wc));
}

template <int NST>
__global__ void __launch_bounds__(RM_WARPS * 32, 1) interp_bulk_kernel(const char *__restrict__ feat, const int64_t *__restrict__ idx,
                                                                       const float *__restrict__ w, int S, uint32_t row_bytes, int RS,
                                                                       long N, long rows_total, float4 *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t head_bytes = 128;                                  // [RS*3] weights (RS <= 8 -> 96 bytes)
    const uint32_t stage_bytes = head_bytes + 3u * RS * row_bytes;
    const uint32_t bar0 = ring_init<NST>(smem, warp, lane);
    const uint32_t data0 = smem_u32(smem) + RM_BAR_AREA + (uint32_t)(warp * NST) * stage_bytes;
    const long groups = (rows_total + RS - 1) / RS;
    const long gw = (long)blockIdx.x * RM_WARPS + warp, nw = (long)gridDim.x * RM_WARPS;
    const long n_my = groups > gw ? (groups - gw + nw - 1) / nw : 0;
    const int C4 = (int)(row_bytes >> 4);

    // lane = 3*u + jn handles neighbour jn of the group's u-th row
    struct Nb { long src; float wt; };
    auto load_nb = [&](long j) -> Nb {
        Nb nb; nb.src = -2; nb.wt = 0.f;
        const long row = (gw + j * nw) * RS + lane / 3;
        if (j >= n_my || lane >= 3 * RS || row >= rows_total) return nb;
        long i = idx[row * 3 + lane % 3];
        float wt = w[row * 3 + lane % 3];
        if ((unsigned long long)i >= (unsigned long long)S) {       // negative indices wrap once; still out of range: contributes nothing
            if (i < 0) i += S;
            if (i < 0 || i >= S) { i = 0; wt = 0.0f; }
        }
        nb.src = (row / N) * S + i; nb.wt = wt;
        return nb;
    };
    auto issue = [&](long j, const Nb &nb) {
        const int st = (int)(j % NST);
        const uint32_t bar = bar0 + 8u * st, base = data0 + (uint32_t)st * stage_bytes;
        const unsigned valid = __ballot_sync(0xffffffffu, nb.src >= 0);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(valid) * row_bytes);
        __syncwarp();
        if (nb.src >= 0) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4u * lane), "f"(nb.wt) : "memory");
            bulk_g2s(base + head_bytes + (uint32_t)lane * row_bytes, feat + (size_t)nb.src * row_bytes, row_bytes, bar);
        }
    };

    Nb pre[NST];                                                     // all index / weight loads of the prologue are in flight together
#pragma unroll
    for (int p = 0; p < NST; ++p) pre[p] = load_nb(p);
#pragma unroll
    for (int p = 0; p < NST; ++p)
        if (p < n_my) issue(p, pre[p]);
    Nb nxt = load_nb(NST);
    for (long i = 0; i < n_my; ++i) {
        const int st = (int)(i % NST);
        const uint32_t base = data0 + (uint32_t)st * stage_bytes;
        mbar_wait(bar0 + 8u * st, (uint32_t)((i / NST) & 1));
        __syncwarp();                                                 // the weights written by the other lanes are visible
        const long row0 = (gw + i * nw) * RS;
        const int nrows = (int)(rows_total - row0 < RS ? rows_total - row0 : RS);
#pragma unroll 2
        for (int u = 0; u < nrows; ++u) {
            const float w0 = lds_f1(base + 12u * u), w1 = lds_f1(base + 12u * u + 4), w2 = lds_f1(base + 12u * u + 8);
            const uint32_t r0 = base + head_bytes + (uint32_t)(3 * u) * row_bytes;
            for (int col = lane; col < C4; col += 32) {
                const float4 a = lds_f4(r0 + 16u * col), b = lds_f4(r0 + row_bytes + 16u * col), c = lds_f4(r0 + 2u * row_bytes + 16u * col);
                float4 o;
                o.x = rm_mix3(a.x, w0, b.x, w1, c.x, w2); o.y = rm_mix3(a.y, w0, b.y, w1, c.y, w2);
                o.z = rm_mix3(a.z, w0, b.z, w1, c.z, w2); o.w = rm_mix3(a.w, w0, b.w, w1, c.w, w2);
                stg_stream(out + (size_t)(row0 + u) * C4 + col, o);
            }
        }
        __syncwarp();                                                 // every lane is done reading the stage
        const long j = i + NST;                                       // the same stage takes the group NST ahead
        if (j < n_my) {
            issue(j, nxt);
            nxt = load_nb(j + 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// group_points: out[b, c, k, s] from rows gathered by idx[b, s, k]  (the Conv2d layout)
// A unit = 32 consecutive centres s x a range of slots k.  Per slot: 32 bulk loads of feature rows into a padded tile
// (row stride chosen so that 16-byte reads of 8 lanes hit 32 different banks), each lane then reads ITS row with
// 16-byte shared loads and writes channel after channel; every store instruction is 128 contiguous bytes.  The unit's
// index block is read once, coalesced, into shared memory.  xyz rows (12 bytes) are read through the LSU.
// ---------------------------------------------------------------------------------------------
template <int NST>
__global__ void __launch_bounds__(RM_WARPS * 32, 1) group_bulk_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                                      const char *__restrict__ feat, const int64_t *__restrict__ idx,
                                                                      int N, int S, int K, int D, int xyz_first, int ksplit,
                                                                      long units_total, float *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    const uint32_t row_stride = row_bytes + (((D >> 2) & 1) ? 32u : 16u);     // (stride / 16) odd -> conflict-free 16-byte reads
    const uint32_t tile_bytes = 32u * row_stride;
    const int kper = (K + ksplit - 1) / ksplit;                       // slots per unit
    const uint32_t idx_stride = (uint32_t)(kper + 1) * 8u;            // padded: 64-bit reads of a column are conflict-free
    const uint32_t idx_bytes = (32u * idx_stride + 127u) & ~127u;
    const uint32_t warp_bytes = idx_bytes + NST * tile_bytes;
    const uint32_t bar0 = ring_init<NST>(smem, warp, lane);
    const uint32_t ibuf = smem_u32(smem) + RM_BAR_AREA + (uint32_t)warp * warp_bytes;
    const uint32_t data0 = ibuf + idx_bytes;
    const long gw = (long)blockIdx.x * RM_WARPS + warp, nw = (long)gridDim.x * RM_WARPS;
    const int s_tiles = (S + 31) / 32;
    const int C = D + 3, xoff = xyz_first ? 0 : D, foff = xyz_first ? 3 : 0, D4 = D >> 2;
    const size_t cstride = (size_t)K * S;
    uint32_t uses = 0;                                                // tiles issued so far by this warp (stage = uses % NST)
    uint32_t done = 0;                                                // tiles consumed so far

    for (long unit = gw; unit < units_total; unit += nw) {
        const int kpart = (int)(unit % ksplit);
        const long bt = unit / ksplit;
        const int stile = (int)(bt % s_tiles), b = (int)(bt / s_tiles);
        const int s0 = stile * 32, k0 = kpart * kper;
        const int nk = K - k0 < kper ? K - k0 : kper;
        const int s = s0 + lane;
        const bool live = s < S;
        // ---- the unit's index block idx[b, s0..s0+31, k0..k0+nk) -> shared, coalesced over (s, k) ----
        __syncwarp();
        for (int e = lane; e < 32 * nk; e += 32) {
            const int sl = e / nk, kk = e - sl * nk;
            long r = -1;
            if (s0 + sl < S) {
                r = idx[((size_t)b * S + s0 + sl) * K + k0 + kk];
                if (r < 0) r += N;
                if (r >= N) r = -1;
            }
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(ibuf + (uint32_t)sl * idx_stride + 8u * kk), "l"(r) : "memory");
        }
        __syncwarp();
        auto my_row = [&](int kk) -> long {
            long r;
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(ibuf + (uint32_t)lane * idx_stride + 8u * kk));
            return r;
        };
        auto issue = [&](int kk) {
            const int st = (int)(uses % NST);
            const uint32_t bar = bar0 + 8u * st, dst = data0 + (uint32_t)st * tile_bytes + (uint32_t)lane * row_stride;
            const long r = my_row(kk);
            const unsigned valid = __ballot_sync(0xffffffffu, r >= 0);
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(valid) * row_bytes);
            __syncwarp();
            if (r >= 0) bulk_g2s(dst, feat + ((size_t)b * N + r) * row_bytes, row_bytes, bar);
            ++uses;
        };
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (live) {
            const float *cr = new_xyz + ((size_t)b * S + s) * 3;
            cx = cr[0]; cy = cr[1]; cz = cr[2];
        }
        const int pro = nk < NST ? nk : NST;
        for (int kk = 0; kk < pro; ++kk) issue(kk);
        for (int kk = 0; kk < nk; ++kk) {
            const long r = my_row(kk);
            const bool ok = r >= 0;
            float *o = out + ((size_t)b * C * K + (k0 + kk)) * S + s;
            // xyz channels through the LSU while the feature rows are (still) in flight
            if (live) {
                const float *xr = xyz + ((size_t)b * N + (ok ? r : 0)) * 3;
                __stcs(o + (size_t)(xoff + 0) * cstride, __fsub_rn(ok ? __ldg(xr + 0) : 0.0f, cx));
                __stcs(o + (size_t)(xoff + 1) * cstride, __fsub_rn(ok ? __ldg(xr + 1) : 0.0f, cy));
                __stcs(o + (size_t)(xoff + 2) * cstride, __fsub_rn(ok ? __ldg(xr + 2) : 0.0f, cz));
            }
            const int st = (int)(done % NST);
            mbar_wait(bar0 + 8u * st, (done / NST) & 1u);
            const uint32_t rowa = data0 + (uint32_t)st * tile_bytes + (uint32_t)lane * row_stride;
            float *of = o + (size_t)foff * cstride;
            if (live) {
#pragma unroll 4
                for (int q = 0; q < D4; ++q) {
                    const float4 v = ok ? lds_f4(rowa + 16u * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    __stcs(of + (size_t)(4 * q + 0) * cstride, v.x); __stcs(of + (size_t)(4 * q + 1) * cstride, v.y);
                    __stcs(of + (size_t)(4 * q + 2) * cstride, v.z); __stcs(of + (size_t)(4 * q + 3) * cstride, v.w);
                }
            }
            ++done;
            __syncwarp();                                             // every lane is done reading the stage
            if (kk + NST < nk) issue(kk + NST);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch plans
// ---------------------------------------------------------------------------------------------
static int pick_rs(long rows, int rs_max, long warps) {
    int rs = rs_max;
    while (rs > 4 && (rows + rs - 1) / rs < 6 * warps) rs >>= 1;      // at least ~6 groups per warp: the last wave is > 85 % full
    return rs;
}

template <typename F>
static int dispatch_nst(int nst, F &&f) {
    if (nst >= 8) return f(std::integral_constant<int, 8>());
    if (nst >= 6) return f(std::integral_constant<int, 6>());
    if (nst >= 4) return f(std::integral_constant<int, 4>());
    if (nst >= 3) return f(std::integral_constant<int, 3>());
    return f(std::integral_constant<int, 2>());
}
static int round_nst(int nst) { return nst >= 8 ? 8 : nst >= 6 ? 6 : nst >= 4 ? 4 : nst >= 3 ? 3 : 2; }

// index_points through the bulk path.  Returns -100 when the shape does not qualify (caller falls back to gather.cu).
int gather_bulk(const float *points, const int64_t *idx, int B, int N, int C, long R, float *out, int *oob, cudaStream_t st) {
    const uint32_t row_bytes = (uint32_t)C * 4u;
    if (C % 4 != 0 || row_bytes < 128 || row_bytes > 48 * 1024) return -100;
    if ((reinterpret_cast<uintptr_t>(points) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -100;
    const long rows = (long)B * R;
    const int sms = sm_count();
    int rs_max = 32;
    while (rs_max > 1 && (size_t)rs_max * row_bytes > 16 * 1024) rs_max >>= 1;
    const int RS = pick_rs(rows, rs_max, (long)sms * RM_WARPS);
    const size_t stage = (size_t)RS * row_bytes;
    int nst = (int)((RM_SMEM_BUDGET - RM_BAR_AREA) / (RM_WARPS * stage));
    if (nst < 2) return -100;
    nst = round_nst(nst);
    const long groups = (rows + RS - 1) / RS;
    const int grid = (int)((groups + RM_WARPS - 1) / RM_WARPS < sms ? (groups + RM_WARPS - 1) / RM_WARPS : sms);
    const size_t smem = RM_BAR_AREA + (size_t)RM_WARPS * nst * stage;
    return dispatch_nst(nst, [&](auto tag) -> int {
        constexpr int NST = decltype(tag)::value;
        auto kern = gather_bulk_kernel<NST>;
        B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, RM_WARPS * 32, smem, st>>>(reinterpret_cast<const char *>(points), idx, N, row_bytes, RS, R, rows,
                                                reinterpret_cast<char *>(out), oob);
        B200PC_LAUNCH_CHECK();
        return B200PC_OK;
    });
}

int interp_bulk(const float *feat, const int64_t *idx, const float *w, int B, int S, int N, int C, float *out, cudaStream_t st) {
    const uint32_t row_bytes = (uint32_t)C * 4u;
    if (C % 4 != 0 || row_bytes < 256 || row_bytes > 8 * 1024) return -100;
    if ((reinterpret_cast<uintptr_t>(feat) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -100;
    const long rows = (long)B * N;
    const int sms = sm_count();
    int rs_max = 8;
    while (rs_max > 1 && (size_t)3 * rs_max * row_bytes > 24 * 1024) rs_max >>= 1;
    const int RS = pick_rs(rows, rs_max, (long)sms * RM_WARPS);
    const size_t stage = 128 + (size_t)3 * RS * row_bytes;
    int nst = (int)((RM_SMEM_BUDGET - RM_BAR_AREA) / (RM_WARPS * stage));
    if (nst < 2) return -100;
    nst = round_nst(nst);
    const long groups = (rows + RS - 1) / RS;
    const int grid = (int)((groups + RM_WARPS - 1) / RM_WARPS < sms ? (groups + RM_WARPS - 1) / RM_WARPS : sms);
    const size_t smem = RM_BAR_AREA + (size_t)RM_WARPS * nst * stage;
    return dispatch_nst(nst, [&](auto tag) -> int {
        constexpr int NST = decltype(tag)::value;
        auto kern = interp_bulk_kernel<NST>;
        B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, RM_WARPS * 32, smem, st>>>(reinterpret_cast<const char *>(feat), idx, w, S, row_bytes, RS, (long)N, rows,
                                                reinterpret_cast<float4 *>(out));
        B200PC_LAUNCH_CHECK();
        return B200PC_OK;
    });
}

int group_bulk(const float *xyz, const float *new_xyz, const float *feat, const int64_t *idx, int B, int N, int S, int K, int D,
               int xyz_first, float *out, cudaStream_t st) {
    if (D < 32 || D % 4 != 0 || D > 1024 || (reinterpret_cast<uintptr_t>(feat) & 15) || K > 256) return -100;
    const int sms = sm_count();
    const long warps = (long)sms * RM_WARPS;
    const long base_units = (long)B * ((S + 31) / 32);
    int ksplit = 1;
    while (ksplit < K && base_units * ksplit < 6 * warps) ksplit <<= 1;
    if (ksplit > K) ksplit = K;
    const int kper = (K + ksplit - 1) / ksplit;
    const long units = base_units * ksplit;
    const uint32_t row_stride = (uint32_t)D * 4u + (((D >> 2) & 1) ? 32u : 16u);
    const size_t tile = (size_t)32 * row_stride;
    const size_t idx_bytes = ((size_t)32 * (kper + 1) * 8 + 127) & ~(size_t)127;
    const size_t avail = (RM_SMEM_BUDGET - RM_BAR_AREA) / RM_WARPS;
    if (avail < idx_bytes + 2 * tile) return -100;
    const int nst = round_nst((int)((avail - idx_bytes) / tile));
    const int grid = (int)((units + RM_WARPS - 1) / RM_WARPS < sms ? (units + RM_WARPS - 1) / RM_WARPS : sms);
    const size_t smem = RM_BAR_AREA + (size_t)RM_WARPS * (idx_bytes + nst * tile);
    return dispatch_nst(nst, [&](auto tag) -> int {
        constexpr int NST = decltype(tag)::value;
        auto kern = group_bulk_kernel<NST>;
        B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, RM_WARPS * 32, smem, st>>>(xyz, new_xyz, reinterpret_cast<const char *>(feat), idx, N, S, K, D, xyz_first, ksplit,
                                                units, out);
        B200PC_LAUNCH_CHECK();
        return B200PC_OK;
    });
}

}  // namespace b200pc
