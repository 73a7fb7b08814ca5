// rowmove.cu -- the HBM-bound row movers on the asynchronous-copy path (sm_100a): feature rows travel global -> shared
// with warp-cooperative 16-byte cp.async (SASS LDGSTS: no register is held while a row is in flight, a 512-byte row is
// ONE instruction), stages complete through cp.async groups, and rows leave either as ONE bulk TMA store
// (cp.async.bulk, SASS UBLKCP) per group of consecutive output rows (index_points) or after a register pass
// (three_interpolate, group_points).
//
//   gather_async_kernel   index_points            Utils/Pointnet2Utils.py:44-61 (= pytorch3d knn_gather, Utils/Layers.py:396,434)
//   interp_async_kernel   three_interpolate       Utils/Layers.py:187-188, Utils/Pointnet2Utils.py:304
//   group_async_kernel    Group.forward tail      Utils/Layers.py:57-66, SA-MSG grouping Utils/Pointnet2Utils.py:243-253
//
// Why: the register-path kernels of gather.cu / group.cu buy memory-level parallelism with registers (8 rows in flight =
// 73 registers, 28 % of the warp slots; 4 rows of three_interpolate = 121 registers), and group_points pays one L1
// wavefront per 32-byte sector of every gathered row because each lane walks its OWN row (580 wavefront-cycles per 8 KB
// of output: the measured 3.5 TB/s is exactly that bound).  Here a warp copies whole rows cooperatively (coalesced), a
// CTA keeps 100-190 KB in flight, every warp runs its own ring of stages (no block barrier at all), and work is dealt
// out in groups small enough that the last wave is > 85 % full.
// Measured dead end, kept out: one cp.async.bulk PER ROW (TMA gather).  The TMA unit serves ~1 request per 46-60 cycles
// per SM whatever its size, i.e. 10 B/clk/SM for 512-byte rows -- three_interpolate went from 60 us to 163 us,
// group_points from 108 us to 167 us (profiles/r02_notes.md).  Bulk copies are used where one request moves >= 4 KB.
#include <type_traits>

#include "common.cuh"

namespace b200pc {

constexpr int RM_BAR_AREA = 0;
constexpr size_t RM_SMEM_BUDGET = 200 * 1024;     // per CTA (one CTA per SM), leaves room for the driver's reservation

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
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

// One row (row_bytes, a multiple of 16) global -> shared by the whole warp; src == nullptr: a row of zeros.
__device__ __forceinline__ void warp_copy_row(uint32_t dst, const char *src, uint32_t row_bytes, int lane) {
    if (src) { for (uint32_t o = 16u * lane; o < row_bytes; o += 512u) cp_async16(dst + o, src + o); }
    else { for (uint32_t o = 16u * lane; o < row_bytes; o += 512u) sts_zero16(dst + o); }
}

// ---------------------------------------------------------------------------------------------
// All index arithmetic below is 32-bit (the launchers check rows, B*N and B*S against 2^31): a 64-bit division is a
// ~100-instruction subroutine call, and these kernels run 4-8 warps per SM -- they are bound by the instruction stream
// of a single warp per scheduler, not by a pipe (ncu of the first version: 258 instructions per interpolated row).
// ---------------------------------------------------------------------------------------------
// index_points: out[row, :] = points[b(row), idx[row], :]
// A group = RS consecutive output rows = one contiguous block of the output: RS row copies in, ONE bulk store out.
// ---------------------------------------------------------------------------------------------
constexpr int GA_WARPS = 8;

template <int NST>
__global__ void __launch_bounds__(GA_WARPS * 32, 1) gather_async_kernel(const char *__restrict__ points, const int64_t *__restrict__ idx,
                                                                        int N, uint32_t row_bytes, int RS, int R, int rows_total,
                                                                        char *__restrict__ out, int *__restrict__ oob) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stage_bytes = (uint32_t)RS * row_bytes;
    const uint32_t data0 = smem_u32(smem) + (uint32_t)(warp * NST) * stage_bytes;
    const int groups = (rows_total + RS - 1) / RS;
    const int gw = blockIdx.x * GA_WARPS + warp, nw = gridDim.x * GA_WARPS;
    const int n_my = groups > gw ? (groups - gw + nw - 1) / nw : 0;

    // source row (b*N + i) of this lane's row in the warp's j-th group; -1: index out of range (row of zeros), -2: no row
    auto load_src = [&](int j) -> int {
        const int row = (gw + j * nw) * RS + lane;
        if (j >= n_my || lane >= RS || row >= rows_total) return -2;
        long i = idx[row];
        if (i < 0) i += N;
        if (i < 0 || i >= N) return -1;
        return (int)((unsigned)row / (unsigned)R) * N + (int)i;
    };
    auto issue = [&](int j, int st, int src) {                        // always commits one cp.async group (possibly empty)
        if (j < n_my) {
            const uint32_t dst = data0 + (uint32_t)st * stage_bytes;
            if (src == -1 && oob) *oob = 1;
            for (int r = 0; r < RS; ++r) {
                const int sr = __shfl_sync(0xffffffffu, src, r);
                if (sr != -2) warp_copy_row(dst + (uint32_t)r * row_bytes, sr >= 0 ? points + (size_t)sr * row_bytes : nullptr, row_bytes, lane);
            }
        }
        cp_async_commit();
    };

    int pre[NST - 1];
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) pre[p] = load_src(p);          // all index loads of the prologue are in flight together
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) issue(p, p, pre[p]);
    int nxt = load_src(NST - 1);
    int st = 0;                                                       // stage of group i; the refill goes to st - 1 (mod NST)
    for (int i = 0; i < n_my; ++i) {
        cp_async_wait<NST - 2>();                                     // group i has landed (this lane's share)
        fence_proxy_async();                                          // ... and is visible to the bulk store
        __syncwarp();
        if (lane == 0) {
            const int row0 = (gw + i * nw) * RS;
            const int nrows = rows_total - row0 < RS ? rows_total - row0 : RS;
            bulk_s2g(out + (size_t)row0 * row_bytes, data0 + (uint32_t)st * stage_bytes, (uint32_t)nrows * row_bytes);
            bulk_commit();
            bulk_wait_read<1>();                                      // the store committed one iteration ago has read its stage
        }
        __syncwarp();
        issue(i + NST - 1, st == 0 ? NST - 1 : st - 1, nxt);          // ... which now takes the group NST-1 ahead
        nxt = load_src(i + NST);
        st = st + 1 == NST ? 0 : st + 1;
    }
    if (lane == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------
// three_interpolate: out[row, :] = (f[i0]*w0 + f[i1]*w1) + f[i2]*w2
// A group = RS dense rows: 3*RS row copies, the (cleaned) weights ride in the stage header, the warp mixes one row
// per iteration from shared memory and stores it with coalesced 16-byte stores.
// ---------------------------------------------------------------------------------------------
constexpr int IN_WARPS = 8;

__device__ __forceinline__ float rm_mix3(float a, float wa, float b, float wb, float c, float wc) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb)), __fmul_rn(c, wc));
}

template <int NST>
__global__ void __launch_bounds__(IN_WARPS * 32, 1) interp_async_kernel(const char *__restrict__ feat, const int64_t *__restrict__ idx,
                                                                        const float *__restrict__ w, int S, uint32_t row_bytes, int RS,
                                                                        int N, int rows_total, float4 *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t head_bytes = 128;                                  // [RS*3] weights (RS <= 8 -> 96 bytes)
    const uint32_t stage_bytes = head_bytes + 3u * RS * row_bytes;
    const uint32_t data0 = smem_u32(smem) + (uint32_t)(warp * NST) * stage_bytes;
    const int groups = (rows_total + RS - 1) / RS;
    const int gw = blockIdx.x * IN_WARPS + warp, nw = gridDim.x * IN_WARPS;
    const int n_my = groups > gw ? (groups - gw + nw - 1) / nw : 0;
    const int C4 = (int)(row_bytes >> 4);
    const int lu = lane / 3, lj = lane - 3 * lu;                     // lane = 3*u + jn handles neighbour jn of the group's u-th row

    struct Nb { int src; float wt; };
    auto load_nb = [&](int j) -> Nb {
        Nb nb; nb.src = -2; nb.wt = 0.f;
        const int row = (gw + j * nw) * RS + lu;
        if (j >= n_my || lane >= 3 * RS || row >= rows_total) return nb;
        long i = idx[(size_t)row * 3 + lj];
        float wt = w[(size_t)row * 3 + lj];
        if ((unsigned long long)i >= (unsigned long long)S) {       // negative indices wrap once; still out of range: contributes nothing
            if (i < 0) i += S;
            if (i < 0 || i >= S) { i = 0; wt = 0.0f; }
        }
        nb.src = (int)((unsigned)row / (unsigned)N) * S + (int)i; nb.wt = wt;
        return nb;
    };
    auto issue = [&](int j, int st, const Nb &nb) {
        if (j < n_my) {
            const uint32_t base = data0 + (uint32_t)st * stage_bytes;
            if (nb.src >= 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 4u * lane), "f"(nb.wt) : "memory");
            const int n3 = 3 * RS;
            for (int r = 0; r < n3; ++r) {
                const int sr = __shfl_sync(0xffffffffu, nb.src, r);
                if (sr >= 0) warp_copy_row(base + head_bytes + (uint32_t)r * row_bytes, feat + (size_t)sr * row_bytes, row_bytes, lane);
            }
        }
        cp_async_commit();
    };

    Nb pre[NST - 1];                                                 // all index / weight loads of the prologue are in flight together
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) pre[p] = load_nb(p);
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) issue(p, p, pre[p]);
    Nb nxt = load_nb(NST - 1);
    int st = 0;
    for (int i = 0; i < n_my; ++i) {
        // the stage consumed one iteration ago takes the group NST-1 ahead, before this group is waited for
        issue(i + NST - 1, st == 0 ? NST - 1 : st - 1, nxt);
        nxt = load_nb(i + NST);
        cp_async_wait<NST - 1>();                                     // group i has landed (this lane's share)
        __syncwarp();                                                 // ... every lane's share, and the weights in the header
        const uint32_t base = data0 + (uint32_t)st * stage_bytes;
        const int row0 = (gw + i * nw) * RS;
        const int nrows = rows_total - row0 < RS ? rows_total - row0 : RS;
        float4 *orow = out + (size_t)row0 * C4;
#pragma unroll 2
        for (int u = 0; u < nrows; ++u) {
            const float w0 = lds_f1(base + 12u * u), w1 = lds_f1(base + 12u * u + 4), w2 = lds_f1(base + 12u * u + 8);
            const uint32_t r0 = base + head_bytes + (uint32_t)(3 * u) * row_bytes;
            for (int col = lane; col < C4; col += 32) {
                const float4 a = lds_f4(r0 + 16u * col), b = lds_f4(r0 + row_bytes + 16u * col), c = lds_f4(r0 + 2u * row_bytes + 16u * col);
                float4 o;
                o.x = rm_mix3(a.x, w0, b.x, w1, c.x, w2); o.y = rm_mix3(a.y, w0, b.y, w1, c.y, w2);
                o.z = rm_mix3(a.z, w0, b.z, w1, c.z, w2); o.w = rm_mix3(a.w, w0, b.w, w1, c.w, w2);
                stg_stream(orow + (size_t)u * C4 + col, o);
            }
        }
        __syncwarp();                                                 // every lane is done reading the stage
        st = st + 1 == NST ? 0 : st + 1;
    }
}

// ---------------------------------------------------------------------------------------------
// group_points: out[b, c, k, s] from rows gathered by idx[b, s, k]  (the Conv2d layout)
// A unit = 32 consecutive centres s x a range of slots k.  Per slot: the 32 feature rows are copied by the whole warp
// (coalesced) into a padded tile (row stride chosen so that 16-byte reads of 8 lanes hit 32 different banks); each lane
// then reads ITS row with 16-byte shared loads and writes channel after channel: every store instruction is 128
// contiguous bytes.  The unit's index block is read once, coalesced, into shared memory (as 32-bit row numbers b*N + i).
// xyz rows (12 bytes) are read through the LSU.
// ---------------------------------------------------------------------------------------------
constexpr int GR_MAX_WARPS = 16;        // warps per CTA are a launch parameter: as many as shared memory allows (the kernel is bound by
                                        // one warp's instruction latency, not by a pipe: ncu 12 % warps active, 25 % issue active at 8 warps)

template <int NST>
__global__ void __launch_bounds__(GR_MAX_WARPS * 32, 1) group_async_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                                                       const char *__restrict__ feat, const int64_t *__restrict__ idx,
                                                                       int N, int S, int K, int D, int xyz_first, int ksplit,
                                                                       int units_total, float *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    const uint32_t row_stride = row_bytes + (((D >> 2) & 1) ? 32u : 16u);     // (stride / 16) odd -> conflict-free 16-byte reads
    const uint32_t tile_bytes = 32u * row_stride;
    const int kper = (K + ksplit - 1) / ksplit;                       // slots per unit
    const uint32_t idx_stride = (uint32_t)(kper | 1) * 4u;            // odd number of words: column reads are conflict-free
    const uint32_t idx_bytes = (32u * idx_stride + 127u) & ~127u;
    const uint32_t warp_bytes = idx_bytes + NST * tile_bytes;
    const uint32_t ibuf = smem_u32(smem) + (uint32_t)warp * warp_bytes;
    const uint32_t data0 = ibuf + idx_bytes;
    const int nwarps = (int)(blockDim.x >> 5);
    const int gw = blockIdx.x * nwarps + warp, nw = gridDim.x * nwarps;
    const int s_tiles = (S + 31) / 32;
    const int C = D + 3, xoff = xyz_first ? 0 : D, foff = xyz_first ? 3 : 0, D4 = D >> 2;
    const size_t cstride = (size_t)K * S;
    // how the warp copies a tile: cpr 16-byte chunks per row; rpp rows per instruction when a row is < 512 bytes
    const int cpr = (int)(row_bytes >> 4);
    const int lanes_per_row = cpr >= 32 ? 32 : cpr;
    const int rpp = 32 / lanes_per_row;
    const int sub = lane / lanes_per_row, chunk = lane - sub * lanes_per_row;

    for (int unit = gw; unit < units_total; unit += nw) {
        const int bt = unit / ksplit, kpart = unit - bt * ksplit;
        const int b = bt / s_tiles, stile = bt - b * s_tiles;
        const int s0 = stile * 32, k0 = kpart * kper;
        const int nk = K - k0 < kper ? K - k0 : kper;
        const int s = s0 + lane;
        const bool live = s < S;
        // ---- the unit's index block idx[b, s0..s0+31, k0..k0+nk) -> shared as row numbers b*N + i (-1: none) ----
        __syncwarp();
        {
            int sl = 0, kk = lane;
            while (kk >= nk) { kk -= nk; ++sl; }
            while (sl < 32) {
                int r = -1;
                if (s0 + sl < S) {
                    long i = idx[((size_t)b * S + s0 + sl) * K + k0 + kk];
                    if (i < 0) i += N;
                    if (i >= 0 && i < N) r = b * N + (int)i;
                }
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(ibuf + (uint32_t)sl * idx_stride + 4u * kk), "r"(r) : "memory");
                kk += 32;
                while (kk >= nk) { kk -= nk; ++sl; }
            }
        }
        __syncwarp();
        auto row_of = [&](int sl, int kk) -> int {
            int r;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(ibuf + (uint32_t)sl * idx_stride + 4u * kk));
            return r;
        };
        auto issue = [&](int kk, int st) {                            // always commits one cp.async group (possibly empty)
            if (kk < nk) {
                const uint32_t tile = data0 + (uint32_t)st * tile_bytes;
                if (cpr >= 32) {
                    for (int sl = 0; sl < 32; ++sl) {
                        const int r = row_of(sl, kk);                 // broadcast read
                        if (r >= 0) warp_copy_row(tile + (uint32_t)sl * row_stride, feat + (size_t)r * row_bytes, row_bytes, lane);
                    }
                } else {
                    uint32_t dst = tile + (uint32_t)sub * row_stride + 16u * chunk;
                    for (int sl = sub; sl < 32; sl += rpp, dst += rpp * row_stride) {   // rpp rows per instruction
                        const int r = row_of(sl, kk);
                        if (r >= 0) cp_async16(dst, feat + (size_t)r * row_bytes + 16u * chunk);
                    }
                }
            }
            cp_async_commit();
        };
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (live) {
            const float *cr = new_xyz + ((size_t)b * S + s) * 3;
            cx = cr[0]; cy = cr[1]; cz = cr[2];
        }
#pragma unroll
        for (int p = 0; p < NST - 1; ++p) issue(p, p);
        int st = 0;
        float *o = out + ((size_t)b * C * K + k0) * S + s;
        for (int kk = 0; kk < nk; ++kk, o += S) {
            issue(kk + NST - 1, st == 0 ? NST - 1 : st - 1);         // into the stage consumed one iteration ago
            const int r = row_of(lane, kk);
            const bool ok = r >= 0;
            // xyz channels through the LSU while the feature rows are (still) in flight
            if (live) {
                const float *xr = xyz + (size_t)(ok ? r : 0) * 3;
                __stcs(o + (size_t)(xoff + 0) * cstride, __fsub_rn(ok ? __ldg(xr + 0) : 0.0f, cx));
                __stcs(o + (size_t)(xoff + 1) * cstride, __fsub_rn(ok ? __ldg(xr + 1) : 0.0f, cy));
                __stcs(o + (size_t)(xoff + 2) * cstride, __fsub_rn(ok ? __ldg(xr + 2) : 0.0f, cz));
            }
            cp_async_wait<NST - 1>();                                 // tile kk has landed (this lane's share)
            __syncwarp();                                             // ... every lane's share
            const uint32_t rowa = data0 + (uint32_t)st * tile_bytes + (uint32_t)lane * row_stride;
            float *of = o + (size_t)foff * cstride;
            if (live) {
#pragma unroll 4
                for (int q = 0; q < D4; ++q, of += 4 * cstride) {
                    const float4 v = ok ? lds_f4(rowa + 16u * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    __stcs(of, v.x); __stcs(of + cstride, v.y); __stcs(of + 2 * cstride, v.z); __stcs(of + 3 * cstride, v.w);
                }
            }
            __syncwarp();                                             // every lane is done reading the stage
            st = st + 1 == NST ? 0 : st + 1;
        }
        cp_async_wait<0>();                                           // drain the (empty) tail groups before the ring restarts
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

template <typename F>
static int dispatch_nst12(int nst, F &&f) {
    if (nst >= 2) return f(std::integral_constant<int, 2>());
    return f(std::integral_constant<int, 1>());
}

// index_points through the asynchronous path.  Returns -100 when the shape does not qualify (caller falls back to gather.cu).
int gather_bulk(const float *points, const int64_t *idx, int B, int N, int C, long R, float *out, int *oob, cudaStream_t st) {
    const uint32_t row_bytes = (uint32_t)C * 4u;
    if (C % 4 != 0 || row_bytes < 128 || row_bytes > 16 * 1024) return -100;
    if ((reinterpret_cast<uintptr_t>(points) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -100;
    const long rows = (long)B * R;
    if (rows >= (1L << 31) - 64 || (long)B * N >= (1L << 31) || R >= (1L << 31)) return -100;      // 32-bit row arithmetic in the kernel
    const int sms = sm_count();
    int rs_max = 32;
    while (rs_max > 1 && (size_t)rs_max * row_bytes > 16 * 1024) rs_max >>= 1;
    const int RS = pick_rs(rows, rs_max, (long)sms * GA_WARPS);
    const size_t stage = (size_t)RS * row_bytes;
    int nst = (int)(RM_SMEM_BUDGET / (GA_WARPS * stage));
    if (nst < 3) return -100;
    nst = round_nst(nst);
    const long groups = (rows + RS - 1) / RS;
    const int grid = (int)((groups + GA_WARPS - 1) / GA_WARPS < sms ? (groups + GA_WARPS - 1) / GA_WARPS : sms);
    const size_t smem = (size_t)GA_WARPS * nst * stage;
    return dispatch_nst(nst, [&](auto tag) -> int {
        constexpr int NST = decltype(tag)::value < 3 ? 3 : decltype(tag)::value;
        auto kern = gather_async_kernel<NST>;
        B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, GA_WARPS * 32, smem, st>>>(reinterpret_cast<const char *>(points), idx, N, row_bytes, RS, (int)R, (int)rows,
                                                reinterpret_cast<char *>(out), oob);
        B200PC_LAUNCH_CHECK();
        return B200PC_OK;
    });
}

int interp_bulk(const float *feat, const int64_t *idx, const float *w, int B, int S, int N, int C, float *out, cudaStream_t st) {
    const uint32_t row_bytes = (uint32_t)C * 4u;
    if (C % 4 != 0 || row_bytes < 256 || row_bytes > 4 * 1024) return -100;
    if ((reinterpret_cast<uintptr_t>(feat) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -100;
    const long rows = (long)B * N;
    if (rows >= (1L << 31) - 64 || (long)B * S >= (1L << 31)) return -100;      // 32-bit row arithmetic in the kernel
    const int sms = sm_count();
    int rs_max = 8;
    while (rs_max > 1 && (size_t)3 * rs_max * row_bytes > 12 * 1024) rs_max >>= 1;
    const int RS = pick_rs(rows, rs_max, (long)sms * IN_WARPS);
    const size_t stage = 128 + (size_t)3 * RS * row_bytes;
    int nst = (int)(RM_SMEM_BUDGET / (IN_WARPS * stage));
    if (nst < 2) return -100;
    nst = round_nst(nst);
    const long groups = (rows + RS - 1) / RS;
    const int grid = (int)((groups + IN_WARPS - 1) / IN_WARPS < sms ? (groups + IN_WARPS - 1) / IN_WARPS : sms);
    const size_t smem = (size_t)IN_WARPS * nst * stage;
    return dispatch_nst(nst, [&](auto tag) -> int {
        constexpr int NST = decltype(tag)::value;
        auto kern = interp_async_kernel<NST>;
        B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, IN_WARPS * 32, smem, st>>>(reinterpret_cast<const char *>(feat), idx, w, S, row_bytes, RS, N, (int)rows,
                                                reinterpret_cast<float4 *>(out));
        B200PC_LAUNCH_CHECK();
        return B200PC_OK;
    });
}

// force == 0: only where measured faster than the register path (16 warps fit: D <= 64 -- 90.7 vs 102.7 us at C3, D = 64;
// with 10 warps at D = 128 it is 210 vs 183 us); force != 0: every shape the kernel can serve (A/B, tests).
int group_bulk(const float *xyz, const float *new_xyz, const float *feat, const int64_t *idx, int B, int N, int S, int K, int D,
               int xyz_first, float *out, int force, cudaStream_t st) {
    if (!force && (long)B * S * K < 32768) return -100;                  // small groupings: one wave of the register kernel is quicker
    if (D < 16 || D % 4 != 0 || D > 1024 || (reinterpret_cast<uintptr_t>(feat) & 15) || K > 256) return -100;
    const int cpr = D / 4;
    if (cpr < 32 && 32 % cpr != 0) return -100;                        // rows shorter than 512 bytes must tile an instruction evenly
    if ((long)B * N >= (1L << 31) || (long)B * ((S + 31) / 32) * K >= (1L << 31)) return -100;     // 32-bit row arithmetic in the kernel
    const int sms = sm_count();
    const uint32_t row_stride = (uint32_t)D * 4u + (((D >> 2) & 1) ? 32u : 16u);
    const size_t tile = (size_t)32 * row_stride;
    // Many warps with a shallow ring beat few warps with a deep one here (the warp's own dependent chain idx -> copy -> LDS ->
    // STG is the limiter): one stage per warp, as many warps as fit; two stages only when 16 warps still fit with them.
    int ksplit = 1, kper = K, nst = 1, nwarps = 0;
    long units = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const long base_units = (long)B * ((S + 31) / 32);
        const long warps_guess = (long)sms * (nwarps > 0 ? nwarps : GR_MAX_WARPS);
        ksplit = 1;
        while (ksplit < K && base_units * ksplit < 6 * warps_guess) ksplit <<= 1;
        if (ksplit > K) ksplit = K;
        kper = (K + ksplit - 1) / ksplit;
        units = base_units * ksplit;
        const size_t idx_bytes = ((size_t)32 * (kper | 1) * 4 + 127) & ~(size_t)127;
        int w1 = (int)(RM_SMEM_BUDGET / (idx_bytes + tile)), w2 = (int)(RM_SMEM_BUDGET / (idx_bytes + 2 * tile));
        nst = w2 >= GR_MAX_WARPS ? 2 : 1;
        nwarps = nst == 2 ? w2 : w1;
        if (nwarps > GR_MAX_WARPS) nwarps = GR_MAX_WARPS;
        if (nwarps < 4 || (!force && nwarps < GR_MAX_WARPS)) return -100;
    }
    const size_t idx_bytes = ((size_t)32 * (kper | 1) * 4 + 127) & ~(size_t)127;
    const int grid = (int)((units + nwarps - 1) / nwarps < sms ? (units + nwarps - 1) / nwarps : sms);
    const size_t smem = (size_t)nwarps * (idx_bytes + nst * tile);
    return dispatch_nst12(nst, [&](auto tag) -> int {
        constexpr int NST = decltype(tag)::value;
        auto kern = group_async_kernel<NST>;
        B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, nwarps * 32, smem, st>>>(xyz, new_xyz, reinterpret_cast<const char *>(feat), idx, N, S, K, D, xyz_first, ksplit,
                                              (int)units, out);
        B200PC_LAUNCH_CHECK();
        return B200PC_OK;
    });
}

}  // namespace b200pc
