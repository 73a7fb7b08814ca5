// search.cu -- brute-force neighbour search on FP32 CUDA cores (sm_100a).
//
// One streaming skeleton serves four reference call sites (file:line relative to the reference):
//   * kNN grouping           Utils/Layers.py:50-53          form 0, top-k
//   * three-NN               Utils/Layers.py:180-182,       form 1, top-3
//                            Utils/Pointnet2Utils.py:297-299
//   * pytorch3d knn_points   Utils/Layers.py:220 ...        form 2, top-K   (also Chamfer, K=1)
//   * query_ball_point       Utils/Pointnet2Utils.py:88-108 form 1, first-nsample-by-index
//
// Design (why it looks the way it does is argued in DESIGN.md):
//   1. pack_refs_kernel turns refs [B,N,3] into 16-byte records grouped in PAIRS:
//      {x0,x1,y0,y1} {z0,z1,w0,w1} (w = |r|^2 rounded as torch does, or 0 for the direct form),
//      padded to a whole tile with records whose distance is +inf; the 4 pairs of a chunk are
//      XOR-swizzled by the chunk number so that the drain's per-lane re-reads avoid bank conflicts.
//   2. search_kernel: one TMA-producer warp streams 8 KB tiles of those records into a 4-stage
//      shared-memory ring with cp.async.bulk + mbarrier; consumer threads own Q queries each and
//      evaluate 2 refs per instruction with FFMA2/FMUL2/FADD2 in EXACTLY the reference's rounding
//      order.  The refs of a tile are read with broadcast LDS.128 (no bank conflicts).
//   3. The per-pair cost is kept at "distance + half a min": a chunk of 8 refs is reduced with
//      FMNMX3 and compared once against the query's current threshold tau; a hit only sets a bit
//      in a register mask.  There is no branch and no list traffic in the hot loop.
//   4. After at most 32 chunks the mask is drained: hit chunks are re-evaluated (bit-identical
//      arithmetic) and true candidates are insertion-sorted into a per-query list that lives in
//      shared memory, laid out [rank][query] so that lanes never conflict.  Refs are visited in
//      increasing index order and insertion uses strict '<', which yields the total order
//      (distance, index): ties go to the lower index, deterministically.
//   5. When there are too few queries to fill 148 SMs the ref range is split over gridDim.z and
//      a small merge kernel combines the partial lists.
#include "search.cuh"

#include <math.h>
#include <math_constants.h>
#include <stdlib.h>

namespace b200pc {

constexpr int TILE = 512;                  // refs per shared-memory tile
constexpr int TILE_BYTES = TILE * 16;      // 8 KB
constexpr int STAGES = 3;                  // ring depth
constexpr int CHUNK = 8;                   // refs per threshold test
constexpr int CHUNKS_PER_TILE = TILE / CHUNK;
constexpr int MAX_SPLIT = 32;
constexpr int BAR_BYTES = 128;

// ---------------------------------------------------------------------------------------------
// 1. ref packing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float torch_sq_norm(float x, float y, float z) {
    // torch.sum(p ** 2, -1): three roundings of the squares, then (xx + yy) + zz, never fused
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__global__ void pack_refs_kernel(const float *__restrict__ ref, int N, int n_pad, int form,
                                 float4 *__restrict__ packed) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;  // pair index
    if (p >= n_pad / 2) return;
    int b = blockIdx.y;
    const float *r = ref + (size_t)b * N * 3;
    float x[2], y[2], z[2], w[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int i = 2 * p + h;
        if (i < N) {
            x[h] = r[i * 3 + 0]; y[h] = r[i * 3 + 1]; z[h] = r[i * 3 + 2];
            w[h] = form == B200PC_FORM_DIRECT ? 0.0f : torch_sq_norm(x[h], y[h], z[h]);
        } else if (form == B200PC_FORM_DIRECT) {
            x[h] = CUDART_INF_F; y[h] = 0.0f; z[h] = 0.0f; w[h] = 0.0f;   // (inf - q)^2 = inf
        } else {
            x[h] = 0.0f; y[h] = 0.0f; z[h] = 0.0f; w[h] = CUDART_INF_F;   // 0 + inf = inf
        }
    }
    // a chunk = 4 pairs = 128 bytes = all 32 banks.  Pair q of chunk c is stored in pair-slot q ^ (c & 3), so
    // that lanes re-visiting DIFFERENT chunks in the drain spread over the banks instead of all starting at
    // bank 0 (measured: 11.4 wavefronts per LDS.128 without the swizzle).  The hot loop takes a min over the
    // whole chunk, so the order of the pairs inside a chunk does not matter to it.
    const int chunk = p >> 2;
    if (form == B200PC_FORM_DIRECT) {
        // the direct form needs no norm word: 24 bytes per pair, a chunk is 6 records
        // [A0 A1 Z01 A2 A3 Z23] with A = {x0,x1,y0,y1} and Z = {z0,z1 of the first pair, z0,z1 of the second}
        // (6 instead of 8 LDS.128 per chunk in the hot loop).  96-byte chunks rotate over the banks by themselves.
        const int pc = p & 3;
        float4 *base = packed + ((size_t)b * (n_pad / 8) + chunk) * 6 + (pc >> 1) * 3;
        base[pc & 1] = make_float4(x[0], x[1], y[0], y[1]);
        float2 *zz = reinterpret_cast<float2 *>(base + 2) + (pc & 1);
        *zz = make_float2(z[0], z[1]);
        return;
    }
    const int slot = (p & 3) ^ (chunk & 3);
    float4 *o = packed + ((size_t)b * (n_pad / 2) + (size_t)chunk * 4 + slot) * 2;
    o[0] = make_float4(x[0], x[1], y[0], y[1]);
    o[1] = make_float4(z[0], z[1], w[0], w[1]);
}

// shared-memory layout of a chunk (8 refs) per distance form
template <int FORM>
struct Lay {
    static constexpr int HREC = FORM == B200PC_FORM_DIRECT ? 3 : 4;   // 16-byte records per half chunk (4 refs)
    static constexpr int REC = 2 * HREC;                               // per chunk
    static constexpr int TILE_COPY = (TILE / 8) * REC * 16;            // bytes one tile occupies in the packed stream
};

// ---------------------------------------------------------------------------------------------
// 2. distance of one query against a PAIR of refs, in the reference's rounding order
// ---------------------------------------------------------------------------------------------
struct QueryConst {  // per-query splatted constants
    f32x2 a0, a1, a2, a3;
};

template <int FORM>
__device__ __forceinline__ QueryConst make_query(float x, float y, float z) {
    QueryConst q;
    if (FORM == B200PC_FORM_DIRECT) {
        q.a0 = splat2(-x); q.a1 = splat2(-y); q.a2 = splat2(-z); q.a3 = 0ull;
    } else {
        // -2*(s.d) == s.(-2d) exactly: scaling by a power of two commutes with every rounding
        q.a0 = splat2(-2.0f * x); q.a1 = splat2(-2.0f * y); q.a2 = splat2(-2.0f * z);
        q.a3 = splat2(torch_sq_norm(x, y, z));
    }
    return q;
}

template <int FORM>
__device__ __forceinline__ f32x2 pair_dist(const float4 &A, const float4 &Bv, const QueryConst &q) {
    f32x2 X = pack2(A.x, A.y), Y = pack2(A.z, A.w), Z = pack2(Bv.x, Bv.y);
    if (FORM == B200PC_FORM_DIRECT) {
        // pytorch3d: d = fma(dz,dz, fma(dy,dy, dx*dx)); (r-q)^2 == (q-r)^2 bit for bit
        f32x2 dx = add2(X, q.a0), dy = add2(Y, q.a1), dz = add2(Z, q.a2);
        f32x2 t = mul2(dx, dx);
        t = fma2(dy, dy, t);
        return fma2(dz, dz, t);
    }
    // torch CPU (MKL sgemm, K=3): dot = fma(z,z', fma(y,y', x*x')); then two separate adds
    f32x2 W = pack2(Bv.z, Bv.w);
    f32x2 t = mul2(X, q.a0);
    t = fma2(Y, q.a1, t);
    t = fma2(Z, q.a2, t);
    if (FORM == B200PC_FORM_REF_NORM_FIRST) {
        t = add2(t, W);         // dist += |src|^2   (src = refs at the kNN call site)
        return add2(t, q.a3);   // dist += |dst|^2
    } else {
        t = add2(t, q.a3);      // src = queries (ball query, three-NN)
        return add2(t, W);
    }
}

// Hot-loop variant.  At the kNN call site (form 0) the query norm is the LAST addend and is constant
// per query, so the prefilter ranks on t = fl(dot' + |r|^2) and skips that add; the threshold it is
// compared with is moved into t-space conservatively (prefilter_threshold), and the drain re-tests the
// exact distance, so results are unchanged -- the hot loop just does 4 packed instructions instead of 5.
template <int FORM>
__device__ __forceinline__ f32x2 pair_prefilter(const float4 &A, const float4 &Bv, const QueryConst &q) {
    if (FORM != B200PC_FORM_REF_NORM_FIRST) return pair_dist<FORM>(A, Bv, q);
    f32x2 X = pack2(A.x, A.y), Y = pack2(A.z, A.w), Z = pack2(Bv.x, Bv.y), W = pack2(Bv.z, Bv.w);
    f32x2 t = mul2(X, q.a0);
    t = fma2(Y, q.a1, t);
    t = fma2(Z, q.a2, t);
    return add2(t, W);
}
// every t with fl(t + nq) < tau satisfies t < (tau - nq) + (2|tau| + |nq|) * 2^-24; one more bit of margin
template <int FORM>
__device__ __forceinline__ float prefilter_threshold(float tau, float nq) {
    if (FORM != B200PC_FORM_REF_NORM_FIRST) return tau;
    return (tau - nq) + (2.0f * fabsf(tau) + fabsf(nq)) * 1.1920929e-7f;
}

__device__ __forceinline__ f32x2 pair_dist_direct(const float4 &A, f32x2 Z, const QueryConst &q) {
    // pytorch3d: d = fma(dz,dz, fma(dy,dy, dx*dx)); (r-q)^2 == (q-r)^2 bit for bit
    const f32x2 dx = add2(pack2(A.x, A.y), q.a0), dy = add2(pack2(A.z, A.w), q.a1), dz = add2(Z, q.a2);
    f32x2 t = mul2(dx, dx);
    t = fma2(dy, dy, t);
    return fma2(dz, dz, t);
}

// prefilter value of the 4 refs of half a chunk given its HREC records -> min of the four
template <int FORM>
__device__ __forceinline__ float half_chunk_min(const float4 (&H)[Lay<FORM>::HREC], const QueryConst &q) {
    float d0, d1, d2, d3;
    if (FORM == B200PC_FORM_DIRECT) {
        unpack2(pair_dist_direct(H[0], pack2(H[2].x, H[2].y), q), d0, d1);
        unpack2(pair_dist_direct(H[1], pack2(H[2].z, H[2].w), q), d2, d3);
    } else {
        unpack2(pair_prefilter<FORM>(H[0], H[Lay<FORM>::HREC - 3], q), d0, d1);
        unpack2(pair_prefilter<FORM>(H[Lay<FORM>::HREC - 2], H[Lay<FORM>::HREC - 1], q), d2, d3);
    }
    return fminf(min3(d0, d1, d2), d3);
}

template <int FORM>
__device__ __forceinline__ float chunk_min(const float4 (&R)[Lay<FORM>::REC], const QueryConst &q) {
    constexpr int H = Lay<FORM>::HREC;
    float4 lo[H], hi[H];
#pragma unroll
    for (int i = 0; i < H; ++i) { lo[i] = R[i]; hi[i] = R[H + i]; }
    return fminf(half_chunk_min<FORM>(lo, q), half_chunk_min<FORM>(hi, q));
}

// exact distances of logical pair p (refs 2p, 2p+1) of a chunk whose records start at `cb` (drain path)
template <int FORM>
__device__ __forceinline__ f32x2 chunk_pair_dist(const float4 *cb, int chunk, int p, const QueryConst &q) {
    if (FORM == B200PC_FORM_DIRECT) {
        const float4 *h = cb + (p >> 1) * 3;
        const float4 Z = h[2];
        return pair_dist_direct(h[p & 1], (p & 1) ? pack2(Z.z, Z.w) : pack2(Z.x, Z.y), q);
    }
    const float4 *pp = cb + 2 * (p ^ (chunk & 3));       // un-swizzle: where pair p of this chunk lives
    return pair_dist<FORM>(pp[0], pp[1], q);
}

// ---------------------------------------------------------------------------------------------
// 3. the streaming search kernel
// ---------------------------------------------------------------------------------------------
struct SearchArgs {
    const float4 *packed;  // [B][n_pad/2][2]
    const float *qry;      // [B][S][3]
    int N, n_pad, S, k;
    float r2;              // ball radius^2 (fp32)
    int debug_nodrain;     // measurement only: start with tau = -inf so nothing ever hits
    int n_split, tiles_per_split;
    int64_t *idx_out;      // [B][S][k]   (n_split == 1)
    float *dist_out;       // [B][S][k] or null
    float *part_d;         // [B][S][n_split][k]  (top-k partial lists)
    int *part_i;           // [B][S][n_split][k]
    int *part_cnt;         // [B][S][n_split]     (ball partial counts)
};

// Order-preserving map fp32 -> u32 (negative expanded-form distances included), so that a
// 64-bit key (map(d) << 32 | index) realises the total order (distance, index) with one compare.
__device__ __forceinline__ uint32_t order_key(float d) {
    const uint32_t b = __float_as_uint(d);
    return b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
constexpr unsigned long long HEAP_SENTINEL = 0xFF800000FFFFFFFFull;  // (+inf, max index)

__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// max-heap of 64-bit keys in shared memory: element e of this query lives at byte address hb + e*SB.
// Put `nk` at the root of a heap holding nB/SB elements and sift it down (depth ceil(log2 n): every
// lane of a warp runs the same short loop).  PADDED = the element right after the heap is readable and
// holds key 0, so the right child needs no bounds check (true while streaming, false during heapsort).
template <bool PADDED>
__device__ __forceinline__ void heap_sift_root(uint32_t hb, uint32_t SB, uint32_t nB, unsigned long long nk) {
    uint32_t pos = 0;
    while (true) {
        const uint32_t l = 2 * pos + SB;
        if (l >= nB) break;
        unsigned long long kc = lds_u64(hb + l);
        uint32_t c = l;
        if (PADDED || l + SB < nB) {
            const unsigned long long kr = lds_u64(hb + l + SB);
            if (kr > kc) { kc = kr; c = l + SB; }
        }
        if (kc <= nk) break;
        sts_u64(hb + pos, kc);
        pos = c;
    }
    sts_u64(hb + pos, nk);
}

constexpr int CAND_CAP = 32;  // per-query buffer: one u16 entry (chunk << 8 | candidate mask) per hit chunk of a sub-tile

template <int FORM>
__device__ __forceinline__ float tile_dist(const float4 *tp, int off, const QueryConst &q) {
    float lo, hi;
    const int chunk = off >> 3;
    unpack2(chunk_pair_dist<FORM>(tp + chunk * Lay<FORM>::REC, chunk, (off >> 1) & 3, q), lo, hi);   // same packed arithmetic as the hot loop
    return (off & 1) ? hi : lo;
}

constexpr int MAX_WARPS = 16;

// Every warp is a consumer; there is no dedicated producer warp.  The ring is refilled by whichever
// warp happens to release a stage LAST: after arriving on the stage's "empty" barrier each warp
// probes it once (no spinning) and, if the phase is complete, claims the refill with a shared-memory
// compare-and-swap and issues the bulk copy.  blockDim.x = warps * 32 is a RUNTIME value so that the
// planner can size the grid as whole waves of resident CTAs.
template <int FORM, int MODE, int Q>
__global__ void __launch_bounds__(MAX_WARPS * 32) search_kernel(const SearchArgs P) {
    const int NCW = (int)(blockDim.x >> 5);       // warps
    const int NCT = NCW * 32;                     // threads
    const int QPB = NCT * Q;                      // queries per block
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem[];
    const float4 *tiles = reinterpret_cast<const float4 *>(smem);
    const uint32_t bar_base = smem_u32(smem + STAGES * TILE_BYTES);   // full[s] at +8s, empty[s] at +8(STAGES+s)
    int *issued = reinterpret_cast<int *>(smem + STAGES * TILE_BYTES + 8 * 2 * STAGES);   // tiles issued so far
    // top-k: heap [k][QPB] u64, then candidate buffer [CAND_CAP][QPB] u16.   ball: list [k][QPB] u32
    unsigned long long *heap_all = reinterpret_cast<unsigned long long *>(smem + STAGES * TILE_BYTES + BAR_BYTES);
    unsigned short *cand_all = reinterpret_cast<unsigned short *>(heap_all + (size_t)(P.k + 1) * QPB);
    int *list_all = reinterpret_cast<int *>(heap_all);

    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y, split = blockIdx.z;
    const int tile0 = split * P.tiles_per_split;
    const int tile1 = min(tile0 + P.tiles_per_split, P.n_pad / TILE);
    const int ntiles = tile1 - tile0;
    const int k = P.k;
    constexpr int REC = Lay<FORM>::REC, HREC = Lay<FORM>::HREC, TCOPY = Lay<FORM>::TILE_COPY;
    const char *src = reinterpret_cast<const char *>(P.packed) + ((size_t)b * (P.n_pad / TILE) + (size_t)tile0) * TCOPY;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (STAGES + s), NCW);
        }
        mbar_fence_init();
        const int pre = ntiles < STAGES ? ntiles : STAGES;
        for (int t = 0; t < pre; ++t) {
            mbar_expect_tx(bar_base + 8 * t, TCOPY);
            bulk_g2s(smem_u32(smem + t * TILE_BYTES), src + (size_t)t * TCOPY, TCOPY, bar_base + 8 * t);
        }
        *issued = pre;
    }
    __syncthreads();

    const int ct = threadIdx.x;  // 0 .. NCT-1
    QueryConst qc[Q];
    float tau[Q];
    int cnt[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int qi = blockIdx.x * QPB + j * NCT + ct;
        float x = 0.f, y = 0.f, z = 0.f;
        if (qi < P.S) {
            const float *qp = P.qry + ((size_t)b * P.S + qi) * 3;
            x = qp[0]; y = qp[1]; z = qp[2];
        }
        qc[j] = make_query<FORM>(x, y, z);
        cnt[j] = 0;
        if (MODE == MODE_TOPK) {
            tau[j] = P.debug_nodrain ? -CUDART_INF_F : CUDART_INF_F;
            for (int e = 0; e < k; ++e) heap_all[e * QPB + j * NCT + ct] = HEAP_SENTINEL;
            heap_all[k * QPB + j * NCT + ct] = 0ull;   // pad: the smallest key, never selected as a child
        } else {
            tau[j] = P.r2;
        }
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(bar_base + 8 * s, (t / STAGES) & 1);
        const float4 *tp = tiles + (size_t)s * TILE;
        const int tile_ref0 = (tile0 + t) * TILE;

        int c = 0;
        while (c < CHUNKS_PER_TILE) {
            // warm-up (first tile of the split only): drain after 2,2,4,8,16,32 chunks so that tau
            // tightens quickly; afterwards once per tile (64 chunks, two 32-bit hit masks).
            int nch = CHUNKS_PER_TILE;
            if (t == 0) nch = c < 2 ? 2 : (c < 32 ? c : 32);
            if (nch > CHUNKS_PER_TILE - c) nch = CHUNKS_PER_TILE - c;

            // ---- hot loop: distance + half a min per pair, one compare per chunk, no branches ----
            float thr[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                float lo, hi;
                unpack2(qc[j].a3, lo, hi);
                thr[j] = prefilter_threshold<FORM>(tau[j], lo);
            }
            uint32_t mask[Q][2];
#pragma unroll
            for (int j = 0; j < Q; ++j) { mask[j][0] = 0u; mask[j][1] = 0u; }
            const float4 *cp = tp + c * REC;
            // register double buffer: records are requested ahead of their use and a warp-level memory barrier
            // pins those loads above the math of the current group, so the shared-memory latency is covered
            // (without it ptxas sinks every LDS next to its first use).  Q=2: a whole chunk ahead; Q=1: half a
            // chunk ahead, which keeps the kernel under 72 registers so that twice as many warps stay resident.
            if (Q >= 2) {
                float4 R[REC];
#pragma unroll
                for (int p = 0; p < REC; ++p) R[p] = cp[p];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int n_here = half == 0 ? (nch < 32 ? nch : 32) : nch - 32;
                    uint32_t bit = 1u;
                    for (int cc = 0; cc < n_here; ++cc, bit <<= 1) {
                        cp += REC;     // one chunk past the tile end is still inside the ring / barrier block: harmless
                        float4 Nx[REC];
#pragma unroll
                        for (int p = 0; p < REC; ++p) Nx[p] = cp[p];
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            const float m = chunk_min<FORM>(R, qc[j]);
                            const bool hit = MODE == MODE_TOPK ? (m < thr[j]) : (m <= tau[j]);
                            if (hit) mask[j][half] |= bit;
                        }
#pragma unroll
                        for (int p = 0; p < REC; ++p) R[p] = Nx[p];
                    }
                }
            } else {
                float4 H[HREC];
#pragma unroll
                for (int p = 0; p < HREC; ++p) H[p] = cp[p];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int n_here = half == 0 ? (nch < 32 ? nch : 32) : nch - 32;
                    uint32_t bit = 1u;
                    for (int cc = 0; cc < n_here; ++cc, bit <<= 1) {
                        float4 N1[HREC];
#pragma unroll
                        for (int p = 0; p < HREC; ++p) N1[p] = cp[HREC + p];
                        __syncwarp();
                        float ma[Q];
#pragma unroll
                        for (int j = 0; j < Q; ++j) ma[j] = half_chunk_min<FORM>(H, qc[j]);
                        cp += REC;
                        float4 N2[HREC];
#pragma unroll
                        for (int p = 0; p < HREC; ++p) N2[p] = cp[p];
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            const float m = fminf(ma[j], half_chunk_min<FORM>(N1, qc[j]));
                            const bool hit = MODE == MODE_TOPK ? (m < thr[j]) : (m <= tau[j]);
                            if (hit) mask[j][half] |= bit;
                        }
#pragma unroll
                        for (int p = 0; p < HREC; ++p) H[p] = N2[p];
                    }
                }
            }

            // ---- drain, warp-synchronous so that the lanes' slow work overlaps instead of serialising ----
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                uint32_t m0 = mask[j][0], m1 = mask[j][1];
                const int slot = j * NCT + ct;
                if (MODE == MODE_TOPK) {
                    const uint32_t hb = smem_u32(heap_all + slot), SB = (uint32_t)QPB * 8u;
                    unsigned short *cand = cand_all + slot;
                    while (__any_sync(FULL, (m0 | m1) != 0u)) {
                        // phase 1: every lane revisits its r-th hit chunk; the chunk's candidates (d < stale tau)
                        // are only recorded, as ONE entry (chunk << 8 | 8-bit mask), at most CAND_CAP per pass.
                        int nc = 0;
                        while (__any_sync(FULL, (m0 | m1) != 0u && nc < CAND_CAP)) {
                            if ((m0 | m1) != 0u && nc < CAND_CAP) {
                                int cc;
                                if (m0 != 0u) { cc = __ffs(m0) - 1; m0 &= m0 - 1; }
                                else { cc = 32 + __ffs(m1) - 1; m1 &= m1 - 1; }
                                const float4 *dp = tp + (c + cc) * REC;
                                float d[8];
#pragma unroll
                                for (int p = 0; p < 4; ++p) unpack2(chunk_pair_dist<FORM>(dp, c + cc, p, qc[j]), d[2 * p], d[2 * p + 1]);
                                uint32_t cm = 0u;
#pragma unroll
                                for (int i = 0; i < 8; ++i) cm |= (d[i] < tau[j]) ? (1u << i) : 0u;
                                if (cm != 0u) { cand[nc * QPB] = (unsigned short)((cc << 8) | cm); ++nc; }
                            }
                        }
                        // phase 2: all lanes insert their next candidate together (index order is preserved)
                        uint32_t cur = 0u;
                        int e = 0, base = 0;
                        while (true) {
                            if (cur == 0u && e < nc) {
                                const uint32_t v = cand[e * QPB];
                                ++e;
                                cur = v & 0xffu;
                                base = (c + (int)(v >> 8)) * CHUNK;
                            }
                            if (!__any_sync(FULL, cur != 0u)) break;
                            if (cur != 0u) {
                                const int off = base + __ffs(cur) - 1;
                                cur &= cur - 1;
                                const float d = tile_dist<FORM>(tp, off, qc[j]);
                                if (d < tau[j]) {
                                    heap_sift_root<true>(hb, SB, (uint32_t)k * SB, ((unsigned long long)order_key(d) << 32) | (uint32_t)(tile_ref0 + off));
                                    tau[j] = key_to_float((uint32_t)(lds_u64(hb) >> 32));
                                }
                            }
                        }
                    }
                } else {
                    int *list = list_all + slot;
                    while (__any_sync(FULL, (m0 | m1) != 0u)) {
                        if ((m0 | m1) != 0u) {
                            int cc;
                            if (m0 != 0u) { cc = __ffs(m0) - 1; m0 &= m0 - 1; }
                            else { cc = 32 + __ffs(m1) - 1; m1 &= m1 - 1; }
                            const int off0 = (c + cc) * CHUNK;
                            const float4 *dp = tp + (c + cc) * REC;
                            float d[8];
#pragma unroll
                            for (int p = 0; p < 4; ++p) unpack2(chunk_pair_dist<FORM>(dp, c + cc, p, qc[j]), d[2 * p], d[2 * p + 1]);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (d[i] <= tau[j] && cnt[j] < k) { list[cnt[j] * QPB] = tile_ref0 + off0 + i; ++cnt[j]; }
                            if (cnt[j] == k) { tau[j] = -CUDART_INF_F; m0 = 0u; m1 = 0u; }   // this query is complete
                        }
                    }
                }
            }
            c += nch;
        }

        // ---- release the stage; the warp that completes the release refills it ----
        __syncwarp();
        if (lane == 0) {
            const uint32_t ebar = bar_base + 8 * (STAGES + s);
            mbar_arrive(ebar);
            const int nxt = t + STAGES;                 // the tile that will reuse this stage
            if (nxt < ntiles && mbar_test(ebar, (t / STAGES) & 1)) {
                if (atomicCAS(issued, nxt, nxt + 1) == nxt) {
                    mbar_expect_tx(bar_base + 8 * s, TCOPY);
                    bulk_g2s(smem_u32(smem + s * TILE_BYTES), src + (size_t)nxt * TCOPY, TCOPY, bar_base + 8 * s);
                }
            }
        }
    }

    // ---------------- results ----------------
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const int qi = blockIdx.x * QPB + j * NCT + ct;
        const int slot = j * NCT + ct;
        if (MODE == MODE_TOPK) {
            // in-place heapsort: ascending (distance, index) order
            const uint32_t hb = smem_u32(heap_all + slot), SB = (uint32_t)QPB * 8u;
            for (int n = k; n > 1; --n) {
                const unsigned long long top = lds_u64(hb);
                const unsigned long long last = lds_u64(hb + (uint32_t)(n - 1) * SB);
                heap_sift_root<false>(hb, SB, (uint32_t)(n - 1) * SB, last);
                sts_u64(hb + (uint32_t)(n - 1) * SB, top);
            }
        }
        if (qi >= P.S) continue;
        const size_t row = (size_t)b * P.S + qi;
        if (P.n_split == 1) {
            int64_t *io = P.idx_out ? P.idx_out + row * k : nullptr;
            if (MODE == MODE_TOPK) {
                float *dout = P.dist_out ? P.dist_out + row * k : nullptr;
                for (int e = 0; e < k; ++e) {
                    const unsigned long long key = heap_all[e * QPB + slot];
                    if (io) io[e] = (int64_t)(uint32_t)key;
                    if (dout) dout[e] = key_to_float((uint32_t)(key >> 32));
                }
            } else {
                const int *list = list_all + slot;
                const int first = cnt[j] > 0 ? list[0] : P.N;
                for (int e = 0; e < k; ++e) io[e] = e < cnt[j] ? list[e * QPB] : first;
            }
        } else {
            const size_t prow = (row * P.n_split + split) * k;
            if (MODE == MODE_TOPK) {
                for (int e = 0; e < k; ++e) {
                    const unsigned long long key = heap_all[e * QPB + slot];
                    P.part_d[prow + e] = key_to_float((uint32_t)(key >> 32));
                    P.part_i[prow + e] = (int)(uint32_t)key;
                }
            } else {
                const int *list = list_all + slot;
                P.part_cnt[row * P.n_split + split] = cnt[j];
                for (int e = 0; e < cnt[j]; ++e) P.part_i[prow + e] = list[e * QPB];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 4. merge of partial lists (only when the ref range was split)
// ---------------------------------------------------------------------------------------------
__global__ void merge_topk_kernel(const float *__restrict__ part_d, const int *__restrict__ part_i, int rows,
                                  int n_split, int k, int64_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    unsigned short head[MAX_SPLIT];
#pragma unroll
    for (int s = 0; s < MAX_SPLIT; ++s) head[s] = 0;
    const float *pd = part_d + (size_t)row * n_split * k;
    const int *pi = part_i + (size_t)row * n_split * k;
    for (int o = 0; o < k; ++o) {
        float best = CUDART_INF_F;
        int bs = 0;
        // splits hold disjoint, increasing index ranges: strict '<' keeps the lower index on ties
        for (int s = 0; s < n_split; ++s) {
            const int h = head[s];
            const float d = h < k ? pd[s * k + h] : CUDART_INF_F;
            if (d < best) { best = d; bs = s; }
        }
        const int h = head[bs];
        if (idx_out) idx_out[(size_t)row * k + o] = pi[bs * k + h];
        if (dist_out) dist_out[(size_t)row * k + o] = best;
        head[bs] = h + 1;
    }
}

__global__ void merge_ball_kernel(const int *__restrict__ part_i, const int *__restrict__ part_cnt, int rows,
                                  int n_split, int k, int N, int64_t *__restrict__ idx_out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    int64_t *o = idx_out + (size_t)row * k;
    int n = 0;
    for (int s = 0; s < n_split && n < k; ++s) {
        const int c = part_cnt[(size_t)row * n_split + s];
        const int *pi = part_i + ((size_t)row * n_split + s) * k;
        for (int e = 0; e < c && n < k; ++e) o[n++] = pi[e];
    }
    const int64_t first = n > 0 ? o[0] : (int64_t)N;
    for (; n < k; ++n) o[n] = first;
}

// ---------------------------------------------------------------------------------------------
// 5. planning + launch
// ---------------------------------------------------------------------------------------------
// per-query shared-memory bytes.  top-k: u64 heap [k] + u16 candidate buffer [CAND_CAP];  ball: u32 list [nsample]
static size_t query_bytes(int k, int mode) { return mode == MODE_TOPK ? (size_t)(k + 1) * 8 + (size_t)CAND_CAP * 2 : (size_t)k * 4; }
static const size_t kMaxSmem = 227 * 1024;
static const size_t kFixedSmem = (size_t)STAGES * TILE_BYTES + BAR_BYTES;

// Choose {queries per thread, consumer warps, ref split} so that the grid is (close to) a whole
// number of waves of resident CTAs: CTA count = B * ceil(S / q_per_block) * n_split against
// slots = SMs * CTAs-per-SM.  Among the candidates the one with the best wave efficiency wins;
// ties go to more resident warps, then to Q=2 (half the shared-memory loads per pair).
bool plan_search(int B, int N, int S, int k, int mode, SearchPlan *pl) {
    const int sms = sm_count();
    const size_t qb = query_bytes(k, mode);
    if (kFixedSmem + 32 * qb > kMaxSmem) return false;   // not even one warp of queries fits
    pl->n_pad = (int)align_up((size_t)N, TILE);
    pl->n_tiles = pl->n_pad / TILE;

    int force_q = 0, force_w = 0, force_split = 0;   // tuning / debugging overrides (not part of the ABI)
    if (const char *e = getenv("B200PC_FORCE_Q")) force_q = atoi(e);
    if (const char *e = getenv("B200PC_FORCE_WARPS")) force_w = atoi(e);
    if (const char *e = getenv("B200PC_FORCE_SPLIT")) force_split = atoi(e);

    double best_score = -1.0;
    int best_q = 1, best_w = 1, best_split = 1;
    const double lnk = 1.0 + log(fmax(1.0, (double)N / k));
    for (int q = 2; q >= 1; --q) {
        if (force_q && q != force_q) continue;
        for (int c = 1; c <= 8; ++c) {                       // target CTAs per SM
            const size_t budget = kMaxSmem / c - 1024;          // ~1 KB per resident CTA is reserved by the system
            if (budget <= kFixedSmem + 32 * q * qb) continue;
            int wmax = (int)((budget - kFixedSmem) / (32 * q * qb));
            if (wmax > MAX_WARPS) wmax = MAX_WARPS;
            const long slots = (long)sms * c;
            for (int split = 1; split <= MAX_SPLIT && split <= pl->n_tiles; split = split < 4 ? split + 1 : split * 2) {
                if (force_split && split != (force_split > pl->n_tiles ? pl->n_tiles : force_split)) continue;
                // queries per CTA if the work items (query block x ref split) of each batch item fill the slots
                long ipb = slots / ((long)B * split);           // query blocks per batch item, one wave
                if (ipb < 1) ipb = 1;
                int w = (int)((((long)S + ipb - 1) / ipb + 32 * q - 1) / (32 * q));
                if (w < 1) w = 1;
                if (w > wmax) w = wmax;
                if (force_w) w = force_w > wmax ? wmax : force_w;
                // register file: 64K registers per SM, ~125 (Q=2) / <=72 (Q=1) per thread
                if ((long)w * 32 * c * (q == 2 ? 128 : 72) > 65536) continue;
                const long items = (long)B * (((long)S + 32L * q * w - 1) / (32L * q * w));
                const int tps = (pl->n_tiles + split - 1) / split;
                const int real_split = (pl->n_tiles + tps - 1) / tps;
                const long ctas = items * real_split;
                const long waves = (ctas + slots - 1) / slots;
                const double wave_eff = (double)ctas / (double)(waves * slots);
                const double pad_eff = (double)B * S / ((double)items * 32 * q * w);
                // resident warps per SM: below ~16 the SM cannot hide the drain's latency (and below 4 it idles)
                const double resident = fmin((double)c, (double)ctas / sms) * w;
                const double occ = pow(fmin(1.0, resident / 16.0), 0.8);
                // a split repeats the warm-up of the k-best list in every ref range: ~k(1+ln(n/k)) candidates each
                const double drain = real_split * (1.0 + log(fmax(1.0, (double)N / real_split / k))) / lnk;
                const double work = 0.5 + 0.5 * drain + (real_split > 1 ? 0.05 : 0.0);
                const double score = wave_eff * pad_eff * occ / work + 1e-3 * (q == 2) + 1e-5 * resident;
                if (score > best_score) { best_score = score; best_q = q; best_w = w; best_split = real_split; }
            }
        }
    }
    pl->q_per_thread = best_q;
    pl->consumer_warps = best_w;
    pl->q_per_block = best_q * best_w * 32;
    pl->tiles_per_split = (pl->n_tiles + best_split - 1) / best_split;
    pl->n_split = (pl->n_tiles + pl->tiles_per_split - 1) / pl->tiles_per_split;
    pl->smem_bytes = kFixedSmem + (size_t)pl->q_per_block * qb;
    pl->packed_bytes = align_up((size_t)B * pl->n_pad * 16, 256);
    pl->part_bytes = 0;
    if (pl->n_split > 1) {
        const size_t rows = (size_t)B * S * pl->n_split;
        pl->part_bytes = align_up(rows * k * 4, 256) * (mode == MODE_TOPK ? 2 : 1) + align_up(rows * 4, 256);
    }
    pl->total_bytes = pl->packed_bytes + pl->part_bytes;
    return true;
}

template <int FORM, int MODE, int Q>
static int launch_one(const SearchArgs &a, const SearchPlan &pl, int B, cudaStream_t st) {
    auto kern = search_kernel<FORM, MODE, Q>;
    B200PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    dim3 grid((a.S + pl.q_per_block - 1) / pl.q_per_block, B, pl.n_split);
    kern<<<grid, pl.consumer_warps * 32, pl.smem_bytes, st>>>(a);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}

template <int FORM, int MODE>
static int launch_shape(const SearchArgs &a, const SearchPlan &pl, int B, cudaStream_t st) {
    if (pl.q_per_thread == 2) return launch_one<FORM, MODE, 2>(a, pl, B, st);
    return launch_one<FORM, MODE, 1>(a, pl, B, st);
}

static int run_search(const float *ref, const float *qry, int B, int N, int S, int k, int form, int mode, float r2,
                      int64_t *idx, float *dist, void *ws, size_t ws_bytes, cudaStream_t st) {
    B200PC_REQUIRE(ref && qry, "search: null input pointer");
    B200PC_REQUIRE(B >= 0 && N >= 1 && S >= 0 && k >= 1, "search: bad sizes B=%d N=%d S=%d k=%d", B, N, S, k);
    B200PC_REQUIRE(idx || dist, "search: no output requested");
    if (B == 0 || S == 0) return B200PC_OK;
    {   // small reference sets: warp-per-query kernel, one launch, no workspace (small_search.cu)
        const int rc_small = run_small(ref, qry, B, N, S, k, form, mode, r2, idx, dist, st);
        if (rc_small != -100) return rc_small;
    }
    SearchPlan pl;
    if (!plan_search(B, N, S, k, mode, &pl)) {
        set_error("search: list length k=%d does not fit in shared memory", k);
        return B200PC_EINVAL;
    }
    if (!ws || ws_bytes < pl.total_bytes) {
        set_error("search: workspace too small (%zu < %zu bytes)", ws_bytes, pl.total_bytes);
        return B200PC_EWORKSPACE;
    }
    B200PC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "search: workspace must be 16-byte aligned");

    char *w = static_cast<char *>(ws);
    float4 *packed = reinterpret_cast<float4 *>(w);
    {
        dim3 grid((pl.n_pad / 2 + 255) / 256, B);
        pack_refs_kernel<<<grid, 256, 0, st>>>(ref, N, pl.n_pad, form, packed);
        B200PC_LAUNCH_CHECK();
    }
    SearchArgs a;
    a.packed = packed; a.qry = qry; a.N = N; a.n_pad = pl.n_pad; a.S = S; a.k = k; a.r2 = r2;
    a.n_split = pl.n_split; a.tiles_per_split = pl.tiles_per_split;
    a.debug_nodrain = getenv("B200PC_DEBUG_NODRAIN") != nullptr;
    a.idx_out = idx; a.dist_out = dist; a.part_d = nullptr; a.part_i = nullptr; a.part_cnt = nullptr;
    if (pl.n_split > 1) {
        const size_t rows = (size_t)B * S * pl.n_split;
        char *p = w + pl.packed_bytes;
        a.part_i = reinterpret_cast<int *>(p); p += align_up(rows * k * 4, 256);
        if (mode == MODE_TOPK) { a.part_d = reinterpret_cast<float *>(p); p += align_up(rows * k * 4, 256); }
        a.part_cnt = reinterpret_cast<int *>(p);
    }
    int rc;
    if (mode == MODE_BALL) rc = launch_shape<B200PC_FORM_QRY_NORM_FIRST, MODE_BALL>(a, pl, B, st);
    else if (form == B200PC_FORM_REF_NORM_FIRST) rc = launch_shape<B200PC_FORM_REF_NORM_FIRST, MODE_TOPK>(a, pl, B, st);
    else if (form == B200PC_FORM_QRY_NORM_FIRST) rc = launch_shape<B200PC_FORM_QRY_NORM_FIRST, MODE_TOPK>(a, pl, B, st);
    else rc = launch_shape<B200PC_FORM_DIRECT, MODE_TOPK>(a, pl, B, st);
    if (rc != B200PC_OK) return rc;

    if (pl.n_split > 1) {
        const int rows = B * S;
        if (mode == MODE_TOPK)
            merge_topk_kernel<<<(rows + 127) / 128, 128, 0, st>>>(a.part_d, a.part_i, rows, pl.n_split, k, idx, dist);
        else
            merge_ball_kernel<<<(rows + 127) / 128, 128, 0, st>>>(a.part_i, a.part_cnt, rows, pl.n_split, k, N, idx);
        B200PC_LAUNCH_CHECK();
    }
    return B200PC_OK;
}

int run_topk(const float *ref, const float *qry, int B, int N, int S, int k, int form, int64_t *idx, float *dist,
             void *ws, size_t ws_bytes, cudaStream_t st) {
    B200PC_REQUIRE(form >= 0 && form <= 2, "knn: unknown distance form %d", form);
    B200PC_REQUIRE(k <= N, "knn: k=%d exceeds the number of reference points N=%d", k, N);
    return run_search(ref, qry, B, N, S, k, form, MODE_TOPK, 0.f, idx, dist, ws, ws_bytes, st);
}

int run_ball(const float *ref, const float *qry, int B, int N, int S, float r2, int nsample, int64_t *idx,
             void *ws, size_t ws_bytes, cudaStream_t st) {
    B200PC_REQUIRE(idx, "ball_query: null output pointer");
    return run_search(ref, qry, B, N, S, nsample, B200PC_FORM_QRY_NORM_FIRST, MODE_BALL, r2, idx, nullptr, ws,
                      ws_bytes, st);
}

}  // namespace b200pc

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace b200pc;

extern "C" size_t b200pc_search_workspace_bytes(int B, int N, int S, int k) {
    if (B <= 0 || N <= 0 || S <= 0 || k <= 0) return 256;
    SearchPlan a, c;
    size_t need = 256;
    if (plan_search(B, N, S, k, MODE_TOPK, &a)) need = a.total_bytes > need ? a.total_bytes : need;
    if (plan_search(B, N, S, k, MODE_BALL, &c)) need = c.total_bytes > need ? c.total_bytes : need;
    return need;
}

extern "C" int b200pc_knn(const float *ref, const float *qry, int B, int N, int S, int k, int form, int64_t *idx,
                          float *dist, void *workspace, size_t workspace_bytes, b200pc_stream_t stream) {
    return run_topk(ref, qry, B, N, S, k, form, idx, dist, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int b200pc_ball_query(const float *xyz, const float *new_xyz, int B, int N, int S, float r2, int nsample,
                                 int64_t *idx, void *workspace, size_t workspace_bytes, b200pc_stream_t stream) {
    B200PC_REQUIRE(nsample >= 1, "ball_query: nsample must be >= 1");
    return run_ball(xyz, new_xyz, B, N, S, r2, nsample, idx, workspace, workspace_bytes, as_stream(stream));
}
