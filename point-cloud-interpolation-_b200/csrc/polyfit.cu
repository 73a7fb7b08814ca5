// polyfit.cu -- PolyPCI's per-point polynomial fit + evaluation as one weighted sum on the device.
//
// Reference: PolyPCI/Models/Models_V1.py:116-124 (fitting_and_predict) and its call site :191-219 -- for every batch item
// and every coordinate the stacked frames [F,N] go to the HOST, np.polyfit fits a degree-d polynomial through the F
// time stamps independently for each of the N points, the polynomial is evaluated at t, and the result goes back to
// the device.  Least squares is linear in the data, so  value[n] = sum_f w[f] * y[f,n]  with ONE weight vector per
// batch item, w = V(t) . polyfit(T, I_F) (computed by b200pc.polypci.poly_weights with the reference's own numpy
// calls, in float64).  This kernel applies it: float64 accumulation like numpy's, one rounding to fp32 at the end like
// the reference's torch.tensor(...).to(torch.float32).  HBM-bound: 4 (F + 1) bytes per output element.
#include "common.cuh"

namespace b200pc {

constexpr int POLY_MAX_FRAMES = 16;
struct FramePtrs { const float *p[POLY_MAX_FRAMES]; };

// VEC = 4: one thread produces 4 consecutive outputs from 16-byte streaming loads (per_batch % 4 == 0, aligned pointers)
template <int VEC>
__global__ void __launch_bounds__(256) poly_predict_kernel(const FramePtrs fr, const double *__restrict__ w, int F, long per_batch,
                                                           long total, float *__restrict__ out) {
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;     // first element of this thread
    if (i >= total) return;
    const double *wb = w + (i / per_batch) * F;                            // VEC consecutive elements share a batch item
    double acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
    for (int f = 0; f < F; ++f) {
        const double wf = __ldg(wb + f);
        if (VEC == 4) {
            const float4 y = ldg_stream(reinterpret_cast<const float4 *>(fr.p[f] + i));
            acc[0] = fma(wf, (double)y.x, acc[0]); acc[1] = fma(wf, (double)y.y, acc[1]);
            acc[2] = fma(wf, (double)y.z, acc[2]); acc[3] = fma(wf, (double)y.w, acc[3]);
        } else {
            acc[0] = fma(wf, (double)__ldcs(fr.p[f] + i), acc[0]);
        }
    }
    if (VEC == 4) stg_stream(reinterpret_cast<float4 *>(out + i), make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]));
    else out[i] = (float)acc[0];
}

}  // namespace b200pc

using namespace b200pc;

extern "C" int b200pc_poly_predict(const float *const *frames, const double *weights, int B, int F, int64_t per_batch, float *out,
                                   b200pc_stream_t stream) {
    B200PC_REQUIRE(F >= 1 && F <= POLY_MAX_FRAMES, "poly_predict: F=%d frames, supported 1..%d", F, POLY_MAX_FRAMES);
    B200PC_REQUIRE(B >= 0 && per_batch >= 0, "poly_predict: bad sizes");
    const long total = (long)B * per_batch;
    if (total == 0) return B200PC_OK;
    B200PC_REQUIRE(frames && weights && out, "poly_predict: null pointer");
    FramePtrs fp;
    for (int f = 0; f < POLY_MAX_FRAMES; ++f) fp.p[f] = f < F ? frames[f] : nullptr;
    for (int f = 0; f < F; ++f) B200PC_REQUIRE(fp.p[f], "poly_predict: frame %d is null", f);
    bool vec = per_batch % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int f = 0; f < F; ++f) vec = vec && (reinterpret_cast<uintptr_t>(fp.p[f]) & 15) == 0;
    const long threads = vec ? total / 4 : total, blocks = (threads + 255) / 256;
    B200PC_REQUIRE(blocks < (1L << 31), "poly_predict: problem too large for one launch");
    if (vec) poly_predict_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(fp, weights, F, per_batch, total, out);
    else poly_predict_kernel<1><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(fp, weights, F, per_batch, total, out);
    B200PC_LAUNCH_CHECK();
    return B200PC_OK;
}
