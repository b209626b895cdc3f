// ffnet_act.cuh -- activation functions of get_AF (ptranking/base/utils.py:101-143) with derivatives.
#pragma once
#include "common.cuh"

namespace ptrb200 {

// ---- exact-erf GELU without erff -------------------------------------------------------------------------------------
// nn.GELU() (get_AF 'GE', base/utils.py:125) is x * Phi(x) with the exact normal CDF.  erff() costs ~35 instructions of
// branch-free coefficient selects per element, and the layer kernels are issue-bound on exactly this prologue.
// Phi is evaluated directly instead:   h(u) = 0.5 erfc(u / sqrt 2) = 2^(u R(u) - 1),  u = |x|,
// R a degree-8 polynomial fitted (weighted minimax, tools/fit_gelu.py) to log2(erfc(u / sqrt 2)) / u on [0, 5.75]:
// approximation error 1.5e-9 in Phi, i.e. far below fp32 rounding;  Phi(x) = x >= 0 ? 1 - h : h  (no cancellation on the
// negative side, where 0.5 * (1 + erf) loses relative accuracy).  One MUFU.EX2, 9 FFMA, ~14 instructions in all.
// Measured against float64 over [-8, 8]: max |GELU error| 3.9e-7, rms 4.9e-8 -- the figures a correctly rounded erff gives
// (4.5e-7 / 5.5e-8; both are dominated by the final x * Phi rounding).  Beyond |x| = 5.75, h < 4.5e-9 is flushed to 0.
static __device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
static __device__ __forceinline__ float normal_cdf(float x) {
    const float u = fabsf(x);
    const float uc = fminf(u, 5.75f);
    float r = 5.077291253e-07f;
    r = fmaf(r, uc, -9.483431188e-06f);
    r = fmaf(r, uc, 7.450972771e-05f);
    r = fmaf(r, uc, -2.828807353e-04f);
    r = fmaf(r, uc, 1.236852252e-05f);
    r = fmaf(r, uc, 6.933792046e-03f);
    r = fmaf(r, uc, -5.243624784e-02f);
    r = fmaf(r, uc, -4.592210540e-01f);
    r = fmaf(r, uc, -1.151104305e+00f);
    float h = ex2_approx(fmaf(uc, r, -1.0f));
    h = u > 5.75f ? 0.0f : h;
    return x >= 0.0f ? 1.0f - h : h;
}
// x * phi(x) = x exp(-x^2 / 2) / sqrt(2 pi), the second term of GELU'
static __device__ __forceinline__ float x_normal_pdf(float x) {
    return (x * 0.3989422804014327f) * ex2_approx(x * x * -0.7213475204444817f);
}

struct ActOut { float y, dy; };
static __device__ __forceinline__ ActOut activate(int af, float x) {
    ActOut r;
    switch (af) {
        case PTRB200_AF_RELU: r.y = fmaxf(x, 0.0f); r.dy = x > 0.0f ? 1.0f : 0.0f; break;
        case PTRB200_AF_GELU: {
            const float cdf = normal_cdf(x);
            r.y = x * cdf;
            r.dy = cdf + x_normal_pdf(x);
        } break;
        case PTRB200_AF_SIGM: { const float s = __fdividef(1.0f, 1.0f + expf(-x)); r.y = s; r.dy = s * (1.0f - s); } break;
        case PTRB200_AF_TANH: { const float t = tanhf(x); r.y = t; r.dy = 1.0f - t * t; } break;
        case PTRB200_AF_CELU:
        case PTRB200_AF_ELU: { const float e = expf(x); r.y = x > 0.0f ? x : e - 1.0f; r.dy = x > 0.0f ? 1.0f : e; } break;
        case PTRB200_AF_LRELU: r.y = x > 0.0f ? x : 0.01f * x; r.dy = x > 0.0f ? 1.0f : 0.01f; break;
        case PTRB200_AF_SELU: {
            const float sc = 1.0507009873554805f, al = 1.6732632423543772f, e = expf(x);
            r.y = sc * (x > 0.0f ? x : al * (e - 1.0f));
            r.dy = sc * (x > 0.0f ? 1.0f : al * e);
        } break;
        default: r.y = x; r.dy = 1.0f; break;
    }
    return r;
}


}  // namespace ptrb200
