// ffnet_act.cuh -- activation functions of get_AF (ptranking/base/utils.py:101-143) with derivatives.
#pragma once
#include "common.cuh"

namespace ptrb200 {

struct ActOut { float y, dy; };
static __device__ __forceinline__ ActOut activate(int af, float x) {
    ActOut r;
    switch (af) {
        case PTRB200_AF_RELU: r.y = fmaxf(x, 0.0f); r.dy = x > 0.0f ? 1.0f : 0.0f; break;
        case PTRB200_AF_GELU: {
            const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
            r.y = x * cdf;
            r.dy = cdf + x * 0.3989422804014327f * expf(-0.5f * x * x);
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
