// optim.cu -- the optimizer step of the training loop as ONE kernel over the flat parameter / gradient buffers.
//
// Reference: NeuralRanker.config_optimizer, ptranking/base/ranker.py:512-525 builds torch.optim.Adam(params, lr,
// weight_decay) (PyTorch defaults betas=(0.9, 0.999), eps=1e-8, amsgrad=False) and every loss class ends its train_op with
// optimizer.step() (e.g. ptranking/ltr_adhoc/listwise/lambdarank.py:58-60).  torch runs that step as ~8 multi-tensor
// launches over 18 small tensors; here parameters, gradients and both moment buffers are flat fp32 arrays with identical
// offsets (dist.GradBucket), so the whole update is one elementwise pass.  Operation order follows torch's
// _multi_tensor_adam: g += wd*p; m = lerp(m, g, 1-b1); v = v*b2 + (1-b2)*g*g; p += -(lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps).
#include "common.cuh"

namespace ptrb200 {

struct AdamCfg { float step_size, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, bc2_sqrt; };

static __device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamCfg& c) {
    if (c.weight_decay != 0.0f) g = fmaf(c.weight_decay, p, g);
    m = fmaf(c.one_minus_beta1, g - m, m);
    v = fmaf(c.one_minus_beta2 * g, g, v * c.beta2);
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
    p = fmaf(-c.step_size, m / denom, p);
}

__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, size_t n, AdamCfg c) {
    const size_t n4 = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        adam1(P.x, G.x, M.x, V.x, c); adam1(P.y, G.y, M.y, V.y, c); adam1(P.z, G.z, M.z, V.z, c); adam1(P.w, G.w, M.w, V.w, c);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) adam1(p[i], g[i], m[i], v[i], c);
}

}  // namespace ptrb200

using namespace ptrb200;

extern "C" int ptrb200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count,
                                 double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                                 ptrb200_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || count <= 0 || step < 1) { set_error("adam_step: bad arguments"); return PTRB200_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) {
        set_error("adam_step: buffers must be 16-byte aligned");
        return PTRB200_ERR_INVALID;
    }
    // hyper-parameters arrive as the Python doubles they are; every derived scalar is formed in double like the Python
    // reference does and rounded to fp32 once (1 - 0.999 in fp32 would already be off by 5e-5 relative)
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamCfg c{(float)(lr / bc1), (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay, (float)sqrt(bc2)};
    const size_t n4 = (size_t)count / 4 + 1;
    const unsigned blocks = (unsigned)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
    PTRB200_LAUNCH(adam_step_kernel, blocks, 256, 0, stream, param, grad, exp_avg, exp_avg_sq, (size_t)count, c);
    return check_launch("adam_step");
}
