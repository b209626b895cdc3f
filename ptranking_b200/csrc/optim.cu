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

// ---- Adagrad / RMSprop: the other two optimizers config_optimizer offers (ranker.py:517-520; Adagrad is the listsf
// default, ltr_adhoc/eval/parameter.py:157-162).  One state buffer each, same flat layout.  Operation order follows
// torch's _multi_tensor_adagrad / _multi_tensor_rmsprop (PyTorch defaults: Adagrad lr_decay=0, eps=1e-10,
// initial_accumulator_value=0; RMSprop alpha=0.99, eps=1e-8, momentum=0, centered=False):
//   Adagrad: g += wd*p; sum += g*g;                    p += -clr * g / (sqrt(sum) + eps)
//   RMSprop: g += wd*p; sq = alpha*sq + (1-alpha)*g*g; p += -lr  * g / (sqrt(sq)  + eps)
struct AccCfg { float lr, decay, one_minus_decay, eps, weight_decay; };

template <bool RMS>
static __device__ __forceinline__ void acc1(float& p, float g, float& st, const AccCfg& c) {
    if (c.weight_decay != 0.0f) g = fmaf(c.weight_decay, p, g);
    if (RMS) st = fmaf(c.one_minus_decay * g, g, st * c.decay);
    else st = fmaf(g, g, st);
    const float denom = sqrtf(st) + c.eps;
    p = fmaf(-c.lr, g / denom, p);
}

template <bool RMS>
__global__ void accum_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ st, size_t n, AccCfg c) {
    const size_t n4 = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], S = reinterpret_cast<float4*>(st)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        acc1<RMS>(P.x, G.x, S.x, c); acc1<RMS>(P.y, G.y, S.y, c); acc1<RMS>(P.z, G.z, S.z, c); acc1<RMS>(P.w, G.w, S.w, c);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(st)[i] = S;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc1<RMS>(p[i], g[i], st[i], c);
}

static int check_flat(const char* who, const void* a, const void* b, const void* c, int64_t count) {
    if (!a || !b || !c || count <= 0) { set_error("%s: bad arguments", who); return PTRB200_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) {
        set_error("%s: buffers must be 16-byte aligned", who);
        return PTRB200_ERR_INVALID;
    }
    return PTRB200_OK;
}
static unsigned flat_blocks(int64_t count) {
    const size_t n4 = (size_t)count / 4 + 1;
    return (unsigned)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
}

}  // namespace ptrb200

using namespace ptrb200;

extern "C" int ptrb200_adagrad_step(float* param, const float* grad, float* state_sum, int64_t count,
                                    double lr, double lr_decay, double eps, double weight_decay, int step,
                                    ptrb200_stream_t stream) {
    int rc = check_flat("adagrad_step", param, grad, state_sum, count);
    if (rc) return rc;
    if (step < 1) { set_error("adagrad_step: step must be >= 1"); return PTRB200_ERR_INVALID; }
    const double clr = lr / (1.0 + (double)(step - 1) * lr_decay);
    AccCfg c{(float)clr, 1.0f, 0.0f, (float)eps, (float)weight_decay};
    PTRB200_LAUNCH_TAG("adagrad_step_kernel", accum_step_kernel<false>, flat_blocks(count), 256, 0, stream, param, grad, state_sum, (size_t)count, c);
    return check_launch("adagrad_step");
}

extern "C" int ptrb200_rmsprop_step(float* param, const float* grad, float* square_avg, int64_t count,
                                    double lr, double alpha, double eps, double weight_decay,
                                    ptrb200_stream_t stream) {
    int rc = check_flat("rmsprop_step", param, grad, square_avg, count);
    if (rc) return rc;
    AccCfg c{(float)lr, (float)alpha, (float)(1.0 - alpha), (float)eps, (float)weight_decay};
    PTRB200_LAUNCH_TAG("rmsprop_step_kernel", accum_step_kernel<true>, flat_blocks(count), 256, 0, stream, param, grad, square_avg, (size_t)count, c);
    return check_launch("rmsprop_step");
}

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
