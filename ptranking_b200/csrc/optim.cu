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

// ---- data-parallel: the gradient all-reduce (SUM) folded into the optimizer step, over NVLink peer memory -------------
// One process per GPU; every rank's flat gradient buffer (and a small flag pad) lives in memory the other ranks have
// mapped through CUDA IPC (ptrb200_peer_alloc / ptrb200_peer_open).  The step kernel of rank r
//   1. announces "my gradients of step e are complete" with a system-scope release store of e into flag[r] of every
//      rank's pad (CTA 0), and every CTA waits until the local pad shows e from every rank (acquire loads of local memory);
//   2. reads element i of all ranks' buffers straight over NVLink, adds them in rank order 0..W-1 (the same order on
//      every rank: replicas stay bit-identical) and applies the update.
// One launch replaces ncclAllReduce + the optimizer launch; 221 KB x W of peer reads for the default scorer.  The flags
// only grow (the step counter), so nothing is ever reset; the caller alternates between two gradient buffers so that
// a rank already writing step e+1 gradients can never touch what a slower rank is still reading for step e.
struct PeerDev {
    int world, rank;
    const float* grads[PTRB200_MAX_PEERS];
    uint32_t* flags[PTRB200_MAX_PEERS];
    uint32_t epoch;
    int* error;                 // set to 1 + (rank that never arrived) when the wait gives up; may be null
};

static __device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// true when every rank has announced step `epoch`; false after 4 s without one of them (a crashed peer must not hang the GPU)
static __device__ __forceinline__ bool peer_arrive_and_wait(const PeerDev& g) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < g.world) {
        const int r = threadIdx.x;
        if (blockIdx.x == 0) { __threadfence_system(); st_release_sys(g.flags[r] + g.rank, g.epoch); }
        const uint32_t* mine = g.flags[g.rank] + r;
        const unsigned long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(mine) - g.epoch) < 0) {
            if (global_ns() - t0 > 4000000000ull) { ok = 0; if (g.error) atomicExch(g.error, 1 + r); break; }
            __nanosleep(40);
        }
    }
    __syncthreads();
    return ok != 0;
}
template <bool PEER>
static __device__ __forceinline__ float4 grad4(const float* g, const PeerDev& pd, size_t i) {
    if (!PEER) return reinterpret_cast<const float4*>(g)[i];
    float4 s = __ldcg(reinterpret_cast<const float4*>(pd.grads[0]) + i);
    for (int r = 1; r < pd.world; ++r) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(pd.grads[r]) + i);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    return s;
}
template <bool PEER>
static __device__ __forceinline__ float grad1(const float* g, const PeerDev& pd, size_t i) {
    if (!PEER) return g[i];
    float s = __ldcg(pd.grads[0] + i);
    for (int r = 1; r < pd.world; ++r) s += __ldcg(pd.grads[r] + i);
    return s;
}

template <bool PEER>
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, size_t n, AdamCfg c, const PeerDev pd) {
    if (PEER && !peer_arrive_and_wait(pd)) return;
    const size_t n4 = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 G = grad4<PEER>(g, pd, i);
        adam1(P.x, G.x, M.x, V.x, c); adam1(P.y, G.y, M.y, V.y, c); adam1(P.z, G.z, M.z, V.z, c); adam1(P.w, G.w, M.w, V.w, c);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) adam1(p[i], grad1<PEER>(g, pd, i), m[i], v[i], c);
}

// out = sum over ranks of their buffers (the bare exchange: tests, and callers that keep their own optimizer)
__global__ void peer_sum_kernel(float* __restrict__ out, size_t n, const PeerDev pd) {
    if (!peer_arrive_and_wait(pd)) return;
    const size_t n4 = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) reinterpret_cast<float4*>(out)[i] = grad4<true>(nullptr, pd, i);
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = grad1<true>(nullptr, pd, i);
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

template <bool RMS, bool PEER>
__global__ void accum_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ st, size_t n, AccCfg c,
                                  const PeerDev pd) {
    if (PEER && !peer_arrive_and_wait(pd)) return;
    const size_t n4 = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], S = reinterpret_cast<float4*>(st)[i];
        const float4 G = grad4<PEER>(g, pd, i);
        acc1<RMS>(P.x, G.x, S.x, c); acc1<RMS>(P.y, G.y, S.y, c); acc1<RMS>(P.z, G.z, S.z, c); acc1<RMS>(P.w, G.w, S.w, c);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(st)[i] = S;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc1<RMS>(p[i], grad1<PEER>(g, pd, i), st[i], c);
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
// every CTA of a peer kernel spins until all ranks have arrived: the grid must be co-resident (<= 4 CTAs per SM here)
static unsigned peer_blocks(int64_t count) {
    const unsigned b = flat_blocks(count);
    return b < 592 ? b : 592;
}
static int make_peer(const char* who, const ptrb200_peer_group* grp, PeerDev& pd) {
    if (!grp || grp->world < 1 || grp->world > PTRB200_MAX_PEERS || grp->rank < 0 || grp->rank >= grp->world || grp->epoch == 0) {
        set_error("%s: bad peer group (world 1..%d, rank inside it, epoch >= 1)", who, PTRB200_MAX_PEERS);
        return PTRB200_ERR_INVALID;
    }
    pd.world = grp->world; pd.rank = grp->rank; pd.epoch = grp->epoch; pd.error = grp->error;
    for (int r = 0; r < grp->world; ++r) {
        if (!grp->grads[r] || !grp->flags[r] || (reinterpret_cast<uintptr_t>(grp->grads[r]) & 15)) {
            set_error("%s: rank %d's gradient buffer / flag pad missing or not 16-byte aligned", who, r);
            return PTRB200_ERR_INVALID;
        }
        pd.grads[r] = grp->grads[r]; pd.flags[r] = grp->flags[r];
    }
    return PTRB200_OK;
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
    PTRB200_LAUNCH_TAG("adagrad_step_kernel", (accum_step_kernel<false, false>), flat_blocks(count), 256, 0, stream, param, grad, state_sum, (size_t)count, c, PeerDev{});
    return check_launch("adagrad_step");
}

extern "C" int ptrb200_rmsprop_step(float* param, const float* grad, float* square_avg, int64_t count,
                                    double lr, double alpha, double eps, double weight_decay,
                                    ptrb200_stream_t stream) {
    int rc = check_flat("rmsprop_step", param, grad, square_avg, count);
    if (rc) return rc;
    AccCfg c{(float)lr, (float)alpha, (float)(1.0 - alpha), (float)eps, (float)weight_decay};
    PTRB200_LAUNCH_TAG("rmsprop_step_kernel", (accum_step_kernel<true, false>), flat_blocks(count), 256, 0, stream, param, grad, square_avg, (size_t)count, c, PeerDev{});
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
    PTRB200_LAUNCH_TAG("adam_step_kernel", adam_step_kernel<false>, blocks, 256, 0, stream, param, grad, exp_avg, exp_avg_sq, (size_t)count, c, PeerDev{});
    return check_launch("adam_step");
}

// ---- peer memory: allocation / mapping (CUDA IPC) and the fused exchange + step entry points ------------------------
extern "C" int ptrb200_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
    if (bytes <= 0 || !dev_ptr || !handle64) { set_error("peer_alloc: bad arguments"); return PTRB200_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the header promises 64-byte handles");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { if (p) cudaFree(p); set_error("peer_alloc: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return PTRB200_OK;
}
extern "C" int ptrb200_peer_open(const unsigned char* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) { set_error("peer_open: bad arguments"); return PTRB200_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("peer_open: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
    *dev_ptr = p;
    return PTRB200_OK;
}
extern "C" int ptrb200_peer_close(void* dev_ptr) {
    const cudaError_t e = dev_ptr ? cudaIpcCloseMemHandle(dev_ptr) : cudaSuccess;
    if (e != cudaSuccess) { set_error("peer_close: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
    return PTRB200_OK;
}
extern "C" int ptrb200_peer_free(void* dev_ptr) {
    const cudaError_t e = dev_ptr ? cudaFree(dev_ptr) : cudaSuccess;
    if (e != cudaSuccess) { set_error("peer_free: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
    return PTRB200_OK;
}

extern "C" int ptrb200_peer_allreduce_sum(const ptrb200_peer_group* grp, float* out, int64_t count, ptrb200_stream_t stream) {
    PeerDev pd{};
    int rc = make_peer("peer_allreduce_sum", grp, pd);
    if (rc) return rc;
    if (!out || count <= 0 || (reinterpret_cast<uintptr_t>(out) & 15)) { set_error("peer_allreduce_sum: bad arguments"); return PTRB200_ERR_INVALID; }
    PTRB200_LAUNCH_TAG("peer_sum_kernel", peer_sum_kernel, peer_blocks(count), 256, 0, stream, out, (size_t)count, pd);
    return check_launch("peer_allreduce_sum");
}

extern "C" int ptrb200_adam_step_peer(const ptrb200_peer_group* grp, float* param, float* exp_avg, float* exp_avg_sq, int64_t count,
                                      double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                                      ptrb200_stream_t stream) {
    PeerDev pd{};
    int rc = make_peer("adam_step_peer", grp, pd);
    if (rc) return rc;
    if ((rc = check_flat("adam_step_peer", param, exp_avg, exp_avg_sq, count))) return rc;
    if (step < 1) { set_error("adam_step_peer: step must be >= 1"); return PTRB200_ERR_INVALID; }
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamCfg c{(float)(lr / bc1), (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay, (float)sqrt(bc2)};
    PTRB200_LAUNCH_TAG("adam_step_peer_kernel", adam_step_kernel<true>, peer_blocks(count), 256, 0, stream, param, (const float*)nullptr, exp_avg, exp_avg_sq, (size_t)count, c, pd);
    return check_launch("adam_step_peer");
}

extern "C" int ptrb200_adagrad_step_peer(const ptrb200_peer_group* grp, float* param, float* state_sum, int64_t count,
                                         double lr, double lr_decay, double eps, double weight_decay, int step,
                                         ptrb200_stream_t stream) {
    PeerDev pd{};
    int rc = make_peer("adagrad_step_peer", grp, pd);
    if (rc) return rc;
    if ((rc = check_flat("adagrad_step_peer", param, state_sum, state_sum, count))) return rc;
    if (step < 1) { set_error("adagrad_step_peer: step must be >= 1"); return PTRB200_ERR_INVALID; }
    const double clr = lr / (1.0 + (double)(step - 1) * lr_decay);
    AccCfg c{(float)clr, 1.0f, 0.0f, (float)eps, (float)weight_decay};
    PTRB200_LAUNCH_TAG("adagrad_step_peer_kernel", (accum_step_kernel<false, true>), peer_blocks(count), 256, 0, stream, param, (const float*)nullptr, state_sum, (size_t)count, c, pd);
    return check_launch("adagrad_step_peer");
}

extern "C" int ptrb200_rmsprop_step_peer(const ptrb200_peer_group* grp, float* param, float* square_avg, int64_t count,
                                         double lr, double alpha, double eps, double weight_decay,
                                         ptrb200_stream_t stream) {
    PeerDev pd{};
    int rc = make_peer("rmsprop_step_peer", grp, pd);
    if (rc) return rc;
    if ((rc = check_flat("rmsprop_step_peer", param, square_avg, square_avg, count))) return rc;
    AccCfg c{(float)lr, (float)alpha, (float)(1.0 - alpha), (float)eps, (float)weight_decay};
    PTRB200_LAUNCH_TAG("rmsprop_step_peer_kernel", (accum_step_kernel<true, true>), peer_blocks(count), 256, 0, stream, param, (const float*)nullptr, square_avg, (size_t)count, c, pd);
    return check_launch("rmsprop_step_peer");
}
