// losses_ext.cu -- the sibling ranking losses of SURVEY 8f-4 on the same one-CTA-per-query skeleton, plus the Sinkhorn
// half-step (the reference's only CUDA kernel).
//
// Reference functions replaced (wildltr/ptranking @ f1d366c):
//   RankMSE     ptranking/ltr_adhoc/pointwise/rank_mse.py:13-22
//   RankCosine  ptranking/ltr_adhoc/listwise/rank_cosine.py:33 (nn.CosineSimilarity(dim=1), eps 1e-8)
//   STListNet   ptranking/ltr_adhoc/listwise/st_listnet.py:41-49
//   SoftRank    ptranking/ltr_adhoc/listwise/softrank.py:46-72
//   sinkstep    ptranking/ltr_adhoc/listwise/wassrank/pytorch_wasserstein.py:132-224 (CUDA string) / :277-291 (CPU form)
#include "losses_common.cuh"

namespace ptrb200 {

// ---------------------------------------------------------------------------
// RankMSE: mean over the batch of the per-query sum of squared errors.  loss_q[b] = sum_i (s-y)^2 / B so that the sum
// over queries is the reference's torch.mean(torch.sum(., dim=1)); grad = 2 (s - y) / B.
// ---------------------------------------------------------------------------
__global__ void rankmse_kernel(const float* __restrict__ scores, const float* __restrict__ labels, const int32_t* __restrict__ offsets,
                               float* __restrict__ grad, float* __restrict__ loss_q, int n_uniform, float inv_B) {
    __shared__ float red[33];
    const int b = blockIdx.x;
    const ListSpan sp = list_span(offsets, b, n_uniform);
    float acc = 0.0f;
    for (int i = threadIdx.x; i < sp.n; i += blockDim.x) {
        const float d = scores[sp.base + i] - labels[sp.base + i];
        acc = fmaf(d, d, acc);
        grad[sp.base + i] = (2.0f * d) * inv_B;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) loss_q[b] = acc * inv_B;
}

// ---------------------------------------------------------------------------
// RankCosine: (1 - cos(s, y)) / 0.5 per query with ATen's cosine_similarity: cos = sum_i (s_i / max(|s|, eps)) (y_i / max(|y|, eps)).
// d cos / d s_i = y_i / (S Y) - [|s| > eps] (s . y) s_i / (S^2 |s| Y),  S = max(|s|, eps), Y = max(|y|, eps).
// ---------------------------------------------------------------------------
__global__ void rankcosine_kernel(const float* __restrict__ scores, const float* __restrict__ labels, const int32_t* __restrict__ offsets,
                                  float* __restrict__ grad, float* __restrict__ loss_q, int n_uniform, float eps) {
    __shared__ float red[33];
    const int b = blockIdx.x;
    const ListSpan sp = list_span(offsets, b, n_uniform);
    float ss = 0.0f, yy = 0.0f, sy = 0.0f;
    for (int i = threadIdx.x; i < sp.n; i += blockDim.x) {
        const float s = scores[sp.base + i], y = labels[sp.base + i];
        ss = fmaf(s, s, ss); yy = fmaf(y, y, yy); sy = fmaf(s, y, sy);
    }
    ss = block_sum(ss, red); yy = block_sum(yy, red); sy = block_sum(sy, red);
    const float ns = sqrtf(ss), ny = sqrtf(yy);
    const float S = fmaxf(ns, eps), Y = fmaxf(ny, eps);
    const float cosv = sy / (S * Y);
    const float a = 1.0f / (S * Y);
    const float c = ns > eps ? sy / (S * S * ns * Y) : 0.0f;
    for (int i = threadIdx.x; i < sp.n; i += blockDim.x) {
        const float s = scores[sp.base + i], y = labels[sp.base + i];
        grad[sp.base + i] = -2.0f * (y * a - c * s);
    }
    if (threadIdx.x == 0) loss_q[b] = sp.n > 0 ? (1.0f - cosv) * 2.0f : 0.0f;
}

// ---------------------------------------------------------------------------
// STListNet: ListNet on Gumbel-perturbed scores z = (s + g) / T, g = -log(-log(u + 1e-20) + 1e-20), u ~ U[0,1).
// `unif` (optional) injects the uniforms (parity tests replay the reference's torch.rand draw); otherwise u comes from
// Philox4x32-10 keyed by (seed, offset, flat doc index) as 24-bit fractions, the granularity of torch.rand(float32).
// grad = (softmax(z) - softmax(y)) / T.
// ---------------------------------------------------------------------------
__global__ void stlistnet_kernel(const float* __restrict__ scores, const float* __restrict__ labels, const int32_t* __restrict__ offsets,
                                 const float* __restrict__ unif, float* __restrict__ grad, float* __restrict__ loss_q,
                                 int n_uniform, float inv_T, uint64_t seed, uint64_t offset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* zs = reinterpret_cast<float*>(smem_raw);
    float* ys = zs + n_uniform;
    float* red = ys + n_uniform;
    const int b = blockIdx.x;
    const ListSpan sp = list_span(offsets, b, n_uniform);
    const int n = sp.n;
    float mz = -INFINITY, my = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float u;
        if (unif) u = unif[sp.base + i];
        else u = (float)(dropout_bits(seed, offset, (uint64_t)(sp.base + i)) >> 8) * (1.0f / 16777216.0f);
        const float g = -logf(-logf(u + 1e-20f) + 1e-20f);
        const float z = (scores[sp.base + i] + g) * inv_T;
        const float y = labels[sp.base + i];
        zs[i] = z; ys[i] = y;
        mz = fmaxf(mz, z); my = fmaxf(my, y);
    }
    mz = block_max(mz, red);
    my = block_max(my, red);
    float sz = 0.0f, sy = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sz += expf(zs[i] - mz); sy += expf(ys[i] - my); }
    sz = block_sum(sz, red);
    sy = block_sum(sy, red);
    const float log_sz = logf(sz);
    float loss = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float py = expf(ys[i] - my) / sy;
        const float lsm = (zs[i] - mz) - log_sz;
        loss -= py * lsm;
        grad[sp.base + i] = (expf(lsm) - py) * inv_T;
    }
    loss = block_sum(loss, red);
    if (threadIdx.x == 0) loss_q[b] = n > 0 ? loss : 0.0f;
}

// ---------------------------------------------------------------------------
// SoftRank (metric nDCG): expected rank r_i = 1 + sum_{j != i} 0.5 erfc((s_i - s_j) / den), den = sqrt(2 * 2 delta^2);
// loss_q = - sum_{i < k} G(y_i) / (log2(r_i + 1) iDCG), iDCG over the labels as given (presort contract, softrank.py:40).
// c_i = d loss / d r_i = [i < k] G_i / (iDCG log2(r_i+1)^2 (r_i+1) ln 2); with e_ij = -exp(-x_ij^2) / (sqrt(pi) den)
// (= d r_i / d s_i contribution of j, symmetric in i,j):  grad_m = sum_{j != m} e_mj (c_m - c_j).
// ---------------------------------------------------------------------------
__global__ void softrank_kernel(const float* __restrict__ scores, const float* __restrict__ labels, const int32_t* __restrict__ offsets,
                                float* __restrict__ grad, float* __restrict__ loss_q, int n_uniform, float inv_den, int top_k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* ss = reinterpret_cast<float*>(smem_raw);
    float* cc = ss + n_uniform;
    float* red = cc + n_uniform;
    const int b = blockIdx.x;
    const ListSpan sp = list_span(offsets, b, n_uniform);
    const int n = sp.n;
    if (n == 0) { if (threadIdx.x == 0) loss_q[b] = 0.0f; return; }
    const float* y = labels + sp.base;
    float part = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { ss[i] = scores[sp.base + i]; part += gain_of(y[i]) / log2_rank(i); }
    const float idcg = block_sum(part, red);        // ends with a barrier: ss is visible
    const int K = (top_k > 0 && top_k < n) ? top_k : n;
    float dcg = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float si = ss[i];
        float r = 0.0f;
        for (int j = 0; j < n; ++j) r += (j != i) ? 0.5f * erfcf((si - ss[j]) * inv_den) : 0.0f;
        r += 1.0f;
        float c = 0.0f;
        if (i < K) {
            const float G = gain_of(y[i]);
            const float lg = log2f(r + 1.0f);
            dcg += (G / lg) / idcg;                  // (dist * gain) / idcg, the reference's association
            c = G / (idcg * lg * lg * (r + 1.0f) * 0.6931471805599453f);
        }
        cc[i] = c;
    }
    __syncthreads();
    const float ecoef = -0.5641895835477563f * inv_den;      // -1 / (sqrt(pi) den)
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        const float sm = ss[m], cm = cc[m];
        float acc = 0.0f;
        for (int j = 0; j < n; ++j) {
            const float x = (sm - ss[j]) * inv_den;
            acc = fmaf(expf(-x * x), cm - cc[j], acc);        // j == m contributes exactly 0
        }
        grad[sp.base + m] = ecoef * acc;
    }
    dcg = block_sum(dcg, red);
    if (threadIdx.x == 0) loss_q[b] = -dcg;
}

// ---------------------------------------------------------------------------
// Sinkhorn half-step:  log_v[b][j] = log_nu[b][j] - logsumexp_i( -dist[i][j] / lambda + log_u[b][i] ).
// One CTA per (tile of 32 columns j, b): lane = column (coalesced reads of dist rows), the 8 warps split the reduction
// rows i, each lane keeps an online (max, sum-exp) pair, the eight partials of a column are merged through shared
// memory.  Infinite entries follow the reference's CPU form (the one that runs): plain IEEE arithmetic.
// ---------------------------------------------------------------------------
constexpr int SINK_WARPS = 8;
__global__ void sinkstep_kernel(const float* __restrict__ dist, const float* __restrict__ log_nu, const float* __restrict__ log_u,
                                float* __restrict__ log_v, int d1, int d2, float lambda) {
    __shared__ float smax[SINK_WARPS][33], ssum[SINK_WARPS][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane, b = blockIdx.y;
    float mx = -INFINITY, se = 0.0f;
    if (j < d2) {
        for (int i = warp; i < d1; i += SINK_WARPS) {
            const float v = __fdiv_rn(-dist[(size_t)i * d2 + j], lambda) + log_u[(size_t)b * d1 + i];   // -dist/lambda + log_u, the reference's order
            if (v > mx) { se = se * expf(mx - v) + 1.0f; mx = v; }      // exp(-inf) = 0 on the first finite value
            else if (v > -INFINITY) se += expf(v - mx);
        }
    }
    smax[warp][lane] = mx; ssum[warp][lane] = se;
    __syncthreads();
    if (warp == 0 && j < d2) {
        float M = smax[0][lane];
#pragma unroll
        for (int w = 1; w < SINK_WARPS; ++w) M = fmaxf(M, smax[w][lane]);
        float S = 0.0f;
#pragma unroll
        for (int w = 0; w < SINK_WARPS; ++w) S += smax[w][lane] > -INFINITY ? ssum[w][lane] * expf(smax[w][lane] - M) : 0.0f;
        // pytorch_wasserstein.py:289 (the path that runs): log_nu - logsumexp(...), IEEE arithmetic on the infinities --
        // an all -inf column gives +inf (NaN when log_nu is -inf too); the never-compiled CUDA string would give -inf
        const float lse = M > -INFINITY ? logf(S) + M : -INFINITY;
        log_v[(size_t)b * d2 + j] = log_nu[(size_t)b * d2 + j] - lse;
    }
}

}  // namespace ptrb200

using namespace ptrb200;

extern "C" {

int ptrb200_rankmse_fwd_bwd(const float* scores, const float* labels, const int32_t* offsets, float* grad, float* loss_per_query,
                            int B, int n, ptrb200_stream_t stream) {
    int rc = check_list_args(scores, labels, grad, loss_per_query, B, n);
    if (rc) return rc;
    PTRB200_LAUNCH(rankmse_kernel, B, block_threads(n), 0, stream, scores, labels, offsets, grad, loss_per_query, n, 1.0f / (float)B);
    return check_launch("rankmse");
}

int ptrb200_rankcosine_fwd_bwd(const float* scores, const float* labels, const int32_t* offsets, float* grad, float* loss_per_query,
                               int B, int n, ptrb200_stream_t stream) {
    int rc = check_list_args(scores, labels, grad, loss_per_query, B, n);
    if (rc) return rc;
    PTRB200_LAUNCH(rankcosine_kernel, B, block_threads(n), 0, stream, scores, labels, offsets, grad, loss_per_query, n, 1e-8f);
    return check_launch("rankcosine");
}

int ptrb200_stlistnet_fwd_bwd(const float* scores, const float* labels, const int32_t* offsets, const float* unif,
                              float* grad, float* loss_per_query, int B, int n, float temperature,
                              uint64_t seed, uint64_t offset, ptrb200_stream_t stream) {
    int rc = check_list_args(scores, labels, grad, loss_per_query, B, n);
    if (rc) return rc;
    if (!(temperature > 0.0f)) { set_error("stlistnet: temperature must be positive"); return PTRB200_ERR_INVALID; }
    const size_t smem = (size_t)n * 4 * 2 + 33 * 4;
    if ((rc = allow_smem(stlistnet_kernel, smem))) return rc;
    PTRB200_LAUNCH(stlistnet_kernel, B, block_threads(n), smem, stream, scores, labels, offsets, unif, grad, loss_per_query,
                   n, 1.0f / temperature, seed, offset);
    return check_launch("stlistnet");
}

int ptrb200_softrank_fwd_bwd(const float* scores, const float* labels, const int32_t* offsets, float* grad, float* loss_per_query,
                             int B, int n, float delta, int top_k, ptrb200_stream_t stream) {
    int rc = check_list_args(scores, labels, grad, loss_per_query, B, n);
    if (rc) return rc;
    if (!(delta > 0.0f)) { set_error("softrank: delta must be positive"); return PTRB200_ERR_INVALID; }
    // softrank.py:50-52: pairsub_vars = 2 delta^2 (fp32 tensor arithmetic), denominator sqrt(2 * pairsub_vars)
    const float var2 = 2.0f * (delta * delta);
    const float den = sqrtf(2.0f * var2);
    const size_t smem = (size_t)n * 4 * 2 + 33 * 4;
    if ((rc = allow_smem(softrank_kernel, smem))) return rc;
    PTRB200_LAUNCH(softrank_kernel, B, block_threads(n), smem, stream, scores, labels, offsets, grad, loss_per_query, n, 1.0f / den, top_k);
    return check_launch("softrank");
}

int ptrb200_sinkstep(const float* dist, const float* log_nu, const float* log_u, float* log_v,
                     int B, int d1, int d2, float lambda, ptrb200_stream_t stream) {
    if (!dist || !log_nu || !log_u || !log_v || B <= 0 || d1 <= 0 || d2 <= 0 || !(lambda != 0.0f)) {
        set_error("sinkstep: bad arguments (B=%d d1=%d d2=%d)", B, d1, d2);
        return PTRB200_ERR_INVALID;
    }
    if (B > 65535) { set_error("sinkstep: B=%d exceeds the grid's y extent", B); return PTRB200_ERR_UNSUPPORTED; }
    dim3 grid((d2 + 31) / 32, B);
    PTRB200_LAUNCH(sinkstep_kernel, grid, 32 * SINK_WARPS, 0, stream, dist, log_nu, log_u, log_v, d1, d2, lambda);
    return check_launch("sinkstep");
}

}  // extern "C"
