// attention_tc.cu -- multi-head self-attention with every contraction on tcgen05 tensor cores.
//
// MultiheadAttention.forward, ptranking/base/list_ranker.py:226-248:
//     S = Q K^T / sqrt(d)   ->   A = softmax(S)   ->   A_d = dropout(A)   ->   O = A_d V
// and its autograd.  "Materialised S" design: the six contractions of forward + backward
//     S = Q K^T,   O = A_d V,   dA_d = dO V^T,   dQ = dS K,   dK = dS^T Q,   dV = A_d^T dO
// all run through a batched, strided  C[z] = alpha * op(A[z]) * B[z]^T  kernel (kind::tf32, 3xTF32 split, fp32
// accumulation in TMEM); row softmax / softmax-backward are streaming SIMT kernels over the [n,n] score tensor.  A factor that
// enters a contraction transposed (V, K, Q, dO as [key|query, d]; dS and A_d as [query, key] for dK / dV) is consumed in place
// as an MN-major tcgen05 operand, never transposed in memory.  The [n,n] tensors live in HBM (n <= 1024 per
// list: 134 MB per layer at B=64, n=512, 2 heads).
// Two kernels implement the batched GEMM: bgemm_nt_tc_kernel takes any shape; bgemm_fast_kernel (every extent, pitch and
// base address a multiple of four floats -- the attention core's own shapes) has the operand layouts and the dropout view
// as template parameters, loads one K-chunk ahead and writes full-width tiles through shared memory; both produce the
// same bits.  The row-pitched entry points (_ld) read Q|K|V side by side from one projection output and take per-query key
// counts for padded ragged batches.  A version that kept the score tile in TMEM across QK^T and the softmax was built,
// measured slower (one CTA per SM, strictly serial phases) and removed: DESIGN.md 4.
#include "common.cuh"
#include "tc.cuh"

namespace ptrb200 {

struct BGemmArgs {
    const float *A, *B;
    float* C;
    int M, N, K;                 // C[M,N] = alpha * A[M,K] * B[N,K]^T
    int lda, ldb, ldc;           // row pitches (floats)
    long long sAb, sAh, sBb, sBh, sCb, sCh;   // batch strides (floats) for z = b*H + h
    int H;
    float alpha;
    // optional dropout on A's elements: mode 1: element id = (z*M + row)*K + col ; mode 2 (A is a transposed view of
    // the attention matrix): id = (z*K + col)*M + row
    int drop_mode;
    DropCfg drop;
    // operand storage: 0 = K-major as written above (A[M,K], B[N,K] row-major); 1 = MN-major, the operand is stored
    // transposed ([K,M] resp. [K,N] row-major, pitch lda/ldb between k-rows) and consumed as an MN-major tcgen05 operand
    // (SWIZZLE_128B_BASE32B), so a transposed factor never has to be materialised.  drop_mode 2 goes with a_mn.
    int a_mn, b_mn;
};

constexpr int BG_THREADS = 256, BG_NT = 128;

template <int PASSES>
__global__ void __launch_bounds__(BG_THREADS) bgemm_nt_tc_kernel(BGemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_hi = base;
    unsigned char* a_lo = a_hi + 16384;
    unsigned char* b_hi = a_lo + 16384;
    unsigned char* b_lo = b_hi + 16384;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + 16384);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z, zb = z / g.H, zh = z % g.H;
    const float* A = g.A + zb * g.sAb + zh * g.sAh;
    const float* B = g.B + zb * g.sBb + zh * g.sBh;
    float* C = g.C + zb * g.sCb + zh * g.sCh;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BG_NT;
    const int N = min(BG_NT, g.N - n0), NP = ((N + 15) / 16) * 16;
    const int K = g.K;
    const int nchunks = (K + 31) / 32;
    // The tensor core accumulates with truncation, a bias that grows linearly with the number of accumulation steps.
    // Long contractions therefore alternate between two main accumulators, and the small a_lo*b_hi + a_hi*b_lo correction
    // terms get an accumulator of their own; the epilogue adds them up in round-to-nearest fp32.
    const int nmain = (nchunks > 4 && 3 * NP <= 256) ? 2 : 1;
    const int nacc = nmain + (PASSES == 3 ? 1 : 0);
    const uint32_t need = (uint32_t)(nacc * NP);
    const uint32_t tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : 256;
    if (tid == 0) { tc::mbar_init(mbar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(slot, tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *slot;
    const uint32_t idesc = tc::instr_desc(2, 128, NP) | (g.a_mn ? (1u << 15) : 0u) | (g.b_mn ? (1u << 16) : 0u);
    const bool vecA = (g.lda & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool vecB = (g.ldb & 3) == 0 && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    auto load4 = [&](const float* p, int k, bool vec) -> float4 {      // guarded 4-wide load at column k of a row
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k + 3 < K && vec) return __ldg(reinterpret_cast<const float4*>(p + k));
        if (k < K) v.x = p[k];
        if (k + 1 < K) v.y = p[k + 1];
        if (k + 2 < K) v.z = p[k + 2];
        if (k + 3 < K) v.w = p[k + 3];
        return v;
    };
    auto load4m = [&](const float* p, int c0, int lim, bool vec) -> float4 {   // guarded 4-wide load at column c0 (< lim) of a row
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + 3 < lim && vec) return __ldg(reinterpret_cast<const float4*>(p + c0));
        if (c0 < lim) v.x = p[c0];
        if (c0 + 1 < lim) v.y = p[c0 + 1];
        if (c0 + 2 < lim) v.z = p[c0 + 2];
        if (c0 + 3 < lim) v.w = p[c0 + 3];
        return v;
    };
    for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * 32;
        float4 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = tid + i * BG_THREADS, r = u >> 3, j = u & 7, k = k0 + j * 4;      // K-major: row r, 16-byte unit j
            const int kr = u >> 5, mu = (u & 31) * 4;                                       // MN-major: k-row kr, columns mu..mu+3
            if (g.a_mn) av[i] = (k0 + kr < K) ? load4m(A + (size_t)(k0 + kr) * g.lda, m0 + mu, g.M, vecA) : make_float4(0.f, 0.f, 0.f, 0.f);
            else av[i] = (m0 + r < g.M) ? load4(A + (size_t)(m0 + r) * g.lda, k, vecA) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.b_mn) bv[i] = (k0 + kr < K) ? load4m(B + (size_t)(k0 + kr) * g.ldb, n0 + mu, n0 + N, vecB) : make_float4(0.f, 0.f, 0.f, 0.f);
            else bv[i] = (r < N) ? load4(B + (size_t)(n0 + r) * g.ldb, k, vecB) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (c > 0) tc::mbar_wait(mbar, (c - 1) & 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int u = tid + i * BG_THREADS, r = u >> 3, j = u & 7, k = k0 + j * 4;
            const int kr = u >> 5, mu = (u & 31) * 4;
            const uint32_t off_k = tc::swz_offset(r, j);                                             // K-major slot
            const uint32_t off_mn = (uint32_t)((u & 31) >> 3) * 4096u + tc::swz32_offset(kr, u & 7);   // MN-major: 32-column chunk, k-row, unit
            float4 v = av[i];
            if (g.drop_mode && g.drop.thr) {
                float* e = &v.x;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    // element ids follow the row-major order of the stored attention matrix in both views
                    const uint64_t id = g.a_mn ? ((uint64_t)z * K + (k0 + kr)) * g.M + (m0 + mu + t)
                                               : ((uint64_t)z * g.M + (m0 + r)) * K + (k + t);
                    e[t] = dropout_keep(g.drop.key, id, g.drop.thr) ? e[t] * g.drop.scale : 0.0f;      // (out-of-range elements are already 0)
                }
            }
            float4 h, l;
            tc::split_tf32_rn(v.x, h.x, l.x); tc::split_tf32_rn(v.y, h.y, l.y); tc::split_tf32_rn(v.z, h.z, l.z); tc::split_tf32_rn(v.w, h.w, l.w);
            const uint32_t offa = g.a_mn ? off_mn : off_k;
            *reinterpret_cast<float4*>(a_hi + offa) = PASSES == 3 ? h : v;
            if (PASSES == 3) *reinterpret_cast<float4*>(a_lo + offa) = l;
            const float4 w = bv[i];
            tc::split_tf32_rn(w.x, h.x, l.x); tc::split_tf32_rn(w.y, h.y, l.y); tc::split_tf32_rn(w.z, h.z, l.z); tc::split_tf32_rn(w.w, h.w, l.w);
            const uint32_t offb = g.b_mn ? off_mn : off_k;
            *reinterpret_cast<float4*>(b_hi + offb) = PASSES == 3 ? h : w;
            if (PASSES == 3) *reinterpret_cast<float4*>(b_lo + offb) = l;
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc::fence_after_sync();
            const int ksteps = min(4, (K - k0 + 7) / 8);
            // K-major: 8-row atoms 1024 B apart, +32 B per K-step; MN-major: 32-column chunks 4096 B apart, 4-row groups 512 B
            // apart, +1024 B per K-step (descriptor addresses count 16-byte units)
            uint64_t ah = g.a_mn ? tc::smem_desc_sw128_mn(tc::smem_u32(a_hi), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(a_hi), 1024);
            uint64_t al = g.a_mn ? tc::smem_desc_sw128_mn(tc::smem_u32(a_lo), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(a_lo), 1024);
            uint64_t bh = g.b_mn ? tc::smem_desc_sw128_mn(tc::smem_u32(b_hi), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(b_hi), 1024);
            uint64_t bl = g.b_mn ? tc::smem_desc_sw128_mn(tc::smem_u32(b_lo), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(b_lo), 1024);
            const uint64_t da = g.a_mn ? 64 : 2, db = g.b_mn ? 64 : 2;
            if (tc::elect_one()) {
                const uint32_t t_main = tmem + (uint32_t)((c % nmain) * NP), t_corr = tmem + (uint32_t)(nmain * NP);
                for (int s = 0; s < ksteps; ++s) {
                    const uint32_t acc_m = (c < nmain && s == 0) ? 0u : 1u;
                    if (PASSES == 3) {
                        tc::mma_tf32(t_corr, al, bh, idesc, (c == 0 && s == 0) ? 0u : 1u);
                        tc::mma_tf32(t_corr, ah, bl, idesc, 1u);
                    }
                    tc::mma_tf32(t_main, ah, bh, idesc, acc_m);
                    ah += da; al += da; bh += db; bl += db;
                }
                tc::mma_commit(mbar);
            }
            __syncwarp();
        }
    }
    tc::mbar_wait(mbar, (nchunks - 1) & 1);
    tc::fence_after_sync();
    {   // epilogue: TMEM lane = row; warps (q, half) split the columns
        const int q = warp & 3, half = warp >> 2;
        const int row = m0 + q * 32 + lane;
        const int cols_half = ((NP / 8 + 1) / 2) * 8;
        const int c_begin = half == 0 ? 0 : cols_half, c_end = half == 0 ? min(cols_half, NP) : NP;
        float* crow = C + (size_t)min(row, g.M - 1) * g.ldc + n0;
        const bool vecC = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C + n0) & 15) == 0);
        for (int c0 = c_begin; c0 < c_end; c0 += 8) {
            float v[8];
            tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            for (int a = 1; a < nacc; ++a) {
                float w[8];
                tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NP + c0), w);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += w[e];
            }
            if (row < g.M && c0 < N) {
                if (c0 + 8 <= N && vecC) {
                    *reinterpret_cast<float4*>(crow + c0) = make_float4(v[0] * g.alpha, v[1] * g.alpha, v[2] * g.alpha, v[3] * g.alpha);
                    *reinterpret_cast<float4*>(crow + c0 + 4) = make_float4(v[4] * g.alpha, v[5] * g.alpha, v[6] * g.alpha, v[7] * g.alpha);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) if (c0 + e < N) crow[c0 + e] = v[e] * g.alpha;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

// Alignment-specialised variant of the kernel above: same tiling, same accumulators, same results bit for bit (identical
// operand images and MMA order), but the operand layouts and the dropout view are template parameters and every pitch,
// extent and base address is a multiple of four floats (host-checked).  Row/column validity, global pointers, swizzled
// shared-memory offsets and the dropout counter of a thread's four 16-byte units are then chunk-invariant and leave the
// chunk loop; one 64-bit draw serves the four elements of a unit (the general kernel hashes per element because it
// cannot assume quad alignment); B units beyond the padded tile width are never touched.  The attention core's six
// contractions all qualify; odd shapes keep the general kernel.
template <int PASSES, int A_MN, int B_MN, int DROP>
__global__ void __launch_bounds__(BG_THREADS, 2) bgemm_fast_kernel(BGemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_hi = base;
    unsigned char* a_lo = a_hi + 16384;
    unsigned char* b_hi = a_lo + 16384;
    unsigned char* b_lo = b_hi + 16384;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 128 * (BG_NT + 4) * 4);     // beyond the epilogue's staging tile, which overlays the operands
    (void)b_lo;
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z, zb = z / g.H, zh = z % g.H;
    const float* A = g.A + zb * g.sAb + zh * g.sAh;
    const float* B = g.B + zb * g.sBb + zh * g.sBh;
    float* C = g.C + zb * g.sCb + zh * g.sCh;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BG_NT;
    const int N = min(BG_NT, g.N - n0), NP = ((N + 15) / 16) * 16;
    const int K = g.K;
    const int nchunks = (K + 31) / 32;
    const int nmain = (nchunks > 4 && 3 * NP <= 256) ? 2 : 1;
    const int nacc = nmain + (PASSES == 3 ? 1 : 0);
    const uint32_t need = (uint32_t)(nacc * NP);
    const uint32_t tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : 256;
    if (tid == 0) { tc::mbar_init(mbar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc(slot, tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *slot;
    const uint32_t idesc = tc::instr_desc(2, 128, NP) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);

    // ---- chunk-invariant geometry of this thread's units (unit index i = 0..3) ----
    const int jk = tid & 7, rk = tid >> 3;            // K-major: 16-byte unit jk of row rk + 32 i
    const int cm = (tid & 31) * 4, km = tid >> 5;     // MN-major: columns cm..cm+3 of k-row km + 8 i
    const float* pa; const float* pb;
    size_t sa, sb;                                    // pointer step per unit index
    uint32_t oa, ob;                                  // swizzled offset of unit 0 (+4096 resp. +1024 per unit index)
    uint32_t am = 0, bm = 0;                          // bit i: unit i lies inside the tile along M / N (bit 4+i: and is staged at all)
    if (A_MN) {
        pa = A + (size_t)km * g.lda + m0 + cm; sa = (size_t)8 * g.lda;
        oa = (uint32_t)(cm >> 5) * 4096u + tc::swz32_offset(km, jk);
        am = (m0 + cm < g.M) ? 0xfu : 0u;
    } else {
        pa = A + (size_t)(m0 + rk) * g.lda + jk * 4; sa = (size_t)32 * g.lda;
        oa = tc::swz_offset(rk, jk);
#pragma unroll
        for (int i = 0; i < 4; ++i) am |= (m0 + rk + 32 * i < g.M) ? (1u << i) : 0u;
    }
    if (B_MN) {
        pb = B + (size_t)km * g.ldb + n0 + cm; sb = (size_t)8 * g.ldb;
        ob = (uint32_t)(cm >> 5) * 4096u + tc::swz32_offset(km, jk);
        bm = (cm < N ? 0xfu : 0u) | (cm < NP ? 0xf0u : 0u);
    } else {
        pb = B + (size_t)(n0 + rk) * g.ldb + jk * 4; sb = (size_t)32 * g.ldb;
        ob = tc::swz_offset(rk, jk);
#pragma unroll
        for (int i = 0; i < 4; ++i) bm |= (rk + 32 * i < N ? (1u << i) : 0u) | (rk + 32 * i < NP ? (16u << i) : 0u);
    }
    constexpr uint32_t OA_STEP = A_MN ? 1024u : 4096u, OB_STEP = B_MN ? 1024u : 4096u;
    // dropout counter: key + GOLD * quad, quad = element id / 4 of the unit's first element (ids as in the general kernel)
    constexpr uint64_t GOLD = 0x9e3779b97f4a7c15ull;
    uint64_t dctr = 0, dstep_i = 0, dstep_c = 0;
    if (DROP == 1) {            // id = (z*M + row)*K + col ; unit step: 32 rows ; chunk step: 32 columns
        dctr = g.drop.key + GOLD * ((((uint64_t)z * g.M + (m0 + rk)) * K + jk * 4) >> 2);
        dstep_i = GOLD * (uint64_t)(8 * K); dstep_c = GOLD * 8ull;
    } else if (DROP == 2) {     // id = (z*K + krow)*M + col ; unit step: 8 k-rows ; chunk step: 32 k-rows
        dctr = g.drop.key + GOLD * ((((uint64_t)z * K + km) * g.M + (m0 + cm)) >> 2);
        dstep_i = GOLD * (uint64_t)(2 * g.M); dstep_c = GOLD * (uint64_t)(8 * g.M);
    }
    const bool drop_on = DROP != 0 && g.drop.thr != 0;
    const uint32_t thr = g.drop.thr; const float dscale = g.drop.scale;

    // global -> registers one chunk AHEAD: the loads of chunk c+1 are in flight while chunk c is split, stored, synchronised
    // and multiplied (a CTA has no other way to hide HBM latency: its chunks are strictly sequential)
    float4 nav[4], nbv[4];
    auto load_chunk = [&](int c) {
        const int k0 = c * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ka = A_MN ? (k0 + km + 8 * i < K) : (k0 + jk * 4 < K);
            const bool kb = B_MN ? (k0 + km + 8 * i < K) : (k0 + jk * 4 < K);
            nav[i] = (ka && ((am >> i) & 1u)) ? __ldg(reinterpret_cast<const float4*>(pa + i * sa)) : make_float4(0.f, 0.f, 0.f, 0.f);
            nbv[i] = (kb && ((bm >> i) & 1u)) ? __ldg(reinterpret_cast<const float4*>(pb + i * sb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        pa += A_MN ? (size_t)32 * g.lda : 32; pb += B_MN ? (size_t)32 * g.ldb : 32;
    };
    load_chunk(0);
    for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * 32;
        float4 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { av[i] = nav[i]; bv[i] = nbv[i]; }
        if (c + 1 < nchunks) load_chunk(c + 1);
        if (c > 0) tc::mbar_wait(mbar, (c - 1) & 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v = av[i];
            if (DROP != 0 && drop_on) {
                const uint64_t d = mix64(dctr + i * dstep_i);
                v.x = ((uint32_t)d & 0xffffu) >= thr ? v.x * dscale : 0.0f;
                v.y = ((uint32_t)(d >> 16) & 0xffffu) >= thr ? v.y * dscale : 0.0f;
                v.z = ((uint32_t)(d >> 32) & 0xffffu) >= thr ? v.z * dscale : 0.0f;
                v.w = (uint32_t)(d >> 48) >= thr ? v.w * dscale : 0.0f;
            }
            float4 h, l;
            tc::split_tf32_rn(v.x, h.x, l.x); tc::split_tf32_rn(v.y, h.y, l.y); tc::split_tf32_rn(v.z, h.z, l.z); tc::split_tf32_rn(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(a_hi + oa + i * OA_STEP) = PASSES == 3 ? h : v;
            if (PASSES == 3) *reinterpret_cast<float4*>(a_lo + oa + i * OA_STEP) = l;
            if ((bm >> (4 + i)) & 1u) {
                const float4 w = bv[i];
                tc::split_tf32_rn(w.x, h.x, l.x); tc::split_tf32_rn(w.y, h.y, l.y); tc::split_tf32_rn(w.z, h.z, l.z); tc::split_tf32_rn(w.w, h.w, l.w);
                *reinterpret_cast<float4*>(b_hi + ob + i * OB_STEP) = PASSES == 3 ? h : w;
                if (PASSES == 3) *reinterpret_cast<float4*>(b_lo + ob + i * OB_STEP) = l;
            }
        }
        dctr += dstep_c;
        tc::fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc::fence_after_sync();
            const int ksteps = min(4, (K - k0 + 7) / 8);
            uint64_t ah = A_MN ? tc::smem_desc_sw128_mn(tc::smem_u32(a_hi), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(a_hi), 1024);
            uint64_t al = A_MN ? tc::smem_desc_sw128_mn(tc::smem_u32(a_lo), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(a_lo), 1024);
            uint64_t bh = B_MN ? tc::smem_desc_sw128_mn(tc::smem_u32(b_hi), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(b_hi), 1024);
            uint64_t bl = B_MN ? tc::smem_desc_sw128_mn(tc::smem_u32(b_lo), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(b_lo), 1024);
            constexpr uint64_t da = A_MN ? 64 : 2, db = B_MN ? 64 : 2;
            if (tc::elect_one()) {
                const uint32_t t_main = tmem + (uint32_t)((c % nmain) * NP), t_corr = tmem + (uint32_t)(nmain * NP);
                for (int s = 0; s < ksteps; ++s) {
                    const uint32_t acc_m = (c < nmain && s == 0) ? 0u : 1u;
                    if (PASSES == 3) {
                        tc::mma_tf32(t_corr, al, bh, idesc, (c == 0 && s == 0) ? 0u : 1u);
                        tc::mma_tf32(t_corr, ah, bl, idesc, 1u);
                    }
                    tc::mma_tf32(t_main, ah, bh, idesc, acc_m);
                    ah += da; al += da; bh += db; bl += db;
                }
                tc::mma_commit(mbar);
            }
            __syncwarp();
        }
    }
    tc::mbar_wait(mbar, (nchunks - 1) & 1);
    tc::fence_after_sync();
    const int q = warp & 3, half = warp >> 2;
    const float alpha = g.alpha;
    if (N == BG_NT) {
        // full 128-column tile (the [n,n] score / dP tensors: the 134 MB outputs of the attention core).  TMEM lane = row, so a
        // direct store has every lane of a warp writing into a different row (32-byte pieces of 32 rows per instruction);
        // instead the tile goes through shared memory (the operand buffers are free: every MMA has completed) and leaves as
        // whole 512-byte row segments, one row per warp instruction.
        constexpr int PITCH = BG_NT + 4;                  // floats; 132 = 4 (mod 32): the row-strided float4 writes are conflict-free
        float* tile = reinterpret_cast<float*>(base);     // [128][PITCH] = 67,584 B (operands 65,536 B + the slack the host adds)
        __syncthreads();                                  // (all threads are past their last operand stores; belt and braces)
        for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 8) {
            float v[8];
            tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            for (int a = 1; a < nacc; ++a) {
                float w[8];
                tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NP + c0), w);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += w[e];
            }
            float* t = tile + (q * 32 + lane) * PITCH + c0;
            *reinterpret_cast<float4*>(t) = make_float4(v[0] * alpha, v[1] * alpha, v[2] * alpha, v[3] * alpha);
            *reinterpret_cast<float4*>(t + 4) = make_float4(v[4] * alpha, v[5] * alpha, v[6] * alpha, v[7] * alpha);
        }
        __syncthreads();
        const int rows_here = min(128, g.M - m0);
        for (int r = warp; r < rows_here; r += BG_THREADS / 32)
            *reinterpret_cast<float4*>(C + (size_t)(m0 + r) * g.ldc + n0 + lane * 4) = *reinterpret_cast<const float4*>(tile + r * PITCH + lane * 4);
    } else {   // narrow tile: TMEM lane = row; warps (q, half) split the columns; every extent is a multiple of 4
        const int row = m0 + q * 32 + lane;
        const int cols_half = ((NP / 8 + 1) / 2) * 8;
        const int c_begin = half == 0 ? 0 : cols_half, c_end = half == 0 ? min(cols_half, NP) : NP;
        float* crow = C + (size_t)min(row, g.M - 1) * g.ldc + n0;
        for (int c0 = c_begin; c0 < c_end; c0 += 8) {
            float v[8];
            tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            for (int a = 1; a < nacc; ++a) {
                float w[8];
                tc::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NP + c0), w);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += w[e];
            }
            if (row < g.M) {
                if (c0 < N) *reinterpret_cast<float4*>(crow + c0) = make_float4(v[0] * alpha, v[1] * alpha, v[2] * alpha, v[3] * alpha);
                if (c0 + 4 < N) *reinterpret_cast<float4*>(crow + c0 + 4) = make_float4(v[4] * alpha, v[5] * alpha, v[6] * alpha, v[7] * alpha);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

// in-place row softmax over S[z][i][:] (one warp per row); also emits the per-row log-sum-exp.
// key_lens (optional, [B]): only the first key_lens[b] keys of query b exist (a ragged batch padded to n); the others get
// probability exactly 0, so padded documents influence nothing (their own rows are never read back).
__global__ void softmax_rows_kernel(float* __restrict__ S, float* __restrict__ lse, size_t rows, int n,
                                    const int32_t* __restrict__ key_lens, int rows_per_query) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* s = S + row * n;
    const int nk = key_lens ? max(1, min(n, key_lens[row / (size_t)rows_per_query])) : n;
    float m = -INFINITY;
    for (int j = lane; j < nk; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    float l = 0.0f;
    for (int j = lane; j < nk; j += 32) { const float e = expf(s[j] - m); s[j] = e; l += e; }
    l = warp_sum(l);
    const float inv = 1.0f / l;
    for (int j = lane; j < nk; j += 32) s[j] *= inv;
    for (int j = nk + lane; j < n; j += 32) s[j] = 0.0f;
    if (lane == 0 && lse) lse[row] = m + logf(l);
}

// dS = A * (dA - sum_j A dA) * inv_scale in place over dAd, where dA = dropmask(dAd)  (one warp per row)
__global__ void softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dP, size_t rows, int n,
                                        float inv_scale, DropCfg drop) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = P + row * n;
    float* d = dP + row * n;
    float acc = 0.0f;
    for (int j = lane; j < n; j += 32) {
        float da = d[j];
        if (drop.thr) da = dropout_keep(drop.key, row * (uint64_t)n + j, drop.thr) ? da * drop.scale : 0.0f;
        d[j] = da;
        acc = fmaf(p[j], da, acc);
    }
    acc = warp_sum(acc);
    for (int j = lane; j < n; j += 32) d[j] = p[j] * (d[j] - acc) * inv_scale;
}

template <int PASSES, int A_MN, int B_MN, int DROP>
static int launch_bgemm_fast(const BGemmArgs& g, dim3 grid, size_t smem, cudaStream_t st, const char* tag) {
    const cudaError_t e = cudaFuncSetAttribute(bgemm_fast_kernel<PASSES, A_MN, B_MN, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("bgemm smem attr: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
    PTRB200_LAUNCH_TAG(tag, (bgemm_fast_kernel<PASSES, A_MN, B_MN, DROP>), grid, BG_THREADS, smem, st, g);
    return PTRB200_OK;
}

static bool bgemm_general_forced() {
    const char* e = getenv("PTRB200_BGEMM_GENERAL");      // read per launch: the parity test flips it inside one process
    return e && e[0] == '1';
}

static int launch_bgemm(BGemmArgs& g, int Z, int passes, cudaStream_t st, const char* tag) {
    const size_t smem = 1024 + 4 * 16384 + 64;
    dim3 grid((g.M + 127) / 128, (g.N + BG_NT - 1) / BG_NT, Z);
    // the alignment-specialised kernel: every pitch, stride, extent and base address a multiple of four floats
    const auto q4 = [](long long v) { return (v & 3) == 0; };
    const auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool fast = !bgemm_general_forced() && q4(g.lda) && q4(g.ldb) && q4(g.ldc) && q4(g.M) && q4(g.N) && q4(g.K) &&
                      q4(g.sAb) && q4(g.sAh) && q4(g.sBb) && q4(g.sBh) && q4(g.sCb) && q4(g.sCh) && a16(g.A) && a16(g.B) && a16(g.C) &&
                      (g.drop_mode == 0 || (g.drop_mode == 1 && !g.a_mn) || (g.drop_mode == 2 && g.a_mn));
    if (fast) {
        const int drop = g.drop.thr ? g.drop_mode : 0;
        const size_t smem = 1024 + (size_t)128 * (BG_NT + 4) * 4 + 64;      // operands (64 KB) overlaid by the [128][132] output staging tile
#define PTRB200_BG_CASE(P, AM, BM, D) if ((passes == 3) == (P == 3) && g.a_mn == AM && g.b_mn == BM && drop == D) return launch_bgemm_fast<P, AM, BM, D>(g, grid, smem, st, tag);
        // the attention core's shapes (list_ranker.py:226-248 forward + autograd), 3xTF32 and single-pass
        PTRB200_BG_CASE(3, 0, 0, 0) PTRB200_BG_CASE(3, 0, 1, 0) PTRB200_BG_CASE(3, 0, 1, 1) PTRB200_BG_CASE(3, 1, 1, 0) PTRB200_BG_CASE(3, 1, 1, 2)
        PTRB200_BG_CASE(1, 0, 0, 0) PTRB200_BG_CASE(1, 0, 1, 0) PTRB200_BG_CASE(1, 0, 1, 1) PTRB200_BG_CASE(1, 1, 1, 0) PTRB200_BG_CASE(1, 1, 1, 2)
#undef PTRB200_BG_CASE
    }
    cudaError_t e;
    if (passes == 3) {
        e = cudaFuncSetAttribute(bgemm_nt_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("bgemm smem attr: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
        PTRB200_LAUNCH_TAG(tag, bgemm_nt_tc_kernel<3>, grid, BG_THREADS, smem, st, g);
    } else {
        e = cudaFuncSetAttribute(bgemm_nt_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("bgemm smem attr: %s", cudaGetErrorString(e)); return PTRB200_ERR_CUDA; }
        PTRB200_LAUNCH_TAG(tag, bgemm_nt_tc_kernel<1>, grid, BG_THREADS, smem, st, g);
    }
    return PTRB200_OK;
}

}  // namespace ptrb200

using namespace ptrb200;

extern "C" {

// scratch (floats): the backward pass needs dS[Z,n,n]; the forward pass none (a 4-float minimum keeps allocations non-empty)
int64_t ptrb200_attention_tc_workspace_floats(int B, int n, int H, int D, int backward) {
    const int64_t Z = (int64_t)B * H, nn = (int64_t)n * n;
    (void)D;
    return backward ? Z * nn : 4;
}

// P_out[B*H, n, n] receives the (un-dropped) attention probabilities and must be kept for the backward pass.
int ptrb200_attention_tc_fwd_ld(const float* Q, const float* K, const float* V, float* O, float* P_out, float* scratch,
                                int B, int n, int H, int D, int ld_qkv, int ld_o, const int32_t* key_lens, float dropout_p,
                                uint64_t seed, uint64_t offset, int passes, ptrb200_stream_t stream) {
    if (!Q || !K || !V || !O || !P_out || !scratch || B <= 0 || n <= 0 || H <= 0 || D <= 0) { set_error("attention_tc_fwd: bad arguments"); return PTRB200_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int Z = B * H, HD = H * D;
    const int lq = ld_qkv > 0 ? ld_qkv : HD, lo = ld_o > 0 ? ld_o : HD;
    if (lq < HD || lo < HD) { set_error("attention_tc_fwd: row pitch below H*D"); return PTRB200_ERR_INVALID; }
    const long long sb = (long long)n * lq, sbo = (long long)n * lo, sh = D, nn = (long long)n * n;
    int rc;
    BGemmArgs g{};
    // S = Q K^T / sqrt(D)
    g.A = Q; g.B = K; g.C = P_out; g.M = n; g.N = n; g.K = D; g.lda = lq; g.ldb = lq; g.ldc = n;
    g.sAb = sb; g.sAh = sh; g.sBb = sb; g.sBh = sh; g.sCb = nn * H; g.sCh = nn; g.H = H; g.alpha = 1.0f / sqrtf((float)D);
    if ((rc = launch_bgemm(g, Z, passes, st, "attn_tc_qk"))) return rc;
    const size_t rows = (size_t)Z * n;
    PTRB200_LAUNCH(softmax_rows_kernel, (unsigned)((rows + 7) / 8), 256, 0, st, P_out, (float*)nullptr, rows, n, key_lens, H * n);
    // O = dropout(P) V : V is the [K = key, N = d] row-major factor, consumed MN-major
    (void)scratch;
    BGemmArgs o{};
    o.A = P_out; o.B = V; o.C = O; o.M = n; o.N = D; o.K = n; o.lda = n; o.ldb = lq; o.ldc = lo; o.b_mn = 1;
    o.sAb = nn * H; o.sAh = nn; o.sBb = sb; o.sBh = sh; o.sCb = sbo; o.sCh = sh; o.H = H; o.alpha = 1.0f;
    o.drop_mode = 1; o.drop = make_drop(dropout_p, seed, offset);
    if ((rc = launch_bgemm(o, Z, passes, st, "attn_tc_pv"))) return rc;
    return check_launch("attention_tc_fwd");
}

int ptrb200_attention_tc_fwd(const float* Q, const float* K, const float* V, float* O, float* P_out, float* scratch,
                             int B, int n, int H, int D, float dropout_p, uint64_t seed, uint64_t offset, int passes,
                             ptrb200_stream_t stream) {
    return ptrb200_attention_tc_fwd_ld(Q, K, V, O, P_out, scratch, B, n, H, D, 0, 0, nullptr, dropout_p, seed, offset, passes, stream);
}

int ptrb200_attention_tc_bwd_ld(const float* Q, const float* K, const float* V, const float* P, const float* dO,
                                float* dQ, float* dK, float* dV, float* scratch,
                                int B, int n, int H, int D, int ld_qkv, int ld_o, float dropout_p, uint64_t seed,
                                uint64_t offset, int passes, ptrb200_stream_t stream) {
    if (!Q || !K || !V || !P || !dO || !dQ || !dK || !dV || !scratch || B <= 0 || n <= 0 || H <= 0 || D <= 0) { set_error("attention_tc_bwd: bad arguments"); return PTRB200_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int Z = B * H, HD = H * D;
    const int lq = ld_qkv > 0 ? ld_qkv : HD, lo = ld_o > 0 ? ld_o : HD;
    if (lq < HD || lo < HD) { set_error("attention_tc_bwd: row pitch below H*D"); return PTRB200_ERR_INVALID; }
    const long long sb = (long long)n * lq, sbo = (long long)n * lo, sh = D, nn = (long long)n * n;
    float* dS = scratch;                    // [Z,n,n]
    const DropCfg drop = make_drop(dropout_p, seed, offset);
    const float inv_scale = 1.0f / sqrtf((float)D);
    int rc;
    // dA_d = dO V^T
    BGemmArgs a{};
    a.A = dO; a.B = V; a.C = dS; a.M = n; a.N = n; a.K = D; a.lda = lo; a.ldb = lq; a.ldc = n;
    a.sAb = sbo; a.sAh = sh; a.sBb = sb; a.sBh = sh; a.sCb = nn * H; a.sCh = nn; a.H = H; a.alpha = 1.0f;
    if ((rc = launch_bgemm(a, Z, passes, st, "attn_tc_dp"))) return rc;
    const size_t rows = (size_t)Z * n;
    PTRB200_LAUNCH(softmax_bwd_rows_kernel, (unsigned)((rows + 7) / 8), 256, 0, st, P, dS, rows, n, inv_scale, drop);
    // dQ = dS K : K is the [K = key, N = d] factor (MN-major B)
    BGemmArgs q{};
    q.A = dS; q.B = K; q.C = dQ; q.M = n; q.N = D; q.K = n; q.lda = n; q.ldb = lq; q.ldc = lq; q.b_mn = 1;
    q.sAb = nn * H; q.sAh = nn; q.sBb = sb; q.sBh = sh; q.sCb = sb; q.sCh = sh; q.H = H; q.alpha = 1.0f;
    if ((rc = launch_bgemm(q, Z, passes, st, "attn_tc_dq"))) return rc;
    // dK = dS^T Q : dS itself is the [K = query, M = key] factor (MN-major A), Q the [K = query, N = d] factor (MN-major B)
    BGemmArgs k{};
    k.A = dS; k.B = Q; k.C = dK; k.M = n; k.N = D; k.K = n; k.lda = n; k.ldb = lq; k.ldc = lq; k.a_mn = 1; k.b_mn = 1;
    k.sAb = nn * H; k.sAh = nn; k.sBb = sb; k.sBh = sh; k.sCb = sb; k.sCh = sh; k.H = H; k.alpha = 1.0f;
    if ((rc = launch_bgemm(k, Z, passes, st, "attn_tc_dk"))) return rc;
    // dV = dropout(P)^T dO : same shapes, the dropout mask regenerated through the transposed view
    BGemmArgs v{};
    v.A = P; v.B = dO; v.C = dV; v.M = n; v.N = D; v.K = n; v.lda = n; v.ldb = lo; v.ldc = lq; v.a_mn = 1; v.b_mn = 1;
    v.sAb = nn * H; v.sAh = nn; v.sBb = sbo; v.sBh = sh; v.sCb = sb; v.sCh = sh; v.H = H; v.alpha = 1.0f;
    v.drop_mode = 2; v.drop = drop;
    if ((rc = launch_bgemm(v, Z, passes, st, "attn_tc_dv"))) return rc;
    return check_launch("attention_tc_bwd");
}

int ptrb200_attention_tc_bwd(const float* Q, const float* K, const float* V, const float* P, const float* dO,
                             float* dQ, float* dK, float* dV, float* scratch,
                             int B, int n, int H, int D, float dropout_p, uint64_t seed, uint64_t offset, int passes,
                             ptrb200_stream_t stream) {
    return ptrb200_attention_tc_bwd_ld(Q, K, V, P, dO, dQ, dK, dV, scratch, B, n, H, D, 0, 0, dropout_p, seed, offset, passes, stream);
}

}  // extern "C"
