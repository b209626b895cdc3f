// Issue rate of tcgen05.mma kind::tf32 for K-major vs MN-major shared-memory operands (one CTA per SM, operands resident,
// no loads): is the MN-major (SWIZZLE_128B_BASE32B) path that the weight-gradient kernel uses slower than K-major?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../ptranking_b200/csrc mma_rate.cu -o mma_rate && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
#include "tc.cuh"
using namespace ptrb200;

__global__ void __launch_bounds__(128) rate_kernel(int iters, int N, int a_mn, int b_mn, unsigned long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a = base;                 // 128 x 32 floats (one 128-byte chunk, 4 K-steps)
    unsigned char* b = base + 16384 * 2;     // up to 256 x 32 floats
    uint64_t* mbar = reinterpret_cast<uint64_t*>(base + 16384 * 2 + 32768 * 2);
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
    for (int i = threadIdx.x; i < (16384 * 2 + 32768 * 2) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 1.0f;
    if (threadIdx.x == 0) { tc::mbar_init(mbar, 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(slot, 256);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = tc::instr_desc(2, 128, N) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
        const uint64_t a0 = a_mn ? tc::smem_desc_sw128_mn(tc::smem_u32(a), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(a), 1024);
        const uint64_t b0 = b_mn ? tc::smem_desc_sw128_mn(tc::smem_u32(b), 4096, 512) : tc::smem_desc_sw128(tc::smem_u32(b), 1024);
        const uint64_t da = a_mn ? 64 : 2, db = b_mn ? 64 : 2;
        const unsigned long long t0 = clock64();
        if (tc::elect_one()) {
            for (int it = 0; it < iters; ++it) {
                uint64_t ad = a0, bd = b0;
                for (int s = 0; s < 4; ++s) { tc::mma_tf32(tmem, ad, bd, idesc, 1u); ad += da; bd += db; }
            }
            tc::mma_commit(mbar);
        }
        __syncwarp();
        tc::mbar_wait(mbar, 0);
        const unsigned long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 256);
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 8);
    const size_t smem = 1024 + 16384 * 2 + 32768 * 2 + 64;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    for (int N : {64, 112, 128, 144, 256}) {
        for (int mode = 0; mode < 4; ++mode) {
            const int a_mn = mode & 1, b_mn = mode >> 1;
            rate_kernel<<<148, 128, smem>>>(iters, N, a_mn, b_mn, d);   // warm
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            rate_kernel<<<148, 128, smem>>>(iters, N, a_mn, b_mn, d);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            unsigned long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const double mmas = (double)iters * 4, flop = mmas * 2.0 * 128 * N * 8;
            printf("N=%3d A %s B %s : %7.1f cycles/MMA  %6.1f TFLOP/s (148 SMs, by events)  %s\n", N, a_mn ? "MN" : "K ", b_mn ? "MN" : "K ",
                   (double)h[0] / mmas, flop * 148 / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
        }
    }
    return 0;
}
