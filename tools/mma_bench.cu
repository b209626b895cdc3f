// Microbenchmark (not product): cycles per tcgen05.mma for small tf32 tiles, same vs alternating accumulators.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ptranking_b200/csrc/tc.cuh"
using namespace ptrb200;

__global__ void bench(int N, int nacc, int count, int mn_major, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((float*)base)[i] = 1.0f;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
    tc::fence_proxy_async();
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = tc::instr_desc(2, 128, N);
        uint64_t a, b;
        if (mn_major) { idesc |= (1u << 15) | (1u << 16); a = tc::smem_desc_sw128_mn(tc::smem_u32(base), 4096, 512); b = tc::smem_desc_sw128_mn(tc::smem_u32(base + 16384), 4096, 512); }
        else { a = tc::smem_desc_sw128(tc::smem_u32(base), 1024); b = tc::smem_desc_sw128(tc::smem_u32(base + 16384), 1024); }
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            for (int i = 0; i < count; ++i) tc::mma_tf32(tmem + (uint32_t)(i % nacc) * 128, a, b, idesc, i >= nacc ? 1u : 0u);
            tc::mma_commit(&bar);
            long long t1 = clock64();
            tc::mbar_wait(&bar, rep & 1);
            long long t2 = clock64();
            out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
        }
    }
    tc::fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int Ns[] = {16, 112, 128, 256};
    for (int mn = 0; mn < 2; ++mn)
    for (int N : Ns) for (int nacc : {1, 2, 4}) {
        if (N == 256 && nacc > 2) continue;
        if (mn && N > 128) continue;
        const int count = 96;
        bench<<<1, 128, 66 * 1024 + 1024>>>(N, nacc, count, mn, d);
        long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mn=%d N=%3d nacc=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma  (%s)\n", mn, N, nacc, (double)h[4] / count, (double)h[5] / count, cudaGetErrorString(e));
    }
    return 0;
}
