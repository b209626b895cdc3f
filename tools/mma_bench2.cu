// Microbenchmark (not product): do two warps issuing tcgen05.mma concurrently double the issue rate?
#include <cstdio>
#include <cuda_runtime.h>
#include "../ptranking_b200/csrc/tc.cuh"
using namespace ptrb200;

__global__ void bench(int N, int issuers, int count, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((float*)base)[i] = 1.0f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(&bar[i], 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
    tc::fence_proxy_async();
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = slot;
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        uint32_t idesc = tc::instr_desc(2, 128, N);
        uint64_t a = tc::smem_desc_sw128(tc::smem_u32(base), 1024), b = tc::smem_desc_sw128(tc::smem_u32(base + 16384), 1024);
        long long t0 = clock64();
        for (int i = 0; i < count; ++i) tc::mma_tf32(tmem + (uint32_t)w * 128, a, b, idesc, i ? 1u : 0u);
        tc::mma_commit(&bar[w]);
        tc::mbar_wait(&bar[w], 0);
        long long t2 = clock64();
        out[w] = t2 - t0;
    }
    tc::fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

// plain loop without the setp/predicate wrapper: raw cost of the instruction stream
__global__ void bench_unrolled(int N, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = tc::instr_desc(2, 128, N);
        uint64_t a = tc::smem_desc_sw128(tc::smem_u32(base), 1024), b = tc::smem_desc_sw128(tc::smem_u32(base + 16384), 1024);
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc));
        long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc::fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(bench_unrolled, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int N : {112, 256}) for (int issuers : {1, 2, 4}) {
        const int count = 96;
        bench<<<1, 128, 67 * 1024>>>(N, issuers, count, d);
        long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("N=%3d issuers=%d: per-issuer cycles/mma:", N, issuers);
        for (int w = 0; w < issuers; ++w) printf(" %6.1f", (double)h[w] / count);
        printf("  -> aggregate %.1f cyc/mma (%s)\n", (double)h[0] / (count * issuers), cudaGetErrorString(e));
    }
    for (int N : {16, 112, 256}) {
        bench_unrolled<<<1, 128, 67 * 1024>>>(N, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("unrolled N=%3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma\n", N, (double)h[0] / 32, (double)h[1] / 32);
    }
    return 0;
}
