// Issue cost of the packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against two scalar FFMAs on sm_100a: does packing free issue
// slots for an issue-bound kernel?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a ffma2_rate.cu -o ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int PACKED, int MIX>
__global__ void __launch_bounds__(512) rate(float* out, int iters, float seed) {
    // 8 independent chains per thread; MIX adds one integer instruction per FMA pair (a stand-in for the address / predicate
    // work that shares the issue port in the layer kernels)
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = seed + i;
    unsigned long long p[8];
    for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const unsigned long long kb = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    const unsigned long long kc = ((unsigned long long)__float_as_uint(1e-3f) << 32) | __float_as_uint(-1e-3f);
    unsigned x = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) p[i] = fma2(p[i], kb, kc);
            else { a[2 * i] = fma1(a[2 * i], 0.9999f, -1e-3f); a[2 * i + 1] = fma1(a[2 * i + 1], 1.0001f, 1e-3f); }
            if (MIX) asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, %1, 0x96;" : "+r"(x) : "r"(i));
        }
    }
    float s = 0;
    if (PACKED) for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    else for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
}

template <int PACKED, int MIX>
void run(const char* name, float* d) {
    const int iters = 4096;
    rate<PACKED, MIX><<<148 * 4, 512>>>(d, 16, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate<PACKED, MIX><<<148 * 4, 512>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = 148.0 * 4 * 512 * iters * 16;      // scalar-equivalent FMAs
    printf("%-34s %8.3f ms   %7.2f TFLOP/s fp32   %6.2f FMA lanes/clk/SM @1.965 GHz\n", name, ms, 2 * fmas / (ms * 1e-3) / 1e12, fmas / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    run<0, 0>("scalar FFMA", d);
    run<1, 0>("packed FFMA2", d);
    run<0, 1>("scalar FFMA + 1 LOP3 per pair", d);
    run<1, 1>("packed FFMA2 + 1 LOP3 per pair", d);
    return 0;
}
