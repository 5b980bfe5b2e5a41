// Micro-benchmark: per-SM throughput of ex2.approx.f32 vs ex2.approx.f16x2 vs an FMA-pipe polynomial exp2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mufu_bench tools/micro/mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2poly(float x) {
    x = fmaxf(x, -30.f);
    const float t = x + 12582912.f;
    const float n = t - 12582912.f;
    const float f = x - n;
    float p = fmaf(f, 0.0555041086648f, 0.2402265069591f);
    p = fmaf(p, f, 0.6931471805599f);
    p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
    uint32_t h0 = threadIdx.x, h1 = h0 + 77, h2 = h0 + 99, h3 = h0 + 3;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { a0 = ex2f(a0); a1 = ex2f(a1); a2 = ex2f(a2); a3 = ex2f(a3); a0 -= 1.f; a1 -= 1.f; a2 -= 1.f; a3 -= 1.f; }
        if (MODE == 1) { h0 = ex2h2(h0); h1 = ex2h2(h1); h2 = ex2h2(h2); h3 = ex2h2(h3); h0 ^= 0x80008000u; h1 ^= 0x80008000u; h2 ^= 0x80008000u; h3 ^= 0x80008000u; }
        if (MODE == 2) { a0 = ex2poly(a0) - 1.5f; a1 = ex2poly(a1) - 1.5f; a2 = ex2poly(a2) - 1.5f; a3 = ex2poly(a3) - 1.5f; }
    }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 4096;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(out, clk, iters);
                if (mode == 1) k<1><<<148, threads>>>(out, clk, iters);
                if (mode == 2) k<2><<<148, threads>>>(out, clk, iters);
            }
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
            const double elems = (double)threads * iters * 4 * (mode == 1 ? 2 : 1);
            printf("threads %4d mode %d (%s): %lld clk, %.2f exp/clk/SM\n", threads, mode,
                   mode == 0 ? "ex2.f32" : mode == 1 ? "ex2.f16x2" : "poly fp32", c, elems / c);
        }
    }
    // accuracy of the polynomial
    return 0;
}
