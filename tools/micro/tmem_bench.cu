// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM for W warps (one CTA), and FMNMX3 / FFMA2 checks.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// mode 0: ld only (4 loads in flight, then wait); mode 1: ld, wait each; mode 2: st only
__global__ void __launch_bounds__(512, 1) tmem_kernel(int mode, int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
    tmem_st32(base, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            uint32_t a[32], b[32], c[32], d[32];
            tmem_ld32(base + 0, a);
            tmem_ld32(base + 32, b);
            tmem_ld32(base + 64, c);
            tmem_ld32(base + 96, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j] ^ c[j] ^ d[j];
        }
    } else if (mode == 1) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t a[32];
                tmem_ld32(base + 32 * q, a);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= a[j];
            }
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            tmem_st32(base + 0, v);
            tmem_st32(base + 32, v);
            tmem_st32(base + 64, v);
            tmem_st32(base + 96, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr));
}

// ALU mixes: mode 0: fmnmx 2-input; 1: 3-input max; 2: ffma2 packed; 3: ffma scalar; 4: ex2 MUFU; 5: cvt.f16x2 pack
__global__ void __launch_bounds__(512, 1) alu_kernel(int mode, int iters, long long* cycles, float* sink, float seed) {
    float x[16];
    for (int j = 0; j < 16; ++j) x[j] = seed + threadIdx.x * 1e-3f + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], x[(j + 1) & 15] - 1.f);
        } else if (mode == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(x[(j + 5) & 15]), "f"(seed));
        } else if (mode == 2) {
#pragma unroll
            for (int j = 0; j < 16; j += 2)
                asm volatile("{.reg .b64 a, b;\n\tmov.b64 a, {%0,%1};\n\tmov.b64 b, {%2,%2};\n\tfma.rn.f32x2 a, a, b, b;\n\tmov.b64 {%0,%1}, a;}"
                             : "+f"(x[j]), "+f"(x[j + 1]) : "f"(seed));
        } else if (mode == 3) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], seed, seed);
        } else if (mode == 4) {
#pragma unroll
            for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                uint32_t r;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[j]), "f"(x[j + 1]));
                x[j] = __uint_as_float(r);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0;
    for (int j = 0; j < 16; ++j) s += x[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, 1024 * 8);
    cudaMalloc(&sink, 1024 * 512 * 4);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            tmem_kernel<<<1, warps * 32>>>(mode, iters, cyc, sink);
            tmem_kernel<<<1, warps * 32>>>(mode, iters, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * 4 * 32 * 4 * 32 * warps;
            printf("tmem mode %d (%s) warps %2d: %lld cycles, %.1f B/clk/SM, %.1f cycles per x32 instr per warp  [%s]\n", mode,
                   mode == 0 ? "ld x4 then wait" : mode == 1 ? "ld+wait each" : "st x4 then wait", warps, c, bytes / c,
                   (double)c / (iters * 4), cudaGetErrorString(e));
        }
    const char* names[] = {"fmnmx", "max3", "ffma2", "ffma", "ex2", "cvt.f16x2"};
    for (int mode = 0; mode < 6; ++mode)
        for (int warps : {4, 8, 16}) {
            alu_kernel<<<1, warps * 32>>>(mode, iters, cyc, (float*)sink, 0.5f);
            alu_kernel<<<1, warps * 32>>>(mode, iters, cyc, (float*)sink, 0.5f);
            cudaError_t e = cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double instr = (double)iters * (mode == 2 || mode == 5 ? 8 : 16) * warps;
            printf("alu %-9s warps %2d: %.2f warp-instr/clk/SM  [%s]\n", names[mode], warps, instr / c, cudaGetErrorString(e));
        }
    return 0;
}
