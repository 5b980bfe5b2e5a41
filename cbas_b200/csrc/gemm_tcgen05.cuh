// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   A: activations, row-major [M,K] bf16 (K-major).  W: nn.Linear weight, row-major [N,K] bf16 (K-major).
//   TMA (SWIZZLE_128B boxes of 64 x rows) -> shared-memory ring -> tcgen05.mma (128 x BLOCK_N x 16, fp32
//   accumulators in TMEM, double-buffered) -> tcgen05.ld -> epilogue in registers -> global.
//
// Warp roles (one CTA per SM, 128 + 32*EPI_WARPS threads):
//   warp 0  : TMA producer (one elected lane)
//   warp 1  : MMA issuer   (one elected lane)
//   warp 2  : TMEM allocator / deallocator
//   warp 3  : idle
//   warps 4+: epilogue; warp w reads TMEM lanes 32*(w%4).., column slice (w-4)/4 of the tile.
//
// The epilogues are the ones the DINOv3 block needs (reference: HF modeling_dinov3_vit.py:305-311 QKV bias,
// :385-386 up_proj+GELU, :440-441 / :447-448 LayerScale+residual (LayerScale is folded into W and bias on
// the host), :71-92 patch-embedding rows interleaved behind the CLS/register prefix).
#pragma once
#include "ptx.cuh"

namespace cbas {

enum GemmEpilogue : int {
    EPI_BIAS_BF16 = 0,       // out_bf16[m,n] = acc + bias[n]
    EPI_BIAS_GELU_BF16 = 1,  // out_bf16[m,n] = gelu_erf(acc + bias[n])
    EPI_RESID_F32 = 2,       // resid_f32[m,n] += acc + bias[n]          (residual stream, in place)
    EPI_PATCH_F32 = 3,       // resid_f32[row_map(m),n] = acc + bias[n]  (patch rows behind the prefix tokens)
    EPI_BIAS_F32 = 4,        // out_f32[m,n] = acc + bias[n]
};

struct GemmParams {
    int M, N, K;
    const float* bias;  // [N] fp32 (may be null)
    void* out;          // bf16 or fp32, row stride ldo elements
    int ldo;
    // EPI_PATCH_F32: A row m = frame*rows_in + p  ->  out row frame*rows_out + prefix + p
    int rows_in, rows_out, prefix;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_UMMA_K = 16;

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int kStageA = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KB
    static constexpr int kStageB = BLOCK_N * GEMM_BLOCK_K * 2;
    static constexpr int kStage = kStageA + kStageB;
    static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 192 ? 5 : 6);
    static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int EPI, int EPI_WARPS>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
    static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "EPI_WARPS");
    constexpr int kColSlices = EPI_WARPS / 4;
    constexpr int kColsPerWarp = BLOCK_N / kColSlices;
    static_assert(kColsPerWarp % 32 == 0, "column slice must be a multiple of 32");

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kStageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStage);
    uint64_t* full_bar = bars;                    // [kStages] TMA -> MMA
    uint64_t* empty_bar = bars + kStages;         // [kStages] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kStages;     // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_blocks = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int n_blocks = p.N / BLOCK_N;
    const int num_tiles = m_blocks * n_blocks;
    const int k_blocks = p.K / GEMM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
                    tma_load_2d(smem_a + stage * Cfg::kStageA, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K,
                                m_blk * GEMM_BLOCK_M);
                    tma_load_2d(smem_b + stage * Cfg::kStageB, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                n_blk * BLOCK_N);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
            uint32_t stage = 0, phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
                const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * Cfg::kStageA));
                    const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * Cfg::kStageB));
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the (>>4) address field
                        umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue
        const int ew = warp - 4;
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
        const int col0 = (ew >> 2) * kColsPerWarp;
        int iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
            const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
            const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * GEMM_BLOCK_M + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            long long out_row = row;
            if (EPI == EPI_PATCH_F32) {
                const int f = row / p.rows_in;
                out_row = (long long)f * p.rows_out + p.prefix + (row - f * p.rows_in);
            }
#pragma unroll 1
            for (int c = 0; c < kColsPerWarp; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc * BLOCK_N + col0 + c + (uint32_t(quarter * 32) << 16);
                tmem_ld_32x32(taddr, v);
                tmem_ld_wait();
                const int n0 = n_blk * BLOCK_N + col0 + c;
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    x[j + 0] = __uint_as_float(v[j + 0]) + b.x;
                    x[j + 1] = __uint_as_float(v[j + 1]) + b.y;
                    x[j + 2] = __uint_as_float(v[j + 2]) + b.z;
                    x[j + 3] = __uint_as_float(v[j + 3]) + b.w;
                }
                if (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
                }
                if (row_ok) {
                    if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 q;
                            q.x = pack_bf16(x[j + 0], x[j + 1]);
                            q.y = pack_bf16(x[j + 2], x[j + 3]);
                            q.z = pack_bf16(x[j + 4], x[j + 5]);
                            q.w = pack_bf16(x[j + 6], x[j + 7]);
                            *reinterpret_cast<uint4*>(o + j) = q;
                        }
                    } else {
                        float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 r = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                            if (EPI == EPI_RESID_F32) {
                                const float4 h = *reinterpret_cast<const float4*>(o + j);
                                r.x += h.x; r.y += h.y; r.z += h.z; r.w += h.w;
                            }
                            *reinterpret_cast<float4*>(o + j) = r;
                        }
                    }
                }
            }
            // all of this warp's tcgen05.ld have completed (wait::ld above): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace cbas
