// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   A: activations, row-major [M,K] bf16 (K-major).  W: nn.Linear weight, row-major [N,K] bf16 (K-major).
//   TMA (SWIZZLE_128B boxes of 64 x rows) -> shared-memory ring -> tcgen05.mma (128 x BLOCK_N x 16, fp32
//   accumulators in TMEM, double-buffered) -> tcgen05.ld -> epilogue math in registers -> swizzled
//   shared-memory slab -> TMA store (or TMA add-reduction into the fp32 residual stream) -> global.
//
// Warp roles (one CTA per SM, 512 + 128 threads).  The issuing roles run WARP-UNIFORM code with one elected lane
// (ptx.cuh::elect_one): inside an `if (lane == 0)` region the compiler cannot keep descriptors in uniform registers
// and wraps every tcgen05.mma / TMA instruction in an ELECT + R2UR waterfall of ~100 cycles, which starved the MMA
// issue next to sixteen epilogue warps crunching GELU in the same sub-partitions.
//   warp 16  : TMA producer (whole warp walks the loop, one elected lane issues)
//   warp 17  : MMA issuer   (same; with CTA pairs only the leader CTA's warp issues)
//   warp 18  : TMEM allocator / deallocator
//   warp 19  : idle
//   warps 0-15: epilogue, two groups of eight warps.  Warp w may only read TMEM lanes 32*(w%4)..+31, so a group
//              has two warps per lane quarter and each takes half of the slab's columns (64 B of every row).
//              A group owns one 16 KB staging slab (128 rows x 128 B) and walks the tile's column slabs
//              g, g+2, ...; one thread per group issues the bulk tensor stores.  Four epilogue warps per SM
//              sub-partition keep the TMEM-load / MUFU / store latencies of the K=768 GEMMs under the MMA time.
//
// The epilogues are the ones the DINOv3 block needs (reference: HF modeling_dinov3_vit.py:305-311 QKV bias,
// :385-386 up_proj+GELU, :440-441 / :447-448 LayerScale+residual (LayerScale is folded into W and bias on
// the host), :71-92 patch-embedding rows interleaved behind the CLS/register prefix).  The residual update
// h += acc + bias never loads h into the SM: the slab is reduced into global memory by the TMA unit
// (cp.reduce.async.bulk.tensor .add.f32), so the read-modify-write happens in L2.
#pragma once
#include "ptx.cuh"

namespace cbas {

enum GemmEpilogue : int {
    EPI_BIAS_BF16 = 0,       // out_bf16[m,n] = acc + bias[n]
    EPI_BIAS_GELU_BF16 = 1,  // out_bf16[m,n] = gelu_erf(acc + bias[n])
    EPI_RESID_F32 = 2,       // resid_f32[m,n] += acc + bias[n]          (residual stream, in place)
    EPI_PATCH_F32 = 3,       // resid_f32[row_map(m),n] = acc + bias[n]  (patch rows behind the prefix tokens)
    EPI_BIAS_F32 = 4,        // out_f32[m,n] = acc + bias[n]
    EPI_BIAS_GELU_F32 = 5,   // out_f32[m,n] = gelu_erf(acc + bias[n])
    EPI_BIAS_BF16_VF16 = 6,  // EPI_BIAS_BF16, but columns >= f16_from are stored as IEEE f16 (the V third of QKV)
};

struct GemmParams {
    int M, N, K;
    const float* bias;  // [N] fp32 (may be null)
    void* out;          // bf16 or fp32, row stride ldo elements
    int ldo;
    // EPI_PATCH_F32: A row m = frame*rows_in + p  ->  out row frame*rows_out + prefix + p
    int rows_in, rows_out, prefix;
    int f16_from;  // EPI_BIAS_BF16_VF16: first column stored as f16 (multiple of 64)
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 16;
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_SLAB_BYTES = GEMM_BLOCK_M * 128;  // 128 rows x 128 B

__host__ __device__ constexpr bool gemm_epi_out_bf16(int epi) {
    return epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_BF16_VF16;
}
__host__ __device__ constexpr bool gemm_epi_staged(int epi) { return epi != EPI_PATCH_F32; }
__host__ __device__ constexpr bool gemm_epi_double_stage(int epi) { return epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_GELU_F32; }
__host__ __device__ constexpr int gemm_slab_cols(int epi) { return gemm_epi_out_bf16(epi) ? 64 : 32; }

// CG = 1: one CTA computes a 128 x BLOCK_N tile.  CG = 2: a CTA pair (cluster of two SMs, tcgen05 cta_group::2)
// computes a 256 x BLOCK_N tile; each CTA loads its own 128 A rows but only HALF of the B rows, which cuts the
// L2 -> shared-memory operand traffic per FLOP by a third - the 1-CTA mainloop is bound by exactly that traffic.
// kDoubleStage: two staging slabs per epilogue group (the bulk store of slab s overlaps the math of slab s+1) at
// the price of one mainloop stage - worth it only for the ALU-heavy GELU epilogues.
template <int BLOCK_N, int CG = 1, bool kDoubleStage = false>
struct GemmCfg {
    static constexpr int kStageA = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KB
    static constexpr int kStageB = (BLOCK_N / CG) * GEMM_BLOCK_K * 2;
    static constexpr int kStage = kStageA + kStageB;
    static constexpr int kSlabsPerGroup = kDoubleStage ? 2 : 1;
    static constexpr int kStaging = 2 * kSlabsPerGroup * GEMM_SLAB_BYTES;
    static constexpr int kBudget = 232448 - 1024 - 256 - kStaging;
    static constexpr int kStages = kBudget / kStage > 8 ? 8 : kBudget / kStage;
    static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStage + kStaging + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BLOCK_N, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, CG, gemm_epi_double_stage(EPI)>;
    static_assert(CG == 1 || CG == 2, "cta_group");
    static_assert((BLOCK_N / CG) % 8 == 0 && BLOCK_N % 16 == 0, "UMMA N");
    constexpr int kStages = Cfg::kStages;
    constexpr bool kStaged = gemm_epi_staged(EPI);
    constexpr bool kOutBf16 = gemm_epi_out_bf16(EPI);
    constexpr int kSlabCols = gemm_slab_cols(EPI);
    constexpr int kSlabs = BLOCK_N / kSlabCols;
    static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256, "BLOCK_N");
    static_assert(kSlabs >= 2, "each epilogue group needs at least one slab");

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kStageA;
    uint8_t* smem_stage = smem + kStages * Cfg::kStage;  // 4 x 16 KB, 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStaging);
    uint64_t* full_bar = bars;                      // [kStages] TMA -> MMA
    uint64_t* empty_bar = bars + kStages;           // [kStages] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kStages;       // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    constexpr int kTileM = GEMM_BLOCK_M * CG;
    const int m_blocks = (p.M + kTileM - 1) / kTileM;
    const int n_blocks = p.N / BLOCK_N;
    const int num_tiles = m_blocks * n_blocks;
    const int k_blocks = p.K / GEMM_BLOCK_K;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader of the pair
    const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

    if (CG == 2) cluster_sync_all();  // both CTAs resident before the pair-wide TMEM allocation
    constexpr int kProducerWarp = GEMM_EPI_WARPS, kMmaWarp = GEMM_EPI_WARPS + 1, kAllocWarp = GEMM_EPI_WARPS + 2;
    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (kStaged) tma_prefetch_desc(&tmap_out);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], CG);  // pair: one arrival per CTA, all on the leader's barrier
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], CG * GEMM_EPI_WARPS);  // pair: both CTAs' epilogue warps, leader's barrier
        }
        fence_mbar_init();
    }
    if (warp == kAllocWarp) {
        if (CG == 2) { tmem_alloc_pair(tmem_ptr_smem, Cfg::kTmemCols); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols); tmem_relinquish(); }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == kProducerWarp) {
        // ---------------------------------------------------------------- TMA producer (whole warp, one elected lane)
        uint32_t stage = 0, phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
            const int a_row = (m_blk * CG + (int)cta_rank) * GEMM_BLOCK_M;
            const int b_row = n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / CG);
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
                        tma_load_2d(smem_a + stage * Cfg::kStageA, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, a_row);
                        tma_load_2d(smem_b + stage * Cfg::kStageB, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, b_row);
                    } else {
                        // the leader's barrier counts the bytes of BOTH CTAs' loads
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStage);
                        else mbar_arrive_cluster(&full_bar[stage], 0);
                        tma_load_2d_pair(smem_a + stage * Cfg::kStageA, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K,
                                         a_row);
                        tma_load_2d_pair(smem_b + stage * Cfg::kStageB, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K,
                                         b_row);
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kMmaWarp) {
        if (cta_rank == 0) {
            // ------------------------------------------------------------ MMA issuer (pair: leader CTA only)
            // The whole warp runs the loop and one elected lane issues: with warp-uniform control flow the
            // descriptors live in uniform registers (see elect_one in ptx.cuh).
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BLOCK_N);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t a_u = smem_u32(smem_a), b_u = smem_u32(smem_b);
            uint32_t stage = 0, phase = 0;
            int iter = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++iter) {
                const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_u + acc * BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a_u + stage * Cfg::kStageA);
                    const uint64_t db = umma_desc_sw128(b_u + stage * Cfg::kStageB);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                            // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the (>>4) address field
                            if (CG == 1) umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            else umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                        if (CG == 1) umma_commit(&empty_bar[stage]); else umma_commit_pair(&empty_bar[stage], 3);
                        if (kb == k_blocks - 1) {
                            if (CG == 1) umma_commit(&tmem_full[acc]); else umma_commit_pair(&tmem_full[acc], 3);
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < GEMM_EPI_WARPS) {
        // ---------------------------------------------------------------- epilogue
        const int ew = warp;               // 0..15
        const int group = ew >> 3;         // 0 or 1: which staging slab / which slabs of the tile
        const int sub = (ew >> 2) & 1;     // which half of a slab's columns
        const int quarter = warp & 3;      // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
        const int row_in_tile = quarter * 32 + lane;
        const bool issuer_warp = (ew & 7) == 0;  // one bulk-store issuing warp per group (its elected lane issues)
        constexpr int kHalf = kSlabCols / 2;             // columns per thread per slab: 32 (bf16) or 16 (fp32)
        uint8_t* group_slabs = smem_stage + group * Cfg::kSlabsPerGroup * GEMM_SLAB_BYTES;
        const int swz = row_in_tile & 7;
        uint32_t sbuf = 0;  // which of the group's two staging slabs the next slab uses
        int iter = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++iter) {
            const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
            const int m_base = (m_blk * CG + (int)cta_rank) * GEMM_BLOCK_M;  // first output row of this CTA
            const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr_row = tmem_base + acc * BLOCK_N + (uint32_t(quarter * 32) << 16);

            if constexpr (kStaged) {
#pragma unroll 1
                for (int s = group; s < kSlabs; s += 2) {
                    const int c0 = s * kSlabCols + sub * kHalf;  // this thread's first column inside the tile
                    const int n0 = n_blk * BLOCK_N + c0;         // ... and in the matrix
                    float x[kHalf];
                    if constexpr (kHalf == 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr_row + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
                    } else {
                        uint32_t v[16];
                        tmem_ld_32x16(taddr_row + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(v[j]);
                    }
                    if (s + 2 >= kSlabs) {
                        // last slab of this tile for this warp: the accumulator can go back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cluster(&tmem_empty[acc], 0);
                        }
                    }
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < kHalf; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                            x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
                        }
                    }
                    if (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_F32) {
#pragma unroll
                        for (int j = 0; j < kHalf; ++j) x[j] = gelu_erf_fast<EPI == EPI_BIAS_GELU_BF16 ? 3 : 5>(x[j]);
                    }
                    // the bulk store issued from THIS slab buffer two slabs ago must have finished reading it
                    uint8_t* slab = group_slabs + sbuf * GEMM_SLAB_BYTES;
                    uint8_t* slab_row = slab + row_in_tile * 128;
                    if (Cfg::kSlabsPerGroup == 2) {
                        sbuf ^= 1;
                        if (issuer_warp && elect_one()) tma_wait_group_read<1>();
                    } else {
                        if (issuer_warp && elect_one()) tma_wait_group_read<0>();
                    }
                    named_bar_sync(1 + group, 256);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {  // this thread's 4 of the row's 8 16-byte chunks, XOR-swizzled like TMA
                        uint4 q;
                        if constexpr (kOutBf16) {
                            if (EPI == EPI_BIAS_BF16_VF16 && n0 >= p.f16_from) {
                                q.x = pack_f16(x[8 * j + 0], x[8 * j + 1]);
                                q.y = pack_f16(x[8 * j + 2], x[8 * j + 3]);
                                q.z = pack_f16(x[8 * j + 4], x[8 * j + 5]);
                                q.w = pack_f16(x[8 * j + 6], x[8 * j + 7]);
                            } else {
                                q.x = pack_bf16(x[8 * j + 0], x[8 * j + 1]);
                                q.y = pack_bf16(x[8 * j + 2], x[8 * j + 3]);
                                q.z = pack_bf16(x[8 * j + 4], x[8 * j + 5]);
                                q.w = pack_bf16(x[8 * j + 6], x[8 * j + 7]);
                            }
                        } else {
                            q.x = __float_as_uint(x[4 * j + 0]);
                            q.y = __float_as_uint(x[4 * j + 1]);
                            q.z = __float_as_uint(x[4 * j + 2]);
                            q.w = __float_as_uint(x[4 * j + 3]);
                        }
                        *reinterpret_cast<uint4*>(slab_row + (((sub * 4 + j) ^ swz) << 4)) = q;
                    }
                    fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async proxy)
                    named_bar_sync(1 + group, 256);
                    if (issuer_warp) {
                        const int ns = n_blk * BLOCK_N + s * kSlabCols;
                        if (elect_one()) {
                            if (EPI == EPI_RESID_F32) tma_reduce_add_2d(&tmap_out, slab, ns, m_base);
                            else tma_store_2d(&tmap_out, slab, ns, m_base);
                            tma_commit_group();
                        }
                    }
                }
            } else {
                // direct row-per-thread stores (patch-embedding rows are re-mapped, 0.7 % of the step)
                constexpr int kColsPerWarp = BLOCK_N / 4;
                const int col0 = (group * 2 + sub) * kColsPerWarp;
                const int row = m_base + row_in_tile;
                const bool row_ok = row < p.M;
                const int f = row / p.rows_in;
                const long long out_row = (long long)f * p.rows_out + p.prefix + (row - f * p.rows_in);
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp; c += 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(taddr_row + col0 + c, v);
                    tmem_ld_wait();
                    const int n0 = n_blk * BLOCK_N + col0 + c;
                    if (row_ok) {
                        float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                            *reinterpret_cast<float4*>(o + j) =
                                make_float4(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y,
                                            __uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cluster(&tmem_empty[acc], 0);
                }
            }
        }
        if (kStaged && issuer_warp && elect_one()) tma_wait_group<0>();  // all bulk stores of this CTA have landed
    }

    __syncwarp();
    tc_fence_before();
    // pair: neither CTA may leave while the other can still touch its shared memory, barriers or TMEM
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace cbas
