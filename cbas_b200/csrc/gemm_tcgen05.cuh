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
//   warp 19  : fused-LayerNorm GEMMs only: per tile, one tile ahead of the epilogue, turns the 128 rows' partial sums
//              into {rstd, -mean * rstd} (consumer) or the new row shift (producer) in shared memory and pulls the
//              tile's column constants into L1 (the epilogue warps are the busiest part of these kernels and must not
//              wait for global loads); else idle
//   warps 0-15: epilogue, two groups of eight warps.  Warp w may only read TMEM lanes 32*(w%4)..+31, so a group
//              has two warps per lane quarter and each takes half of the slab's columns (64 B of every row).
//              A group owns one 16 KB staging slab (128 rows x 128 B) and walks the tile's column slabs
//              g, g+2, ...; one thread per group issues the bulk tensor stores.  Four epilogue warps per SM
//              sub-partition keep the TMEM-load / MUFU / store latencies of the K=768 GEMMs under the MMA time.
//
// The epilogues are the ones the DINOv3 block needs (reference: HF modeling_dinov3_vit.py:305-311 QKV bias,
// :385-386 up_proj+GELU, :440-441 / :447-448 LayerScale+residual (LayerScale is folded into W and bias on
// the host), :71-92 patch-embedding rows interleaved behind the CLS/register prefix).  The plain residual update
// h += acc + bias never loads h into the SM: the slab is reduced into global memory by the TMA unit
// (cp.reduce.async.bulk.tensor .add.f32), so the read-modify-write happens in L2.
//
// Fused LayerNorm (HF modeling_dinov3_vit.py:433,445: norm1 / norm2 in front of the QKV and up projections).  There is
// no LayerNorm kernel between the GEMMs of a block; the normalisation is split over the GEMM that PRODUCES the
// residual stream and the one that CONSUMES it:
//   producer (EPI_RESID_LN_F32: proj, down): loads the old fp32 slab of h by TMA (prefetched one slab ahead), adds
//       acc + bias, stores the new fp32 slab, and ALSO stores hb = bf16(h - s_m), a copy shifted by the row's previous
//       mean s_m (so the bf16 rounding acts on centred values even when a row's mean dwarfs its spread), plus per-row
//       partial sums  sum(y), sum(y^2)  of y = h - s_m, one slot per (column tile, epilogue thread of the row): no
//       atomics, no cross-warp exchange, and the result is bit-reproducible.
//   consumer (GemmParams::ln_in set: QKV, up): A = hb, W' = W * gamma (folded on the host).  With mu = mean(y),
//       r = rsqrt(var(y) + eps):   LN(h) W^T + b  =  r * (hb W'^T)  -  r * mu * c1  +  c2,
//       c1[n] = sum_k W'[n,k],  c2[n] = sum_k W[n,k] beta[k] + b[n]   - two FMAs per element in the epilogue.
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
    EPI_RESID_LN_F32 = 7,    // resid_f32 += acc + bias, plus the shifted bf16 copy and row statistics (LayerNorm producer);
                             // old-h slabs double-buffered per epilogue group: for short mainloops (proj, K = D)
    EPI_RESID_LN1_F32 = 8,   // the same with one slab buffer per group and a deeper mainloop ring: for long K (down)
};
__host__ __device__ constexpr bool gemm_epi_ln_producer(int epi) { return epi == EPI_RESID_LN_F32 || epi == EPI_RESID_LN1_F32; }

// Row statistics of the residual stream for the fused LayerNorm: LN_STAT_FLOATS floats per row,
//   [0,16)  partial sums of y = h - shift, slot = 4 * column_tile + 2 * epilogue_group + column_half (unused: zero)
//   [16,32) partial sums of y^2, same slots          [32] shift          [33,36) padding
constexpr int LN_STAT_SLOTS = 16;
constexpr int LN_STAT_FLOATS = 36;

struct GemmParams {
    int M, N, K;
    const float* bias;  // [N] fp32 (may be null)
    void* out;          // bf16 or fp32, row stride ldo elements
    int ldo;
    // EPI_PATCH_F32: A row m = frame*rows_in + p  ->  out row frame*rows_out + prefix + p
    int rows_in, rows_out, prefix;
    int f16_from;  // EPI_BIAS_BF16_VF16: first column stored as f16 (multiple of 64)
    // ---- fused LayerNorm (see the header comment)
    // consumer (bf16-output epilogues, ln_in != null): A is the shifted copy hb, `bias` holds c2, ln_c1 the column sums
    // producer (EPI_RESID_LN_F32): ln_in = statistics of the residual stream BEFORE this update (its exact row mean
    //   becomes the new shift), ln_out = statistics after it, hb = shifted bf16 copy [M, N] with row pitch ldhb
    const float* ln_in;   // row r at ln_in + r * ln_in_stride * LN_STAT_FLOATS
    int ln_in_stride;
    float* ln_out;
    int ln_out_stride;
    const float* ln_c1;   // [N] fp32
    float ln_inv_dim;     // 1 / D (the normalised width; equals 1 / K for a consumer, 1 / N for a producer)
    float ln_eps;
    void* hb;
    int ldhb;
    // split-bf16 GEMMs (head.cu): A holds [hi | lo] (a_wrap columns) while K = 1.5 * a_wrap; K blocks at or past
    // a_wrap read A from column k - a_wrap again (hi * W_lo).  0 = A is K wide.
    int a_wrap;
    // walk the row blocks from the last to the first (encoder.cu: consecutive kernels alternate direction, so each one
    // starts on the rows its predecessor touched last, which are still in L2).  Not supported by the LN producers.
    int reverse;
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
__host__ __device__ constexpr bool gemm_epi_staged(int) { return true; }  // every epilogue leaves through a staged TMA store
__host__ __device__ constexpr bool gemm_epi_double_stage(int epi) { return epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_GELU_F32; }
__host__ __device__ constexpr int gemm_slab_cols(int epi) { return gemm_epi_out_bf16(epi) ? 64 : 32; }
// LayerNorm producer: per epilogue group one or two fp32 slabs (old h in by TMA, new h out, in place); the shifted
// bf16 copy leaves the registers through 32-byte global stores (one full sector per thread), so it needs no staging
__host__ __device__ constexpr int gemm_ln_bufs(int epi) { return epi == EPI_RESID_LN_F32 ? 2 : 1; }

// CG = 1: one CTA computes a 128 x BLOCK_N tile.  CG = 2: a CTA pair (cluster of two SMs, tcgen05 cta_group::2)
// computes a 256 x BLOCK_N tile; each CTA loads its own 128 A rows but only HALF of the B rows, which cuts the
// L2 -> shared-memory operand traffic per FLOP by a third - the 1-CTA mainloop is bound by exactly that traffic.
// kDoubleStage: two staging slabs per epilogue group (the bulk store of slab s overlaps the math of slab s+1) at
// the price of one mainloop stage - worth it only for the ALU-heavy GELU epilogues.
template <int BLOCK_N, int CG = 1, bool kDoubleStage = false, int kLnBufs = 0, bool kLnConsumer = false>
struct GemmCfg {
    static constexpr int kStageA = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KB
    static constexpr int kStageB = (BLOCK_N / CG) * GEMM_BLOCK_K * 2;
    static constexpr int kStage = kStageA + kStageB;
    static constexpr int kSlabsPerGroup = kDoubleStage ? 2 : 1;
    static constexpr int kStaging = 2 * (kLnBufs ? kLnBufs : kSlabsPerGroup) * GEMM_SLAB_BYTES;
    // LayerNorm consumers: per-row scale / offset of the current tile [128] float2, then the per-column constants
    // c1 | c2 of the current and the next tile, [2][2][BLOCK_N] floats (written by the helper warp, read by every
    // epilogue thread as shared-memory broadcasts: as global loads they were the epilogue's longest stall, the ~28 KB
    // of L1 left next to 227 KB of shared memory does not keep them)
    static constexpr int kSide = kLnConsumer ? 1024 + 2 * 2 * BLOCK_N * 4 : 1024;
    static constexpr int kBudget = 232448 - 1024 - 256 - kSide - kStaging;
#ifdef GEMM_FORCE_STAGES  // A/B builds only (tools/build_ref_lib.py WORKTREE -DGEMM_FORCE_STAGES=4)
    static constexpr int kStages = kBudget / kStage > GEMM_FORCE_STAGES ? GEMM_FORCE_STAGES : kBudget / kStage;
#else
    static constexpr int kStages = kBudget / kStage > 8 ? 8 : kBudget / kStage;
#endif
    static_assert(kStages >= 2, "mainloop needs at least two stages");
    static constexpr int kTmemCols = (2 * BLOCK_N <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStage + kStaging + kSide + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int BLOCK_N, int EPI, int CG, bool LNC = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
    constexpr bool kLnProducer = gemm_epi_ln_producer(EPI);
    constexpr int kLnBufs = kLnProducer ? gemm_ln_bufs(EPI) : 0;
    constexpr bool kLnConsumer = LNC;  // fused LayerNorm, consumer side (bf16-output epilogues, p.ln_in set)
    static_assert(!LNC || gemm_epi_out_bf16(EPI), "LayerNorm-consumer epilogues write bf16");
    using Cfg = GemmCfg<BLOCK_N, CG, gemm_epi_double_stage(EPI), kLnBufs, kLnConsumer>;
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
    float2* rowc = reinterpret_cast<float2*>(smem_stage + Cfg::kStaging);  // [128] {rstd, -mean * rstd} (LN consumer)
    float* colc = reinterpret_cast<float*>(smem_stage + Cfg::kStaging + 1024);  // [2 tile parity][c1 | c2][BLOCK_N] (LN consumer)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStaging + Cfg::kSide);
    uint64_t* full_bar = bars;                      // [kStages] TMA -> MMA
    uint64_t* empty_bar = bars + kStages;           // [kStages] MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kStages;       // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
    uint64_t* ln_full = bars + 2 * kStages + 4;     // [2 groups][2 buffers] old-h slab landed (LayerNorm producer)
    uint64_t* rowc_full = bars + 2 * kStages + 8;   // helper warp -> epilogue: rowc holds this tile's values
    uint64_t* rowc_empty = bars + 2 * kStages + 9;  // epilogue -> helper warp: every epilogue warp has read them
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 10);

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
        for (int i = 0; i < 4; ++i) mbar_init(&ln_full[i], 1);
        mbar_init(rowc_full, 1);
        mbar_init(rowc_empty, GEMM_EPI_WARPS);
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
    // everything above overlapped the previous kernel's tail (programmatic dependent launch); its results are needed
    // from here on.  The next kernel in the stream may start its own prologue as SMs free up.
    pdl_wait();
    pdl_trigger();

    if (warp == kProducerWarp) {
        // ---------------------------------------------------------------- TMA producer (whole warp, one elected lane)
        uint32_t stage = 0, phase = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int m_blk = p.reverse ? m_blocks - 1 - tile / n_blocks : tile / n_blocks, n_blk = tile % n_blocks;
            const int a_row = (m_blk * CG + (int)cta_rank) * GEMM_BLOCK_M;
            const int b_row = n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / CG);
            for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int a_col = p.a_wrap && kb * GEMM_BLOCK_K >= p.a_wrap ? kb * GEMM_BLOCK_K - p.a_wrap : kb * GEMM_BLOCK_K;
                if (elect_one()) {
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStage);
                        tma_load_2d(smem_a + stage * Cfg::kStageA, &tmap_a, &full_bar[stage], a_col, a_row);
                        tma_load_2d(smem_b + stage * Cfg::kStageB, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, b_row);
                    } else {
                        // the leader's barrier counts the bytes of BOTH CTAs' loads
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStage);
                        else mbar_arrive_cluster(&full_bar[stage], 0);
                        tma_load_2d_pair(smem_a + stage * Cfg::kStageA, &tmap_a, &full_bar[stage], a_col, a_row);
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
    } else if (warp == GEMM_EPI_WARPS + 3) {
        // ---------------------------------------------------------------- LayerNorm helper (consumer and producer)
        if constexpr (kLnConsumer || kLnProducer) {
            {
                int iter = 0;
                for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++iter) {
                    const int m_blk = p.reverse ? m_blocks - 1 - tile / n_blocks : tile / n_blocks, n_blk = tile % n_blocks;
                    const int m_base = (m_blk * CG + (int)cta_rank) * GEMM_BLOCK_M;
                    // this tile's per-column constants (BLOCK_N floats each): first touch by this warp, L1 hits for the
                    // epilogue warps
                    float4 cc[kLnConsumer ? BLOCK_N / 64 : 1];  // this lane's share of c1 | c2: 2 * BLOCK_N floats over 32 lanes
                    if constexpr (kLnConsumer) {
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i) {
                            const int idx = (i * 32 + lane) * 4;  // 0 .. 2 * BLOCK_N
                            const float* src = idx < BLOCK_N ? p.ln_c1 + n_blk * BLOCK_N + idx
                                                             : p.bias + n_blk * BLOCK_N + (idx - BLOCK_N);
                            cc[i] = __ldg(reinterpret_cast<const float4*>(src));
                        }
                    } else if (lane * 32 < BLOCK_N) {
                        if (p.bias) prefetch_l1(p.bias + n_blk * BLOCK_N + lane * 32);
                    }
                    float2 rc[GEMM_BLOCK_M / 32];
#pragma unroll
                    for (int r = 0; r < GEMM_BLOCK_M / 32; ++r) {
                        const int row = m_base + r * 32 + lane;
                        float t = 0.f, q = 0.f, s_old = 0.f;
                        if (row < p.M) {
                            const float4* st = reinterpret_cast<const float4*>(
                                p.ln_in + (size_t)row * p.ln_in_stride * LN_STAT_FLOATS);
#pragma unroll
                            for (int i = 0; i < LN_STAT_SLOTS / 4; ++i) {
                                const float4 a = ld_cg_f4(st + i);
                                t += (a.x + a.y) + (a.z + a.w);
                                if (!kLnProducer) {
                                    const float4 c = ld_cg_f4(st + LN_STAT_SLOTS / 4 + i);
                                    q += (c.x + c.y) + (c.z + c.w);
                                }
                            }
                            if (kLnProducer) s_old = ld_cg_f1(reinterpret_cast<const float*>(st) + 2 * LN_STAT_SLOTS);
                        }
                        const float mu = t * p.ln_inv_dim, ms = q * p.ln_inv_dim;
                        if (kLnProducer) {
                            // the row's new shift: its exact mean before this update (old shift + mean of the old y)
                            rc[r] = make_float2(s_old + mu, 0.f);
                        } else {
                            const float rstd = rsqrtf(fmaxf(fmaf(-mu, mu, ms), 0.f) + p.ln_eps);
                            rc[r] = make_float2(rstd, -mu * rstd);
                        }
                    }
                    mbar_wait(rowc_empty, (iter & 1) ^ 1);  // the epilogue has taken the previous tile's values
#pragma unroll
                    for (int r = 0; r < GEMM_BLOCK_M / 32; ++r) rowc[r * 32 + lane] = rc[r];
                    if constexpr (kLnConsumer) {
                        // every epilogue warp has started tile iter - 1, so none is still reading tile iter - 2's constants
                        float4* dst = reinterpret_cast<float4*>(colc + (iter & 1) * 2 * BLOCK_N);
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 64; ++i) dst[i * 32 + lane] = cc[i];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rowc_full);
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
        if constexpr (kLnProducer) {
            // -------------------------------------------------------- LayerNorm producer (see the header comment)
            uint8_t* hbuf = smem_stage + group * kLnBufs * GEMM_SLAB_BYTES;  // this group's fp32 slab(s), SWIZZLE_128B
            uint64_t* hfull = ln_full + 2 * group;
            uint32_t j = 0;  // slabs this group has handled
            if (kLnBufs == 2 && issuer_warp && first_tile < num_tiles) {
                // the group's first old-h slab is under way before the first accumulator is
                if (elect_one()) {
                    const int nm = ((first_tile / n_blocks) * CG + (int)cta_rank) * GEMM_BLOCK_M;
                    const int nn = (first_tile % n_blocks) * BLOCK_N + group * 32;
                    mbar_arrive_expect_tx(&hfull[0], GEMM_SLAB_BYTES);
                    tma_load_2d(hbuf, &tmap_out, &hfull[0], nn, nm);
                }
                __syncwarp();
            }
            int iter = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++iter) {
                const int m_blk = p.reverse ? m_blocks - 1 - tile / n_blocks : tile / n_blocks, n_blk = tile % n_blocks;
                const int m_base = (m_blk * CG + (int)cta_rank) * GEMM_BLOCK_M;
                const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
                const int row = m_base + row_in_tile;
                const bool row_ok = row < p.M;
                // new shift = exact mean of the row before this update, prepared one tile ahead by the helper warp
                mbar_wait(rowc_full, iter & 1);
                const float shift = rowc[row_in_tile].x;
                __syncwarp();
                if (lane == 0) mbar_arrive(rowc_empty);
                __nv_bfloat16* hb_row = reinterpret_cast<__nv_bfloat16*>(p.hb) + (size_t)(row_ok ? row : 0) * p.ldhb;
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr_row = tmem_base + acc * BLOCK_N + (uint32_t(quarter * 32) << 16);
                float psum = 0.f, psq = 0.f;
#pragma unroll 1
                for (int s = group; s < kSlabs; s += 2, ++j) {
                    const uint32_t buf = kLnBufs == 2 ? (j & 1) : 0;
                    const uint32_t parity = kLnBufs == 2 ? ((j >> 1) & 1) : (j & 1);
                    const bool last = s + 2 >= kSlabs;
                    if (issuer_warp) {
                        if (elect_one()) {
                            tma_wait_group_read<0>();  // the stores out of this group's buffers have read them
                            if (kLnBufs == 2) {
                                // prefetch the group's next slab (same tile, or the first one of its next tile)
                                int nt = tile, ns = s + 2;
                                if (ns >= kSlabs) { nt = tile + tile_step; ns = group; }
                                if (nt < num_tiles) {
                                    const int nm = ((nt / n_blocks) * CG + (int)cta_rank) * GEMM_BLOCK_M;
                                    const int nn = (nt % n_blocks) * BLOCK_N + ns * 32;
                                    mbar_arrive_expect_tx(&hfull[buf ^ 1], GEMM_SLAB_BYTES);
                                    tma_load_2d(hbuf + (buf ^ 1) * GEMM_SLAB_BYTES, &tmap_out, &hfull[buf ^ 1], nn, nm);
                                }
                            } else {
                                // long mainloop: the load of THIS slab hides behind the accumulator read below
                                mbar_arrive_expect_tx(&hfull[0], GEMM_SLAB_BYTES);
                                tma_load_2d(hbuf, &tmap_out, &hfull[0], n_blk * BLOCK_N + s * 32, m_base);
                            }
                        }
                        __syncwarp();
                    }
                    const int c0 = s * 32 + sub * 16;
                    const int n0 = n_blk * BLOCK_N + c0;
                    uint32_t v[16];
                    tmem_ld_32x16(taddr_row + c0, v);
                    tmem_ld_wait();
                    if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cluster(&tmem_empty[acc], 0);
                        }
                    }
                    float x[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]);
                    if (p.bias) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                            x[i] += b.x; x[i + 1] += b.y; x[i + 2] += b.z; x[i + 3] += b.w;
                        }
                    }
                    mbar_wait(&hfull[buf], parity);
                    uint8_t* hslab = hbuf + buf * GEMM_SLAB_BYTES;
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4* q = reinterpret_cast<float4*>(hslab + row_in_tile * 128 + (((sub * 4 + i) ^ swz) << 4));
                        float4 o = *q;
                        o.x += x[4 * i]; o.y += x[4 * i + 1]; o.z += x[4 * i + 2]; o.w += x[4 * i + 3];
                        *q = o;
                        const float y0 = o.x - shift, y1 = o.y - shift, y2 = o.z - shift, y3 = o.w - shift;
                        psum += (y0 + y1) + (y2 + y3);
                        psq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, psq))));
                        pk[2 * i] = pack_bf16(y0, y1);
                        pk[2 * i + 1] = pack_bf16(y2, y3);
                    }
                    fence_proxy_async();
                    named_bar_sync(1 + group, 256);
                    if (issuer_warp) {
                        if (elect_one()) {
                            tma_store_2d(&tmap_out, hslab, n_blk * BLOCK_N + s * 32, m_base);
                            tma_commit_group();
                        }
                    }
                    // the shifted bf16 copy: this thread's 16 columns are one 32-byte sector of the row
                    if (row_ok) st_global_v8(hb_row + n0, pk);
                }
                if (row_ok) {
                    // this thread's share of the row statistics, in its own slot (fixed order, no atomics)
                    float* so = p.ln_out + (size_t)row * p.ln_out_stride * LN_STAT_FLOATS;
                    const int slot = n_blk * 4 + group * 2 + sub;
                    so[slot] = psum;
                    so[LN_STAT_SLOTS + slot] = psq;
                    if (slot == 0) so[2 * LN_STAT_SLOTS] = shift;
                }
            }
        } else {
        uint32_t sbuf = 0;  // which of the group's two staging slabs the next slab uses
        int iter = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++iter) {
            const int m_blk = p.reverse ? m_blocks - 1 - tile / n_blocks : tile / n_blocks, n_blk = tile % n_blocks;
            const int m_base = (m_blk * CG + (int)cta_rank) * GEMM_BLOCK_M;  // first output row of this CTA
            const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
            // fused LayerNorm, consumer side: this row's rstd and -mean * rstd, prepared by the helper warp
            float ln_alpha = 1.f, ln_ndelta = 0.f;
            if constexpr (kLnConsumer) {
                {
                    mbar_wait(rowc_full, iter & 1);
                    const float2 rc = rowc[row_in_tile];
                    ln_alpha = rc.x; ln_ndelta = rc.y;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rowc_empty);
                }
            }
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
                    if constexpr (kLnConsumer) {
                        // LN(h) W^T + b = rstd * acc - rstd * mean * c1 + c2   (bias holds c2); c1 | c2 from shared memory
                        const float* cc1 = colc + (iter & 1) * 2 * BLOCK_N + c0;
#pragma unroll
                        for (int j = 0; j < kHalf; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(cc1 + BLOCK_N + j);
                            const float4 c = *reinterpret_cast<const float4*>(cc1 + j);
                            x[j] = fmaf(ln_alpha, x[j], fmaf(ln_ndelta, c.x, b.x));
                            x[j + 1] = fmaf(ln_alpha, x[j + 1], fmaf(ln_ndelta, c.y, b.y));
                            x[j + 2] = fmaf(ln_alpha, x[j + 2], fmaf(ln_ndelta, c.z, b.z));
                            x[j + 3] = fmaf(ln_alpha, x[j + 3], fmaf(ln_ndelta, c.w, b.w));
                        }
                    } else if (p.bias) {
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
                    if constexpr (EPI == EPI_PATCH_F32) {
                        // Patch rows sit behind each frame's prefix tokens, so a 128-row slab is not one box of the output
                        // (and TMA stores take no negative start coordinate to clip its head).  The group's 256 threads copy
                        // the slab out themselves, eight lanes per row: every store instruction writes four whole 128-byte
                        // lines, where the row-per-thread stores this replaces wrote 32 half sectors and were LSU-bound.
                        const int ns = n_blk * BLOCK_N + s * kSlabCols;
                        const int tl = (ew & 7) * 32 + lane;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int idx = tl + 256 * i, r = idx >> 3, c = idx & 7;
                            const int m = m_base + r;
                            if (m < p.M) {
                                const int f = m / p.rows_in;
                                const long long orow = (long long)f * p.rows_out + p.prefix + (m - f * p.rows_in);
                                const float4 v = *reinterpret_cast<const float4*>(slab + r * 128 + ((c ^ (r & 7)) << 4));
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + ns + c * 4) = v;
                            }
                        }
                        // (the barrier at the top of the group's next slab orders these reads before the slab is rewritten)
                    } else if (issuer_warp) {
                        const int ns = n_blk * BLOCK_N + s * kSlabCols;
                        if (elect_one()) {
                            if (EPI == EPI_RESID_F32) tma_reduce_add_2d(&tmap_out, slab, ns, m_base);
                            else tma_store_2d(&tmap_out, slab, ns, m_base);
                            tma_commit_group();
                        }
                    }
                }
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
