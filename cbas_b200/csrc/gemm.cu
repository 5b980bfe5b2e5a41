// Host launcher for the tcgen05 GEMM: builds the TMA tensor maps (driver entry point fetched through the
// runtime, so the library does not link libcuda) and dispatches on tile width and epilogue.
#include "common.h"

#include <atomic>
#include "gemm_tcgen05.cuh"

#include <mutex>

namespace cbas {

namespace {

thread_local int g_force_cg = 0;  // 0 = automatic; 1 / 2 force the CTA-group size (tests, A/B timing); per host thread

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Row-major matrix [rows, cols] with row pitch ld elements; box = box_cols x box_rows whose inner extent is
// 128 bytes (64 bf16 or 32 fp32), 128-byte swizzle.  Used for the K-major operands and for the output slabs.
int make_tmap(CUtensorMap* map, const void* base, bool f32, int rows, int cols, int ld, int box_cols, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    const size_t es = f32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

template <int BLOCK_N, int EPI, int CG, bool LNC = false>
int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, CG, gemm_epi_double_stage(EPI), gemm_epi_ln_producer(EPI) ? gemm_ln_bufs(EPI) : 0, LNC>;
    auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI, CG, LNC>;
    // the opt-in is a per-DEVICE attribute: one flag per instantiation and device ordinal
    static DeviceSmemOptIn optin;
    CBAS_CHECK(optin.ensure(kern, Cfg::kSmemBytes));
    CUtensorMap tout;
    if (gemm_epi_ln_producer(EPI)) {
        if (!p.ln_in || !p.ln_out || !p.hb) return fail("LayerNorm-producer GEMM needs statistics in/out and the bf16 copy");
        if (4 * (p.N / BLOCK_N) > LN_STAT_SLOTS) return fail("LayerNorm-producer GEMM: N too wide for the statistics slots");
        if ((reinterpret_cast<uintptr_t>(p.hb) & 31) || (p.ldhb % 16) || (reinterpret_cast<uintptr_t>(p.ln_in) & 15))
            return fail("LayerNorm-producer GEMM: hb must be 32-byte aligned (row pitch too), statistics 16-byte aligned");
    } else if (gemm_epi_out_bf16(EPI) && p.ln_in) {
        if (!p.ln_c1 || !p.bias) return fail("LayerNorm-consumer GEMM needs c1 and c2");
        if (reinterpret_cast<uintptr_t>(p.ln_in) & 15) return fail("LayerNorm statistics must be 16-byte aligned");
    }
    {
        const bool f32 = !gemm_epi_out_bf16(EPI);
        if ((reinterpret_cast<uintptr_t>(p.out) & 15) || (p.ldo % (f32 ? 4 : 8)))
            return fail("GEMM output must be 16-byte aligned with a 16-byte multiple row pitch");
        if (EPI == EPI_PATCH_F32) {
            if (p.rows_in <= 0 || p.rows_out < p.rows_in + p.prefix) return fail("patch GEMM: bad row mapping");
            tout = ta;  // unused: the patch epilogue copies its staged slabs out with plain stores (rows are re-mapped)
        } else if (int rc = make_tmap(&tout, p.out, f32, p.M, p.N, p.ldo, gemm_slab_cols(EPI), GEMM_BLOCK_M)) {
            return rc;
        }
    }
    const int m_blocks = (p.M + GEMM_BLOCK_M * CG - 1) / (GEMM_BLOCK_M * CG);
    const int tiles = m_blocks * (p.N / BLOCK_N);
    const int max_clusters = sm_count() / CG;
    const int grid = CG * (tiles < max_clusters ? tiles : max_clusters);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1 + pdl_attr(&attr[1]);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, p);
    count_launch();
    if (err != cudaSuccess) return check_cuda(err, "gemm_tcgen05_kernel launch");
    return check_cuda(cudaGetLastError(), "gemm_tcgen05_kernel launch");
}

template <int BLOCK_N, int CG>
int launch_bn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const GemmParams& p, int epi,
              cudaStream_t stream) {
    CUtensorMap ta, tb;
    if (int rc = make_tmap(&ta, A, false, p.M, p.a_wrap ? p.a_wrap : p.K, lda, GEMM_BLOCK_K, GEMM_BLOCK_M)) return rc;
    if (int rc = make_tmap(&tb, W, false, p.N, p.K, ldw, GEMM_BLOCK_K, BLOCK_N / CG)) return rc;
    switch (epi) {
        // bf16-output epilogues: p.ln_in selects the fused-LayerNorm consumer instantiation (its own shared-memory plan)
        case EPI_BIAS_BF16:
            if (p.ln_in) return launch_one<BLOCK_N, EPI_BIAS_BF16, CG, true>(ta, tb, p, stream);
            return launch_one<BLOCK_N, EPI_BIAS_BF16, CG>(ta, tb, p, stream);
        case EPI_BIAS_GELU_BF16:
            if (p.ln_in) return launch_one<BLOCK_N, EPI_BIAS_GELU_BF16, CG, true>(ta, tb, p, stream);
            return launch_one<BLOCK_N, EPI_BIAS_GELU_BF16, CG>(ta, tb, p, stream);
        case EPI_RESID_F32: return launch_one<BLOCK_N, EPI_RESID_F32, CG>(ta, tb, p, stream);
        case EPI_PATCH_F32: return launch_one<BLOCK_N, EPI_PATCH_F32, CG>(ta, tb, p, stream);
        case EPI_BIAS_F32: return launch_one<BLOCK_N, EPI_BIAS_F32, CG>(ta, tb, p, stream);
        case EPI_BIAS_GELU_F32: return launch_one<BLOCK_N, EPI_BIAS_GELU_F32, CG>(ta, tb, p, stream);
        case EPI_BIAS_BF16_VF16:
            if (p.ln_in) return launch_one<BLOCK_N, EPI_BIAS_BF16_VF16, CG, true>(ta, tb, p, stream);
            return launch_one<BLOCK_N, EPI_BIAS_BF16_VF16, CG>(ta, tb, p, stream);
        // long mainloop: one old-h slab buffer per epilogue group and a deeper operand ring; short: prefetched slabs
        case EPI_RESID_LN_F32:
        case EPI_RESID_LN1_F32:
            if (p.K > 1536) return launch_one<BLOCK_N, EPI_RESID_LN1_F32, CG>(ta, tb, p, stream);
            return launch_one<BLOCK_N, EPI_RESID_LN_F32, CG>(ta, tb, p, stream);
    }
    return fail("unknown GEMM epilogue " + std::to_string(epi));
}


}  // namespace

int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const GemmParams& p, int epi,
                cudaStream_t stream, int prof_tag) {
    if (p.M <= 0) return 0;
    ProfScope prof(prof_tag, stream);
    if (p.K % GEMM_BLOCK_K != 0 || p.K <= 0) return fail("GEMM K must be a positive multiple of 64");
    if (p.a_wrap && (p.a_wrap % GEMM_BLOCK_K != 0 || p.a_wrap <= 0 || 2 * p.K != 3 * p.a_wrap))
        return fail("GEMM a_wrap must be a multiple of 64 with K = 1.5 * a_wrap");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (lda % 8) || (ldw % 8))
        return fail("GEMM operands must be 16-byte aligned with row pitch a multiple of 8 elements");
    // CTA pairs (256-row tiles) whenever there is enough work to fill the chip with them
    const bool pair = g_force_cg ? g_force_cg == 2 : p.M >= 4096;
    if (pair) {
        if (p.N % 256 == 0) return launch_bn<256, 2>(A, lda, W, ldw, p, epi, stream);
        if (p.N % 192 == 0) return launch_bn<192, 2>(A, lda, W, ldw, p, epi, stream);
        if (p.N % 128 == 0) return launch_bn<128, 2>(A, lda, W, ldw, p, epi, stream);
    } else {
        if (p.N % 256 == 0) return launch_bn<256, 1>(A, lda, W, ldw, p, epi, stream);
        if (p.N % 192 == 0) return launch_bn<192, 1>(A, lda, W, ldw, p, epi, stream);
        if (p.N % 128 == 0) return launch_bn<128, 1>(A, lda, W, ldw, p, epi, stream);
    }
    return fail("GEMM N must be a multiple of 128 (got " + std::to_string(p.N) + ")");
}

void set_gemm_cta_group(int cg) { g_force_cg = cg; }

int make_tmap_2d(CUtensorMap* map, const void* base, bool f32, int rows, int cols, int ld, int box_cols, int box_rows) {
    return make_tmap(map, base, f32, rows, cols, ld, box_cols, box_rows);
}

// bf16 [d2][d1][d0] tensor (d0 contiguous, row pitch ld elements, slab pitch d1*ld), box box0 x box1 x 1, SWIZZLE_128B
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, int d0, int d1, int d2, int ld, int box0, int box1) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)d1};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string((int)r));
    return 0;
}


// fp32 [d2][d1][d0] tensor (d0 contiguous, pitches in BYTES), box box0 x box1 x box2 with a 128-byte inner extent
// (box0 = 32 floats), SWIZZLE_128B; out-of-range elements read as zero
int make_tmap_3d_f32(CUtensorMap* map, const void* base, long long d0, long long d1, long long d2, long long pitch1,
                     long long pitch2, int box0, int box1, int box2) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)pitch1, (cuuint64_t)pitch2};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (3-D fp32) failed with CUresult " + std::to_string((int)r));
    return 0;
}

}  // namespace cbas
