// Host launcher for the tcgen05 GEMM: builds the TMA tensor maps (driver entry point fetched through the
// runtime, so the library does not link libcuda) and dispatches on tile width and epilogue.
#include "common.h"
#include "gemm_tcgen05.cuh"

#include <mutex>

namespace cbas {

namespace {

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

// K-major bf16 matrix [rows, K] with row pitch ld elements; box = 64 (K) x box_rows, 128-byte swizzle.
int make_tmap(CUtensorMap* map, const __nv_bfloat16* base, int rows, int K, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

template <int BLOCK_N, int EPI>
int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    // exact-erf GELU is ALU-heavy: give that epilogue 8 warps (two column slices per TMEM lane quarter)
    constexpr int EW = 8;
    using Cfg = GemmCfg<BLOCK_N>;
    auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI, EW>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        CBAS_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        configured = true;
    }
    const int m_blocks = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int tiles = m_blocks * (p.N / BLOCK_N);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, 128 + 32 * EW, Cfg::kSmemBytes, stream>>>(ta, tb, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "gemm_tcgen05_kernel launch");
}

template <int BLOCK_N>
int launch_bn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const GemmParams& p, int epi,
              cudaStream_t stream) {
    CUtensorMap ta, tb;
    if (int rc = make_tmap(&ta, A, p.M, p.K, lda, GEMM_BLOCK_M)) return rc;
    if (int rc = make_tmap(&tb, W, p.N, p.K, ldw, BLOCK_N)) return rc;
    switch (epi) {
        case EPI_BIAS_BF16: return launch_one<BLOCK_N, EPI_BIAS_BF16>(ta, tb, p, stream);
        case EPI_BIAS_GELU_BF16: return launch_one<BLOCK_N, EPI_BIAS_GELU_BF16>(ta, tb, p, stream);
        case EPI_RESID_F32: return launch_one<BLOCK_N, EPI_RESID_F32>(ta, tb, p, stream);
        case EPI_PATCH_F32: return launch_one<BLOCK_N, EPI_PATCH_F32>(ta, tb, p, stream);
        case EPI_BIAS_F32: return launch_one<BLOCK_N, EPI_BIAS_F32>(ta, tb, p, stream);
    }
    return fail("unknown GEMM epilogue " + std::to_string(epi));
}

}  // namespace

int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const GemmParams& p, int epi,
                cudaStream_t stream, int prof_tag) {
    if (p.M <= 0) return 0;
    ProfScope prof(prof_tag, stream);
    if (p.K % GEMM_BLOCK_K != 0 || p.K <= 0) return fail("GEMM K must be a positive multiple of 64");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (lda % 8) || (ldw % 8))
        return fail("GEMM operands must be 16-byte aligned with row pitch a multiple of 8 elements");
    if (p.N % 256 == 0) return launch_bn<256>(A, lda, W, ldw, p, epi, stream);
    if (p.N % 192 == 0) return launch_bn<192>(A, lda, W, ldw, p, epi, stream);
    if (p.N % 128 == 0) return launch_bn<128>(A, lda, W, ldw, p, epi, stream);
    return fail("GEMM N must be a multiple of 128 (got " + std::to_string(p.N) + ")");
}

}  // namespace cbas
