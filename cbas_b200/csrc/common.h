// Host-side helpers shared by the translation units of libcbas_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

namespace cbas {

// thread-local error message behind cbas_b200_last_error()
void set_error(const std::string& msg);
int fail(const std::string& msg);               // set_error + return 1
int check_cuda(cudaError_t e, const char* what);  // 0 if cudaSuccess, else records and returns 1
void count_launch(int n = 1);
int sm_count();  // of the calling thread's current device
// Programmatic dependent launch for the kernels that call pdl_wait() (GEMM, attention, LayerNorm): off unless
// CBAS_B200_PDL=1 (parity-green, but no measurable gain on the power-capped step: profiles/pdl_ab_r02.txt).  Fills `attr` and returns 1 when the launch should carry the attribute, else 0.
bool pdl_enabled();
inline int pdl_attr(cudaLaunchAttribute* attr) {
    if (!pdl_enabled()) return 0;
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = 1;
    return 1;
}

// Per-device "configured" state for cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which is a per-DEVICE
// attribute: one entry per device ordinal.  The value is published only after the attribute call has succeeded, so
// a second host thread either repeats the (idempotent) call or sees it done - it never launches ahead of it.
struct DeviceSmemOptIn {
    std::atomic<long long> cur[64] = {};
    template <typename Kernel>
    cudaError_t ensure(Kernel kern, long long bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::atomic<long long>& c = cur[dev & 63];
        if (bytes <= c.load(std::memory_order_acquire)) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        long long seen = c.load(std::memory_order_relaxed);
        while (seen < bytes && !c.compare_exchange_weak(seen, bytes, std::memory_order_release)) {}
        return cudaSuccess;
    }
};

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: a handle created
// on cuda:1 works from a host thread whose current device is cuda:0 (workthreads.py: one EncodeThread /
// ClassificationThread pair per device, all in one process).
struct DeviceGuard {
    int prev = -1;
    bool good = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { good = false; prev = -1; return; }
        if (prev != device) {
            if (cudaSetDevice(device) != cudaSuccess) { good = false; prev = -1; }
        } else {
            prev = -1;  // nothing to restore
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    bool ok() const { return good; }
};

// Optional per-kernel timing (cbas_b200_profile_*): CUDA events recorded on the launch stream around a launch.
enum ProfTag : int {
    PROF_PREPROCESS = 0, PROF_PATCH_GEMM, PROF_LAYERNORM, PROF_QKV_GEMM, PROF_ATTENTION, PROF_PROJ_GEMM, PROF_UP_GEMM,
    PROF_DOWN_GEMM, PROF_FINAL_LN, PROF_HEAD_SPLIT, PROF_HEAD_PROJ_GEMM, PROF_HEAD_FEATURES, PROF_HEAD_LIN0_GEMM,
    PROF_HEAD_CENTER, PROF_HEAD_IH_GEMM, PROF_HEAD_LSTM, PROF_ACTOGRAM, PROF_OTHER, PROF_NUM_TAGS
};
struct ProfScope {
    ProfScope(int tag, cudaStream_t s);
    ~ProfScope();
    int slot;
    cudaStream_t stream;
};

struct GemmParams;
// C[M,N] = A[M,K] W[N,K]^T with one of the GemmEpilogue modes; lda/ldw in elements.
int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const GemmParams& p, int epi,
                cudaStream_t stream, int prof_tag = PROF_OTHER);

// Row-major [rows, cols] tensor map, box box_cols x box_rows with a 128-byte inner extent, SWIZZLE_128B.
int make_tmap_2d(CUtensorMap* map, const void* base, bool f32, int rows, int cols, int ld, int box_cols, int box_rows);
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, int d0, int d1, int d2, int ld, int box0, int box1);
int make_tmap_3d_f32(CUtensorMap* map, const void* base, long long d0, long long d1, long long d2, long long pitch1,
                     long long pitch2, int box0, int box1, int box2);
void set_gemm_cta_group(int cg);  // 0 auto, 1 single-CTA tiles, 2 CTA-pair tiles

#define CBAS_CHECK(expr)                                   \
    do {                                                   \
        if (int _rc = ::cbas::check_cuda((expr), #expr)) return _rc; \
    } while (0)

}  // namespace cbas
