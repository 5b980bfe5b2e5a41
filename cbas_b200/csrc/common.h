// Host-side helpers shared by the translation units of libcbas_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace cbas {

// thread-local error message behind cbas_b200_last_error()
void set_error(const std::string& msg);
int fail(const std::string& msg);               // set_error + return 1
int check_cuda(cudaError_t e, const char* what);  // 0 if cudaSuccess, else records and returns 1
void count_launch(int n = 1);
int sm_count();

struct GemmParams;
// C[M,N] = A[M,K] W[N,K]^T with one of the GemmEpilogue modes; lda/ldw in elements.
int launch_gemm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, const GemmParams& p, int epi,
                cudaStream_t stream);

#define CBAS_CHECK(expr)                                   \
    do {                                                   \
        if (int _rc = ::cbas::check_cuda((expr), #expr)) return _rc; \
    } while (0)

}  // namespace cbas
