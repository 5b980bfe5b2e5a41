// Error reporting, launch accounting and device queries shared by the C ABI.
#include "../../include/cbas_b200.h"
#include "common.h"

#include <atomic>

namespace cbas {

namespace {
thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
}  // namespace

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(const std::string& msg) {
    set_error(msg);
    return 1;
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return 1;
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace cbas

extern "C" {
const char* cbas_b200_last_error(void) { return cbas::g_last_error.c_str(); }
int cbas_b200_abi_version(void) { return CBAS_B200_ABI_VERSION; }
unsigned long long cbas_b200_launch_count(void) { return cbas::g_launches.load(); }
}
