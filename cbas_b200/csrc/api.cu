// Error reporting, launch accounting and device queries shared by the C ABI.
#include "../../include/cbas_b200.h"
#include "common.h"

#include <atomic>
#include <mutex>
#include <vector>

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

namespace {
struct ProfRec { int tag; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::atomic<bool> g_prof_on{false};
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

ProfScope::ProfScope(int tag, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{tag, prof_event(), prof_event()};
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, stream);
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("CBAS_B200_PDL");
        return e && e[0] == '1';
    }();
    return on;
}

int sm_count() {
    static std::atomic<int> cache[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace cbas

extern "C" {
const char* cbas_b200_last_error(void) { return cbas::g_last_error.c_str(); }
int cbas_b200_abi_version(void) { return CBAS_B200_ABI_VERSION; }
unsigned long long cbas_b200_launch_count(void) { return cbas::g_launches.load(); }

int cbas_b200_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(cbas::g_prof_mu);
    for (auto& r : cbas::g_prof) { cbas::g_prof_pool.push_back(r.a); cbas::g_prof_pool.push_back(r.b); }
    cbas::g_prof.clear();
    cbas::g_prof_on.store(on != 0);
    return 0;
}

int cbas_b200_profile_read(int tag, double* total_ms, long long* launches) {
    if (!total_ms || !launches) return cbas::fail("null argument");
    if (tag < 0 || tag >= cbas::PROF_NUM_TAGS) return cbas::fail("profile tag out of range");
    std::lock_guard<std::mutex> lk(cbas::g_prof_mu);
    double ms = 0.0; long long n = 0;
    for (auto& r : cbas::g_prof) {
        if (r.tag != tag) continue;
        if (cudaEventSynchronize(r.b) != cudaSuccess) return cbas::fail("profile event synchronize failed");
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return cbas::fail("profile event elapsed failed");
        ms += t; ++n;
    }
    *total_ms = ms; *launches = n;
    return 0;
}
}
