// Library-level entry points of include/mage_b200.h: error string, device probe, version.
#include "common.cuh"

#include <mutex>
#include <vector>

namespace mage {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
} // namespace mage

namespace mage {
namespace {
struct ProfState {
    bool on = false;
    std::mutex mu;
    struct Pending { int slot; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> open_begin = std::vector<cudaEvent_t>(PROF_SLOTS, nullptr);
    std::vector<cudaEvent_t> pool;
    double total_ms[PROF_SLOTS] = {0};
    long long launches[PROF_SLOTS] = {0};
    cudaEvent_t get() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
};
ProfState g_prof;
const char* kProfNames[PROF_SLOTS] = {"k_resize", "k_blur", "k_fast", "k_select", "k_orient_describe", "k_match_dir", "k_match_emit", "k_ba_step"};
} // namespace
bool prof_enabled() { return g_prof.on; }
void prof_begin(int slot, cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_prof.mu);
    cudaEvent_t e = g_prof.get();
    cudaEventRecord(e, s);
    g_prof.open_begin[slot] = e;
}
void prof_end(int slot, cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_prof.mu);
    cudaEvent_t e = g_prof.get();
    cudaEventRecord(e, s);
    g_prof.pending.push_back({slot, g_prof.open_begin[slot], e});
}
} // namespace mage

extern "C" int mage_profile_enable(int on)
{
    mage::g_prof.on = on != 0;
    return MAGE_OK;
}
// Synchronises the device and folds all recorded (begin, end) event pairs into per-kernel totals.
extern "C" int mage_profile_collect(void)
{
    using namespace mage;
    MAGE_CUDA_TRY(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (auto& p : g_prof.pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { g_prof.total_ms[p.slot] += ms; g_prof.launches[p.slot]++; }
        g_prof.pool.push_back(p.a); g_prof.pool.push_back(p.b);
    }
    g_prof.pending.clear();
    return MAGE_OK;
}
extern "C" int mage_profile_reset(void)
{
    int rc = mage_profile_collect();
    for (int i = 0; i < mage::PROF_SLOTS; i++) { mage::g_prof.total_ms[i] = 0; mage::g_prof.launches[i] = 0; }
    return rc;
}
extern "C" int mage_profile_slots(void) { return mage::PROF_SLOTS; }
extern "C" const char* mage_profile_name(int slot) { return (slot >= 0 && slot < mage::PROF_SLOTS) ? mage::kProfNames[slot] : ""; }
extern "C" int mage_profile_get(int slot, double* total_ms, long long* groups)
{
    MAGE_REQUIRE(slot >= 0 && slot < mage::PROF_SLOTS && total_ms && groups, MAGE_ERR_INVALID, "mage_profile_get: bad argument");
    *total_ms = mage::g_prof.total_ms[slot]; *groups = mage::g_prof.launches[slot];
    return MAGE_OK;
}

extern "C" const char* mage_last_error(void) { return mage::g_err; }

extern "C" int mage_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { mage::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); cudaGetLastError(); return MAGE_ERR_CUDA; }
    return n;
}

extern "C" const char* mage_version(void) { return "mageslam_b200 0.1 (sm_100a)"; }

// convenience for host-language bindings that hold raw device pointers (tests, bench): synchronous device -> host copy
extern "C" int mage_memcpy_d2h(void* dst, const void* src, size_t bytes)
{
    MAGE_REQUIRE(dst && src, MAGE_ERR_INVALID, "mage_memcpy_d2h: null argument");
    MAGE_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return MAGE_OK;
}
