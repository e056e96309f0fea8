// Library-level entry points of include/mage_b200.h: error string, device probe, version.
#include "common.cuh"

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

extern "C" const char* mage_last_error(void) { return mage::g_err; }

extern "C" int mage_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { mage::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); cudaGetLastError(); return MAGE_ERR_CUDA; }
    return n;
}

extern "C" const char* mage_version(void) { return "mageslam_b200 0.1 (sm_100a)"; }
