// Shared helpers for the sm_100a kernels and their C-ABI wrappers.
#pragma once
#include <algorithm>
#include <mutex>
#include <vector>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mage_b200.h"

namespace mage {

void set_error(const char* fmt, ...);

#define MAGE_CUDA_TRY(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            ::mage::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            return MAGE_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

#define MAGE_REQUIRE(cond, code, ...)                 \
    do {                                              \
        if (!(cond)) {                                \
            ::mage::set_error(__VA_ARGS__);           \
            return (code);                            \
        }                                             \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// The stream-ordered memory pool the library allocates from on the current device (set up on first use, once per device, thread-safe).
// Default: the device's default pool with its release threshold raised to the maximum, so blocks freed with cudaFreeAsync stay
// cached across synchronisations (at the default threshold of 0 they go back to the driver at the next sync and every new window
// pays a fresh allocation). That is a process-wide setting the host application shares; MAGE_POOL_PRIVATE=1 gives the library a
// pool of its own instead -- measured slower and less steady on B200 (a fresh local-BA window 2.0 - 4.4 ms instead of 1.8 ms), which
// is why it is the option and not the default. nullptr when the device / driver has no stream-ordered allocator.
inline cudaMemPool_t library_pool()
{
    constexpr int kMaxDev = 64;
    static cudaMemPool_t pools[kMaxDev] = {};
    static std::once_flag once[kMaxDev];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    std::call_once(once[dev], [dev]() {
        if (!getenv("MAGE_POOL_PRIVATE")) {
            cudaMemPool_t dp = nullptr;
            if (cudaDeviceGetDefaultMemPool(&dp, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(dp, cudaMemPoolAttrReleaseThreshold, &keep);
                pools[dev] = dp;
            } else cudaGetLastError();
            return;
        }
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pools[dev] = pool;
        } else cudaGetLastError();
    });
    return pools[dev];
}

// Stream-ordered allocation from the library's pool (the default pool, untouched, when the private one could not be created).
inline cudaError_t pool_malloc_async(void** ptr, size_t bytes, cudaStream_t s)
{
    cudaMemPool_t pool = library_pool();
    return pool ? cudaMallocFromPoolAsync(ptr, bytes, pool, s) : cudaMallocAsync(ptr, bytes, s);
}

// One device allocation carved into aligned sub-buffers. pooled = true takes it from the library's stream-ordered memory pool
// (cudaMallocFromPoolAsync, release threshold raised so freed blocks stay cached): a few microseconds instead of the ~1-4 ms cudaFree +
// cudaMalloc of a multi-megabyte block cost when a new problem is set up per call (a BundlerLib instance per local-BA window).
// A pooled arena must only be released when the work using it has been synchronised (cudaFreeAsync does not wait like cudaFree).
struct DeviceArena {
    uint8_t* base = nullptr;
    size_t size = 0, used = 0;
    bool pooled = false;
    size_t reserve(size_t bytes, size_t align = 256) { used = align_up(used, align); size_t o = used; used += bytes; return o; }
    cudaError_t commit()
    {
        size = used;
        if (!pooled) return cudaMalloc(&base, size ? size : 256);
        cudaMemPool_t pool = library_pool();
        cudaError_t e = pool ? cudaMallocFromPoolAsync(reinterpret_cast<void**>(&base), size ? size : 256, pool, static_cast<cudaStream_t>(0)) : cudaErrorNotSupported;
        if (e == cudaSuccess) return cudaStreamSynchronize(static_cast<cudaStream_t>(0));    // usable from any stream afterwards
        if (getenv("MAGE_DEBUG_POOL")) fprintf(stderr, "[pool] stream-ordered allocation of %zu bytes failed (%s, pool %p): plain cudaMalloc\n", size, cudaGetErrorString(e), (void*)pool);
        cudaGetLastError();                                    // no stream-ordered pool on this device / driver: plain allocation
        pooled = false; base = nullptr;
        return cudaMalloc(&base, size ? size : 256);
    }
    template <class T> T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
    void release()
    {
        if (base) { if (pooled) cudaFreeAsync(base, static_cast<cudaStream_t>(0)); else cudaFree(base); }
        base = nullptr;
    }
};

// Pinned host staging buffers, pooled process-wide (cudaHostAlloc costs more than most of the calls that need one): the structure upload
// of a bundle-adjustment problem takes one (the batched call's worker threads one each), the one-call host paths of the matchers pack their
// inputs into one and land their results in it -- a copy between the device and a caller's pageable array is a staged, synchronous
// transfer of the driver (about 8 us each), one DMA from / to pinned memory plus a memcpy is not.
struct PinnedStage { uint8_t* p = nullptr; size_t cap = 0; };
struct PinnedPool { std::mutex mu; std::vector<PinnedStage> free; };
inline PinnedPool& pinned_pool() { static PinnedPool pool; return pool; }
inline PinnedStage stage_acquire(size_t bytes)
{
    PinnedPool& P = pinned_pool();
    PinnedStage st;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        size_t best = P.free.size();
        for (size_t i = 0; i < P.free.size(); i++)                // best fit: a small request must not take the one large buffer
            if (P.free[i].cap >= bytes && (best == P.free.size() || P.free[i].cap < P.free[best].cap)) best = i;
        if (best < P.free.size()) { st = P.free[best]; P.free.erase(P.free.begin() + best); return st; }
    }
    // nothing large enough: a new buffer (buffers that are too small stay pooled -- cudaFreeHost synchronises the device, which a
    // latency path must never do on somebody else's behalf)
    const size_t cap = align_up(std::max<size_t>(bytes + bytes / 4, 1 << 20), 1 << 20);
    if (cudaHostAlloc(reinterpret_cast<void**>(&st.p), cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); st.p = nullptr; return st; }
    st.cap = cap;
    return st;
}
inline void stage_release(PinnedStage st)
{
    if (!st.p) return;
    PinnedPool& P = pinned_pool();
    PinnedStage drop;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        P.free.push_back(st);
        if (P.free.size() > 48) {                                 // far more than any call pattern keeps in flight: let the smallest one go
            size_t k = 0;
            for (size_t i = 1; i < P.free.size(); i++) if (P.free[i].cap < P.free[k].cap) k = i;
            drop = P.free[k]; P.free.erase(P.free.begin() + k);
        }
    }
    if (drop.p) cudaFreeHost(drop.p);
}

// Device scratch of the one-call host paths, pooled like the pinned staging (a per-frame object such as the spatial index would otherwise
// allocate its scratch anew every frame: a stream-ordered allocation + a synchronisation, 20 us of a 100 us call). A buffer is handed
// back only after the borrowing call has synchronised its stream, so the next borrower may use it from any stream.
struct DevStage { uint8_t* p = nullptr; size_t cap = 0; int dev = -1; };
struct DevPool { std::mutex mu; std::vector<DevStage> free; };
inline DevPool& dev_pool() { static DevPool pool; return pool; }
inline DevStage dev_stage_acquire(size_t bytes)
{
    DevPool& P = dev_pool();
    DevStage st;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(P.mu);
        size_t best = P.free.size();
        for (size_t i = 0; i < P.free.size(); i++)
            if (P.free[i].dev == dev && P.free[i].cap >= bytes && (best == P.free.size() || P.free[i].cap < P.free[best].cap)) best = i;
        if (best < P.free.size()) { st = P.free[best]; P.free.erase(P.free.begin() + best); return st; }
    }
    const size_t cap = align_up(std::max<size_t>(bytes + bytes / 4, 1 << 20), 1 << 20);
    if (cudaMalloc(reinterpret_cast<void**>(&st.p), cap) != cudaSuccess) { cudaGetLastError(); st.p = nullptr; return st; }
    st.cap = cap; st.dev = dev;
    return st;
}
inline void dev_stage_release(DevStage st)
{
    if (!st.p) return;
    DevPool& P = dev_pool();
    DevStage drop;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        P.free.push_back(st);
        if (P.free.size() > 32) {
            size_t k = 0;
            for (size_t i = 1; i < P.free.size(); i++) if (P.free[i].cap < P.free[k].cap) k = i;
            drop = P.free[k]; P.free.erase(P.free.begin() + k);
        }
    }
    if (drop.p) cudaFree(drop.p);
}

// the stream of a host-path call that brings none of its own: one per calling thread (and device), created at first use
inline cudaStream_t thread_stream()
{
    static thread_local cudaStream_t s = nullptr;
    static thread_local int sdev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!s || sdev != dev) { s = nullptr; if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); s = nullptr; } sdev = dev; }
    return s;
}

// Optional per-kernel timing (CUDA events on the launching stream), used by bench.py for the live roofline figure.
enum ProfSlot { PROF_RESIZE = 0, PROF_BLUR, PROF_FAST, PROF_SELECT, PROF_ORIENT_DESCRIBE, PROF_MATCH_DIR, PROF_MATCH_EMIT, PROF_BA_STEP, PROF_SLOTS };
bool prof_enabled();
void prof_begin(int slot, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);
struct ProfScope {
    int slot; cudaStream_t s; bool on;
    ProfScope(int slot_, cudaStream_t s_) : slot(slot_), s(s_), on(prof_enabled()) { if (on) prof_begin(slot, s); }
    ~ProfScope() { if (on) prof_end(slot, s); }
};

__device__ __forceinline__ int warp_reduce_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

} // namespace mage
