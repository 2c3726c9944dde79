// Shared host-side plumbing of libsir_b200: error reporting, launch counting, small device helpers.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/sir_b200.h"

namespace sir {

extern thread_local char g_error[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Optional per-stage device timing (sir_profile_enable / sir_profile_read): a pair of CUDA events on the
// launching stream around each named stage.  Off by default; bench.py turns it on for a separate pass.
extern bool g_profile;
void prof_mark(const char* name, cudaStream_t st, bool begin);
struct ProfScope {
    const char* name;
    cudaStream_t st;
    ProfScope(const char* n, cudaStream_t s) : name(n), st(s) {
        if (g_profile) prof_mark(name, st, true);
    }
    ~ProfScope() {
        if (g_profile) prof_mark(name, st, false);
    }
};

#define SIR_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::sir::fail(SIR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                                \
    } while (0)

#define SIR_CHECK_LAUNCH(name)                                                                           \
    do {                                                                                                 \
        cudaError_t _e = cudaGetLastError();                                                             \
        if (_e != cudaSuccess)                                                                           \
            return ::sir::fail(SIR_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e));   \
        ::sir::count_launch();                                                                           \
    } while (0)

// cudaFuncSetAttribute and the SM count are per DEVICE, a process may hold handles on several: one-time setup is
// remembered per device ordinal, not per process.
struct DeviceOnce {
    std::atomic<uint64_t> done{0};
    bool need(int& dev) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        return ((done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0;
    }
    void mark(int dev) { done.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

#define SIR_SMEM_OPTIN(kern, bytes)                                                                      \
    do {                                                                                                 \
        static ::sir::DeviceOnce _once;                                                                  \
        int _dev = 0;                                                                                    \
        if (_once.need(_dev)) {                                                                          \
            SIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _once.mark(_dev);                                                                            \
        }                                                                                                \
    } while (0)

// SM count of the current device (cached per device ordinal).
inline int device_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

// Grow-only device buffer owned by a handle.
struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t want) {
        if (want <= bytes) return SIR_OK;
        if (ptr) {
            SIR_CUDA(cudaDeviceSynchronize());
            SIR_CUDA(cudaFree(ptr));
            ptr = nullptr;
            bytes = 0;
        }
        SIR_CUDA(cudaMalloc(&ptr, want));
        bytes = want;
        return SIR_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

}  // namespace sir
