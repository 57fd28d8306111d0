// Library-level entry points: version, error reporting, device check.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include "common.cuh"

namespace tgcn {
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tgcn

extern "C" long long tgcn_launch_count(void) { return tgcn::g_launches.load(); }

extern "C" int tgcn_version(void) { return 100; }  // 0.1.0

extern "C" const char* tgcn_last_error(void) { return tgcn::g_err; }

extern "C" int tgcn_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return major == 10 ? 1 : 0;
}
