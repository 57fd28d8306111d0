// Library-level entry points: version, error reporting, device check.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
static const char* const kTuneNames[kTuneCount] = {"SPMM_PIPE", "RES_TC", "RES_ENT", "SPMM_CSM", "SPMM_RTILE"};
// SPMM_PIPE: persistent pipelined SpMM, 4 blocks per SM
static const int kTuneDefaults[kTuneCount] = {4, 0, 0, 106, 2};
static std::atomic<int> g_tune[kTuneCount] = {{-1}, {-1}, {-1}, {-1}, {-1}};
static_assert(kTuneCount == 5, "update the tuning tables");

int tuning_value(int key) {
    int v = g_tune[key].load(std::memory_order_relaxed);
    if (v >= 0) return v;
    char name[64];
    snprintf(name, sizeof(name), "TGCN_%s", kTuneNames[key]);
    const char* e = getenv(name);
    v = e ? atoi(e) : kTuneDefaults[key];
    if (v < 0) v = kTuneDefaults[key];
    g_tune[key].store(v, std::memory_order_relaxed);
    return v;
}

unsigned long long peer_timeout_ns() {
    static const unsigned long long v = [] {
        const char* e = getenv("TGCN_PEER_TIMEOUT_S");
        double sec = e ? atof(e) : 120.0;
        if (!(sec > 0.0)) sec = 120.0;
        return (unsigned long long)(sec * 1e9);
    }();
    return v;
}
}  // namespace tgcn

// Select a kernel variant at run time (tests, sweeps): key in {"SPMM_PIPE", "RES_TC", "RES_ENT", "SPMM_CSM", "SPMM_RTILE"};
// returns 0, or -1 for an unknown key.  RES_TC = 1: contraction of the resident forward kernel on tcgen05 (3xTF32).  SPMM_PIPE = blocks per
// SM of the persistent pipelined SpMM kernel (0: one-thread-per-float4 kernel); SPMM_CSM = blocks per SM (4, 5, 6, 8) the block-staged-CSR
// SpMM kernel is compiled for (0: off; 100 + b, the default 106: only for slabs of 96 MB or more); SPMM_RTILE = use registered row-tile plans.
extern "C" int tgcn_set_tuning(const char* key, int value) {
    for (int i = 0; i < tgcn::kTuneCount; ++i)
        if (key && strcmp(key, tgcn::kTuneNames[i]) == 0) {
            tgcn::g_tune[i].store(value < 0 ? tgcn::kTuneDefaults[i] : value);
            return 0;
        }
    return -1;
}

extern "C" long long tgcn_launch_count(void) { return tgcn::g_launches.load(); }

extern "C" int tgcn_version(void) { return 200; }  // ABI version; tgcn_b200/_lib.py ABI_VERSION must match

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
