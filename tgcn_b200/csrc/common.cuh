// Shared host/device helpers for the tgcn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tgcn_b200.h"

namespace tgcn {

// Records a message for tgcn_last_error() (thread-local) and returns `code`.
int set_error(int code, const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Number of kernel launches this library has issued (host-side counter, for bench.py's
// `gpu_launches`; launches recorded during graph capture count once per capture).
void count_launch();

// Run-time tuning knobs (tgcn_set_tuning / environment TGCN_<NAME> read once); -1 = built-in default.
enum TuneKey { kTuneSpmmTile = 0, kTuneSpmmPipe = 1, kTuneSpmmStaged = 2, kTuneResTc = 3, kTuneResEnt = 4, kTuneCount = 5 };
int tuning_value(int key);

}  // namespace tgcn

#define TGCN_REQUIRE(cond, ...)                                            \
    do {                                                                   \
        if (!(cond)) return tgcn::set_error(TGCN_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define TGCN_SUPPORTED(cond, ...)                                              \
    do {                                                                       \
        if (!(cond)) return tgcn::set_error(TGCN_ERR_UNSUPPORTED, __VA_ARGS__); \
    } while (0)

#define TGCN_LAUNCH_CHECK(name)                                                                  \
    do {                                                                                         \
        tgcn::count_launch();                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess)                                                                  \
            return tgcn::set_error(TGCN_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__));      \
    } while (0)

#define TGCN_PROPAGATE(expr)          \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != TGCN_OK) return rc__; \
    } while (0)
