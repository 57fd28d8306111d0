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
enum TuneKey { kTuneSpmmTile = 0, kTuneSpmmPipe = 1, kTuneSpmmStaged = 2, kTuneResTc = 3, kTuneResEnt = 4, kTuneSpmmWarpRow = 5, kTuneSpmmCsm = 6, kTuneSpmmRtile = 7, kTuneCount = 8 };
int tuning_value(int key);

}  // namespace tgcn

// acc += w * x on a float4 as TWO packed FFMA2 instructions (Blackwell fma.rn.f32x2: two IEEE fp32 fused
// multiply-adds per issue slot, bit-identical to four scalar fmaf) -- the gather kernels are issue-bound, not
// bandwidth-bound, so halving the FMA instruction count is what buys throughput.
#ifdef __CUDACC__
__device__ __forceinline__ void fma4_packed(float4& acc, float w, const float4& x) {
    unsigned long long a01, a23, x01, x23, ww;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "f"(acc.z), "f"(acc.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x01) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(x23) : "f"(x.z), "f"(x.w));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a01) : "l"(ww), "l"(x01));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a23) : "l"(ww), "l"(x23));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a23));
}
#endif

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
