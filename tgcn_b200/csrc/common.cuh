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
enum TuneKey { kTuneSpmmPipe = 0, kTuneResTc = 1, kTuneResEnt = 2, kTuneSpmmCsm = 3, kTuneSpmmRtile = 4, kTuneCount = 5 };
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

// ---- fused dropout (nn.Dropout between ReLU and the pool / fc2 in the reference models, pytorch_hcp_tgcn.py:106,125,136,150)
// Counter-based: the keep decision of activation element `elem` is a hash of (seed, step counter, elem), so the
// forward needs no mask tensor and a CUDA-graph replay draws a new mask whenever the caller's step counter moved.
// The backward never regenerates the mask: a positive pooled / post-ReLU output implies its source element was kept,
// so the gradient is the routed gradient times `scale` = 1/(1-p).
#ifdef __CUDACC__
namespace tgcn {
struct DropCfg { float scale; uint32_t thresh; uint32_t key; };      // scale == 0: dropout off
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// torch semantics: out = x * mask / (1 - p); a dropped NaN stays NaN (NaN * 0)
__device__ __forceinline__ float drop_apply(float v, uint64_t elem, const DropCfg& d) {
    const uint32_t h = mix32((uint32_t)elem ^ mix32((uint32_t)(elem >> 32) ^ d.key));
    return v * (h >= d.thresh ? d.scale : 0.f);
}
// resolved on the device at kernel start: key from the seed and the step counter the caller advances every step
__device__ __forceinline__ DropCfg drop_resolve(float p, uint32_t seed, const uint32_t* step) {
    DropCfg d;
    if (!(p > 0.f)) { d.scale = 0.f; d.thresh = 0u; d.key = 0u; return d; }
    d.scale = 1.0f / (1.0f - p);
    d.thresh = (uint32_t)fminf(p * 4294967296.0f, 4294967040.0f);
    d.key = mix32(seed + 0x9E3779B9u * (step ? *step : 0u));
    return d;
}
}  // namespace tgcn
#endif

// ---- cross-rank flags over NVLink peer memory (peer.cu, bighead.cu, halo reads) --------------------------------
#ifdef __CUDACC__
namespace tgcn {
constexpr int kPeerMaxWorld = 8;
// Bound of every cross-rank wait, in nanoseconds of %globaltimer (TGCN_PEER_TIMEOUT_S, default 120 s: ordinary rank
// skew -- a validation pass, a checkpoint, a first-step lazy load on one rank -- must not kill the job; a lost
// peer still cannot hang the device for ever: the waiter traps, which surfaces as a CUDA error on the host).
unsigned long long peer_timeout_ns();
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *slot >= want (acquire, system scope); traps after limit_ns
__device__ __forceinline__ void peer_wait_flag(const unsigned int* slot, unsigned int want, unsigned long long limit_ns) {
    if (ld_acquire_sys(slot) >= want) return;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(slot) < want) {
        __nanosleep(64);
        if (global_ns() - t0 > limit_ns) __trap();
    }
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
}  // namespace tgcn
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
