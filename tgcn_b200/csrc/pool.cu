// K3: permuted max-pool over coarsened graphs (siblings are contiguous after the binary-tree
// permutation), with the argmax needed for the gradient, optionally fused with the ReLU that
// precedes it in every reference model.  Pure streaming: one thread per float4 of the output.
#include "common.cuh"

namespace tgcn {

// torch.max(dim) CPU/CUDA rule: first maximal element wins; a NaN beats any number (first NaN wins).
__device__ __forceinline__ void take_max(float cand, int s, float& best, int& arg) {
    const bool better = (cand > best) || (cand != cand && best == best);
    if (better) {
        best = cand;
        arg = s;
    }
}

template <int P, bool RELU, int VEC>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx, int64_t rows_out,
                int G, float drop_p, uint32_t drop_seed, const uint32_t* drop_step) {
    const DropCfg drop = drop_resolve(RELU ? drop_p : 0.f, drop_seed, drop_step);
    // rows_out = Q * N/P pooled rows; a pooled row reads P consecutive input rows of G floats
    const int GV = G / VEC;
    const int64_t total = rows_out * GV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / GV;
        const int gv = (int)(i - r * GV);
        float best[VEC];
        int arg[VEC];
#pragma unroll
        for (int s = 0; s < P; ++s) {
            float v[VEC];
            const float* src = x + (r * P + s) * G + gv * VEC;
            if (VEC == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(src));
                v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            } else {
                v[0] = __ldg(src);
            }
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                float u = v[c];
                if (RELU) u = (u != u) ? u : fmaxf(u, 0.f);   // F.relu keeps NaN
                if (RELU && drop.scale != 0.f) u = drop_apply(u, (uint64_t)((r * P + s) * G + gv * VEC + c), drop);
                if (s == 0) { best[c] = u; arg[c] = 0; }
                else take_max(u, s, best[c], arg[c]);
            }
        }
        float* dst = y + r * G + gv * VEC;
        uint8_t* di = idx + r * G + gv * VEC;
        if (VEC == 4) {
            *reinterpret_cast<float4*>(dst) = make_float4(best[0], best[1 % VEC], best[2 % VEC], best[3 % VEC]);
            *reinterpret_cast<uchar4*>(di) = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1 % VEC],
                                                         (unsigned char)arg[2 % VEC], (unsigned char)arg[3 % VEC]);
        } else {
            dst[0] = best[0];
            di[0] = (uint8_t)arg[0];
        }
    }
}

template <int P, bool RELU>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, const float* __restrict__ x,
                const float* __restrict__ y, float drop_scale, float* __restrict__ dx, int64_t rows_out, int G) {
    const int64_t total = rows_out * G;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / G;
        const int g = (int)(i - r * G);
        const float gval = __ldg(dy + i);
        const int a = idx[i];
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const int64_t o = (r * P + s) * G + g;
            float v = (s == a) ? gval : 0.f;
            if (RELU && s == a && y != nullptr) {
                // gate on the pooled OUTPUT: y > 0 <=> its source passed the ReLU and was kept by the dropout
                const float yv = __ldg(y + i);
                v = (yv > 0.f || yv != yv) ? gval * drop_scale : 0.f;
            } else if (RELU && s == a) {
                const float xv = __ldg(x + o);
                // relu'(x) = 1 for x > 0, 0 for x <= 0; NaN input propagates NaN like autograd's threshold_backward
                v = (xv > 0.f) ? gval : ((xv != xv) ? gval : 0.f);
            }
            dx[o] = v;
        }
    }
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_pool_max_fwd(const float* x, float* y, uint8_t* idx, int Q, int N, int G, int p, int relu,
                                 const tgcn_dropout_t* drop, void* stream) {
    const float dp = (drop && drop->p > 0.f) ? drop->p : 0.f;
    TGCN_REQUIRE(dp < 1.f, "tgcn_pool_max_fwd: dropout p = %g must be < 1", (double)dp);
    TGCN_REQUIRE(dp == 0.f || relu, "tgcn_pool_max_fwd: dropout is fused behind the ReLU only");
    const uint32_t dseed = drop ? drop->seed : 0u;
    const uint32_t* dstep = drop ? drop->step : nullptr;
    TGCN_REQUIRE(Q >= 0 && N >= 0 && G >= 1, "tgcn_pool_max_fwd: bad sizes");
    TGCN_SUPPORTED(p == 2 || p == 4, "tgcn_pool_max_fwd: pool size %d (reference has gcn_pool p=2, gcn_pool_4 p=4)", p);
    TGCN_REQUIRE(N % p == 0, "tgcn_pool_max_fwd: vertex count %d not divisible by pool size %d", N, p);
    const int64_t rows_out = (int64_t)Q * (N / p);
    if (rows_out == 0) return TGCN_OK;
    TGCN_REQUIRE(x && y && idx, "tgcn_pool_max_fwd: null pointer");
    cudaStream_t st = as_stream(stream);
    const bool vec = (G % 4 == 0) && aligned16(x) && aligned16(y) && ((reinterpret_cast<uintptr_t>(idx) & 3u) == 0);
    const int64_t total = rows_out * (vec ? G / 4 : G);
    const unsigned blocks = (unsigned)min64(ceil_div(total, 256), (int64_t)kNumSMs * 32);
#define TGCN_POOL_FWD(P, R)                                                                            \
    do {                                                                                               \
        if (vec) pool_fwd_kernel<P, R, 4><<<blocks, 256, 0, st>>>(x, y, idx, rows_out, G, dp, dseed, dstep);             \
        else pool_fwd_kernel<P, R, 1><<<blocks, 256, 0, st>>>(x, y, idx, rows_out, G, dp, dseed, dstep);                 \
    } while (0)
    if (p == 2) { if (relu) TGCN_POOL_FWD(2, true); else TGCN_POOL_FWD(2, false); }
    else        { if (relu) TGCN_POOL_FWD(4, true); else TGCN_POOL_FWD(4, false); }
#undef TGCN_POOL_FWD
    TGCN_LAUNCH_CHECK("pool_max_fwd");
    return TGCN_OK;
}

extern "C" int tgcn_pool_max_bwd(const float* dy, const uint8_t* idx, const float* x, const float* y, float* dx, int Q,
                                 int N, int G, int p, int relu, const tgcn_dropout_t* drop, void* stream) {
    const float dp = (drop && drop->p > 0.f) ? drop->p : 0.f;
    TGCN_REQUIRE(dp < 1.f, "tgcn_pool_max_bwd: dropout p = %g must be < 1", (double)dp);
    TGCN_REQUIRE(dp == 0.f || (relu && y), "tgcn_pool_max_bwd: the dropout backward gates on the pooled output y");
    const float dscale = 1.0f / (1.0f - dp);
    TGCN_REQUIRE(Q >= 0 && N >= 0 && G >= 1, "tgcn_pool_max_bwd: bad sizes");
    TGCN_SUPPORTED(p == 2 || p == 4, "tgcn_pool_max_bwd: pool size %d", p);
    TGCN_REQUIRE(N % p == 0, "tgcn_pool_max_bwd: vertex count %d not divisible by pool size %d", N, p);
    const int64_t rows_out = (int64_t)Q * (N / p);
    if (rows_out == 0) return TGCN_OK;
    TGCN_REQUIRE(dy && idx && dx, "tgcn_pool_max_bwd: null pointer");
    TGCN_REQUIRE(!relu || x || y, "tgcn_pool_max_bwd: relu backward needs the pool input x or the pooled output y");
    cudaStream_t st = as_stream(stream);
    const unsigned blocks = (unsigned)min64(ceil_div(rows_out * G, 256), (int64_t)kNumSMs * 32);
    if (p == 2) {
        if (relu) pool_bwd_kernel<2, true><<<blocks, 256, 0, st>>>(dy, idx, x, y, dscale, dx, rows_out, G);
        else pool_bwd_kernel<2, false><<<blocks, 256, 0, st>>>(dy, idx, x, y, dscale, dx, rows_out, G);
    } else {
        if (relu) pool_bwd_kernel<4, true><<<blocks, 256, 0, st>>>(dy, idx, x, y, dscale, dx, rows_out, G);
        else pool_bwd_kernel<4, false><<<blocks, 256, 0, st>>>(dy, idx, x, y, dscale, dx, rows_out, G);
    }
    TGCN_LAUNCH_CHECK("pool_max_bwd");
    return TGCN_OK;
}
