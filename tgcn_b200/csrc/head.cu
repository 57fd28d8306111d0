// Fused classifier head of the reference models:
//     x[Q,I] -> fc1 (I -> Hd) -> BatchNorm1d (training: batch statistics) -> ReLU -> fc2 (Hd -> C) -> log_softmax
// (pytorch_hcp_tgcn.py:143-155: `x.view(..)`, fc1, dense1_bn, relu, fc2, log_softmax).  The reference runs it as
// ~10 ATen launches forward and ~12 backward; at batch 64 every one of them is launch-latency bound.  Here:
//   head_fwd1: one CTA per block of FB hidden features: the fc1 dot products for ALL samples, then the batch
//              statistics of those features (the whole batch is in the CTA), normalise, ReLU, running-stat update;
//   head_fwd2: fc2 + log_softmax, one warp per sample;
//   head_bwd1: log_softmax / fc2 / ReLU / BatchNorm backward in one CTA (Q x Hd elements), producing dh and the
//              small gradients (dW2, db2, dgamma, dbeta, db1);
//   head_bwd2: dW1 = dh^T x and dx = dh W1 by column blocks with dh resident in shared memory.
// fp32 throughout, fixed summation orders (deterministic).
#include "common.cuh"
#include "tc_common.cuh"

namespace tgcn {

constexpr int kHeadFB = 2;        // hidden features per CTA in head_fwd1 / head_bwd1
constexpr int kHeadThreads = 256;
constexpr int kHeadFwd1Threads = 512;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid = ceil(Hd / FB); block 256 = 8 warps; warp w handles samples q = w, w+8, ... for the CTA's FB features
__global__ void __launch_bounds__(kHeadFwd1Threads)
head_fwd1_kernel(const float* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ b1,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
                 float momentum, float eps, int training,
                 float* __restrict__ act, float* __restrict__ xhat, float* __restrict__ invstd_out,
                 int Q, int I, int Hd, float drop_p, uint32_t drop_seed, const uint32_t* drop_step) {
    extern __shared__ float hs[];                  // [Q][FB] pre-activations
    const DropCfg drop = drop_resolve(drop_p, drop_seed, drop_step);
    __shared__ float s_mean[kHeadFB], s_inv[kHeadFB];
    const int f0 = blockIdx.x * kHeadFB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const bool vec = (I & 3) == 0 && aligned16(x) && aligned16(W1);
    for (int q0 = warp * 2; q0 < Q; q0 += nw * 2) {          // two samples per pass: 2 x FB accumulators
        float acc[2][kHeadFB];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int f = 0; f < kHeadFB; ++f) acc[u][f] = 0.f;
        const bool two = q0 + 1 < Q;
        if (vec) {
            const float4* xa = reinterpret_cast<const float4*>(x + (int64_t)q0 * I);
            const float4* xb = reinterpret_cast<const float4*>(x + (int64_t)(two ? q0 + 1 : q0) * I);
#pragma unroll 4
            for (int i = lane; i < I / 4; i += 32) {
                const float4 a = __ldg(xa + i), b = __ldg(xb + i);
#pragma unroll
                for (int f = 0; f < kHeadFB; ++f) {
                    if (f0 + f < Hd) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(W1 + (int64_t)(f0 + f) * I) + i);
                        acc[0][f] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[0][f]))));
                        acc[1][f] = fmaf(b.x, w.x, fmaf(b.y, w.y, fmaf(b.z, w.z, fmaf(b.w, w.w, acc[1][f]))));
                    }
                }
            }
        } else {
            for (int i = lane; i < I; i += 32) {
                const float a = __ldg(x + (int64_t)q0 * I + i), b = __ldg(x + (int64_t)(two ? q0 + 1 : q0) * I + i);
#pragma unroll
                for (int f = 0; f < kHeadFB; ++f) {
                    if (f0 + f < Hd) {
                        const float w = __ldg(W1 + (int64_t)(f0 + f) * I + i);
                        acc[0][f] = fmaf(a, w, acc[0][f]);
                        acc[1][f] = fmaf(b, w, acc[1][f]);
                    }
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kHeadFB; ++f) {
            const float s0 = warp_sum(acc[0][f]), s1 = warp_sum(acc[1][f]);
            if (lane == 0) {
                const float bb = (f0 + f < Hd && b1) ? __ldg(b1 + f0 + f) : 0.f;
                hs[q0 * kHeadFB + f] = s0 + bb;
                if (two) hs[(q0 + 1) * kHeadFB + f] = s1 + bb;
            }
        }
    }
    __syncthreads();
    // batch statistics per feature (biased variance for the normalisation, unbiased for the running estimate)
    if (warp < kHeadFB && f0 + warp < Hd) {
        const int f = warp;
        float mean, inv;
        if (training) {
            float s = 0.f;
            for (int q = lane; q < Q; q += 32) s += hs[q * kHeadFB + f];
            mean = warp_sum(s) / (float)Q;
            float v = 0.f;
            for (int q = lane; q < Q; q += 32) { const float d = hs[q * kHeadFB + f] - mean; v = fmaf(d, d, v); }
            const float var = warp_sum(v) / (float)Q;
            inv = 1.0f / sqrtf(var + eps);
            if (lane == 0 && running_mean) {
                const float unb = Q > 1 ? var * (float)Q / (float)(Q - 1) : var;
                running_mean[f0 + f] = (1.f - momentum) * running_mean[f0 + f] + momentum * mean;
                running_var[f0 + f] = (1.f - momentum) * running_var[f0 + f] + momentum * unb;
            }
        } else {
            mean = running_mean[f0 + f];
            inv = 1.0f / sqrtf(running_var[f0 + f] + eps);
        }
        if (lane == 0) { s_mean[f] = mean; s_inv[f] = inv; if (invstd_out) invstd_out[f0 + f] = inv; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Q * kHeadFB; i += blockDim.x) {
        const int q = i / kHeadFB, f = i - q * kHeadFB;
        if (f0 + f >= Hd) continue;
        const float xh = (hs[i] - s_mean[f]) * s_inv[f];
        const float y = fmaf(xh, gamma ? __ldg(gamma + f0 + f) : 1.f, beta ? __ldg(beta + f0 + f) : 0.f);
        if (xhat) xhat[(int64_t)q * Hd + f0 + f] = xh;
        float a = fmaxf(y, 0.f);
        if (drop.scale != 0.f) a = drop_apply(a, (uint64_t)((int64_t)q * Hd + f0 + f), drop);   // drop2, pytorch_hcp_tgcn.py:150
        act[(int64_t)q * Hd + f0 + f] = a;
    }
}

// one warp per sample: logits = act W2^T + b2, log_softmax.  The sample's activation row is read once into
// registers (all loads independent), then dotted with the C rows of W2.
constexpr int kHeadActRegs = 8;           // 32 lanes x 8 = 256 hidden units per pass
__global__ void __launch_bounds__(kHeadThreads)
head_fwd2_kernel(const float* __restrict__ act, const float* __restrict__ W2, const float* __restrict__ b2,
                 float* __restrict__ logp, int Q, int Hd, int C) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= Q) return;
    float mine = 0.f;                                // lane c (< 32) keeps logit c
    float part[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) part[c] = 0.f;
    for (int f0 = 0; f0 < Hd; f0 += 32 * kHeadActRegs) {
        float a[kHeadActRegs];
#pragma unroll
        for (int u = 0; u < kHeadActRegs; ++u) {
            const int f = f0 + u * 32 + lane;
            a[u] = f < Hd ? __ldg(act + (int64_t)q * Hd + f) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (c >= C) break;
            float s = 0.f;
#pragma unroll
            for (int u = 0; u < kHeadActRegs; ++u) {
                const int f = f0 + u * 32 + lane;
                if (f < Hd) s = fmaf(a[u], __ldg(W2 + (int64_t)c * Hd + f), s);
            }
            part[c] += s;
        }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        if (c >= C) break;
        const float s = warp_sum(part[c]) + (b2 ? __ldg(b2 + c) : 0.f);
        if (lane == c) mine = s;
    }
    float m = lane < C ? mine : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = lane < C ? expf(mine - m) : 0.f;
    const float lse = logf(warp_sum(e)) + m;
    if (lane < C) logp[(int64_t)q * C + lane] = mine - lse;
}

// one CTA per block of FB hidden features (warp f <-> feature f0+f): log_softmax backward (recomputed per CTA, it
// is Q x C), the CTA's columns of dW2 and of da = dlogits W2, ReLU mask, BatchNorm backward, dh[:, features]
__global__ void __launch_bounds__(kHeadThreads)
head_bwd1_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp, const float* __restrict__ act,
                 const float* __restrict__ xhat, const float* __restrict__ invstd, const float* __restrict__ gamma,
                 const float* __restrict__ W2, float* __restrict__ dW2, float* __restrict__ db2,
                 float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ db1, float* __restrict__ dh,
                 int Q, int Hd, int C, float drop_scale) {
    extern __shared__ float sm[];                  // dlogits [Q][C]
    float* dlog = sm;
    const int tid = threadIdx.x, T = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int f = blockIdx.x * kHeadFB + warp;
    for (int q = tid; q < Q; q += T) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += __ldg(dlogp + (int64_t)q * C + c);
        for (int c = 0; c < C; ++c)
            dlog[q * C + c] = __ldg(dlogp + (int64_t)q * C + c) - expf(__ldg(logp + (int64_t)q * C + c)) * s;
    }
    __syncthreads();
    if (blockIdx.x == 0 && db2)
        for (int c = tid; c < C; c += T) {
            float s = 0.f;
            for (int q = 0; q < Q; ++q) s += dlog[q * C + c];
            db2[c] = s;
        }
    if (warp >= kHeadFB || f >= Hd) return;
    // this warp's feature: lanes run over the samples
    for (int c = 0; c < C; ++c) {                  // dW2[c][f] = sum_q dlogits[q][c] act[q][f]
        float s = 0.f;
        for (int q = lane; q < Q; q += 32) s = fmaf(dlog[q * C + c], __ldg(act + (int64_t)q * Hd + f), s);
        s = warp_sum(s);
        if (lane == 0) dW2[(int64_t)c * Hd + f] = s;
    }
    float sg = 0.f, sb = 0.f;
    for (int q = lane; q < Q; q += 32) {           // dy = relu'(.) * (dlogits W2)[q][f]
        float d = 0.f;
        for (int c = 0; c < C; ++c) d = fmaf(dlog[q * C + c], __ldg(W2 + (int64_t)c * Hd + f), d);
        d = __ldg(act + (int64_t)q * Hd + f) > 0.f ? d * drop_scale : 0.f;   // act > 0: passed the ReLU and kept by the dropout
        sg = fmaf(d, __ldg(xhat + (int64_t)q * Hd + f), sg);
        sb += d;
    }
    sg = warp_sum(sg); sb = warp_sum(sb);
    if (lane == 0) { if (dgamma) dgamma[f] = sg; if (dbeta) dbeta[f] = sb; }
    const float k = (gamma ? __ldg(gamma + f) : 1.f) * __ldg(invstd + f) / (float)Q;
    float s1 = 0.f;
    for (int q = lane; q < Q; q += 32) {
        float d = 0.f;
        for (int c = 0; c < C; ++c) d = fmaf(dlog[q * C + c], __ldg(W2 + (int64_t)c * Hd + f), d);
        d = __ldg(act + (int64_t)q * Hd + f) > 0.f ? d * drop_scale : 0.f;
        const float v = k * ((float)Q * d - sb - __ldg(xhat + (int64_t)q * Hd + f) * sg);
        dh[(int64_t)q * Hd + f] = v;
        s1 += v;
    }
    s1 = warp_sum(s1);
    if (lane == 0 && db1) db1[f] = s1;
}

// column blocks of CB input features: dW1[:, cols] = dh^T x[:, cols];  dx[:, cols] = dh W1[:, cols].
// dh, the x columns and the W1 columns are staged in shared memory; register tiles of 4 hidden features (dW1) and
// 4 samples x 4 hidden features per step (dx) keep the shared-memory traffic at ~0.5 loads per FMA.
constexpr int kHeadCB = 16;
__global__ void __launch_bounds__(kHeadThreads)
head_bwd2_kernel(const float* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ W1,
                 float* __restrict__ dW1, float* __restrict__ dx, int Q, int I, int Hd, int HdP) {
    extern __shared__ __align__(16) float sm2[];   // dh [Q][HdP] | xs [Q][CB] | ws [HdP][CB]
    float* dhs = sm2;
    float* xs = dhs + Q * HdP;
    float* ws = xs + Q * kHeadCB;
    const int i0 = blockIdx.x * kHeadCB, tid = threadIdx.x, T = blockDim.x;
    const int cb = min(kHeadCB, I - i0);
    // staging: when the shapes allow it every row is one 1-D bulk copy (dh as a whole, 128-byte row segments of x
    // and W1), all completing on one mbarrier -- one round trip to memory instead of a loop of dependent loads
    __shared__ __align__(8) uint64_t bar;
    const bool bulk = HdP == Hd && cb == kHeadCB && (I & 3) == 0 && aligned16(dh) && aligned16(x) && aligned16(W1) &&
                      (size_t)Q * Hd * 4 < (1u << 20);
    if (bulk) {
        if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)Q * Hd * 4u + (uint32_t)Q * (kHeadCB * 4u) + (dx ? (uint32_t)Hd * (kHeadCB * 4u) : 0u);
            tc::mbar_arrive_expect_tx(&bar, bytes);
            tc::bulk_g2s(dhs, dh, (uint32_t)Q * Hd * 4u, &bar);
        }
        __syncthreads();
        const int nrows = Q + (dx ? Hd : 0);
        for (int r = tid; r < nrows; r += T) {
            if (r < Q) tc::bulk_g2s(xs + r * kHeadCB, x + (int64_t)r * I + i0, kHeadCB * 4u, &bar);
            else tc::bulk_g2s(ws + (r - Q) * kHeadCB, W1 + (int64_t)(r - Q) * I + i0, kHeadCB * 4u, &bar);
        }
        tc::mbar_wait(&bar, 0);
    } else {
        for (int i = tid; i < Q * HdP; i += T) {
            const int q = i / HdP, f = i - q * HdP;
            dhs[i] = f < Hd ? __ldg(dh + (int64_t)q * Hd + f) : 0.f;
        }
        for (int i = tid; i < Q * kHeadCB; i += T) {
            const int q = i / kHeadCB, c = i - q * kHeadCB;
            xs[i] = c < cb ? __ldg(x + (int64_t)q * I + i0 + c) : 0.f;
        }
        if (dx)
            for (int i = tid; i < HdP * kHeadCB; i += T) {
                const int f = i / kHeadCB, c = i - f * kHeadCB;
                ws[i] = (c < cb && f < Hd) ? __ldg(W1 + (int64_t)f * I + i0 + c) : 0.f;
            }
        __syncthreads();
    }
    // dW1: tiles of 4 hidden features x 1 column; lanes run along the columns (coalesced stores)
    for (int t = tid; t < (HdP / 4) * kHeadCB; t += T) {
        const int fg = t / kHeadCB, c = t - fg * kHeadCB;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < Q; ++q) {
            const float4 d = *reinterpret_cast<const float4*>(dhs + q * HdP + 4 * fg);
            const float xv = xs[q * kHeadCB + c];
            acc.x = fmaf(d.x, xv, acc.x); acc.y = fmaf(d.y, xv, acc.y);
            acc.z = fmaf(d.z, xv, acc.z); acc.w = fmaf(d.w, xv, acc.w);
        }
        if (c < cb) {
            const float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * fg + u < Hd) dW1[(int64_t)(4 * fg + u) * I + i0 + c] = o[u];
        }
    }
    if (!dx) return;
    // dx: tiles of 4 samples x 1 column, hidden features consumed four at a time
    for (int t = tid; t < ((Q + 3) / 4) * kHeadCB; t += T) {
        const int qg = t / kHeadCB, c = t - qg * kHeadCB;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int f = 0; f < HdP; f += 4) {
            const float w0 = ws[(f + 0) * kHeadCB + c], w1 = ws[(f + 1) * kHeadCB + c];
            const float w2 = ws[(f + 2) * kHeadCB + c], w3 = ws[(f + 3) * kHeadCB + c];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = min(4 * qg + u, Q - 1);
                const float4 d = *reinterpret_cast<const float4*>(dhs + q * HdP + f);
                acc[u] = fmaf(d.x, w0, fmaf(d.y, w1, fmaf(d.z, w2, fmaf(d.w, w3, acc[u]))));
            }
        }
        if (c < cb) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (4 * qg + u < Q) dx[(int64_t)(4 * qg + u) * I + i0 + c] = acc[u];
        }
    }
}

// bighead.cu
bool bighead_applies(int Q, int I, int Hd);
int bighead_workspace_chunks(int I);
int bighead_fwd(const float* x, const float* W1, const float* b1, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float momentum, float eps, int training, float drop_p, uint32_t drop_seed,
                const uint32_t* drop_step, float* act, float* xhat, float* invstd, float* partial, int Q, int I, int Hd,
                cudaStream_t st);
int bighead_bwd(const float* dh, const float* x, float* W1, float* dW1, float* dx, const tgcn_fc1_update_t* upd, int Q, int I,
                int Hd, cudaStream_t st);

}  // namespace tgcn

using namespace tgcn;

extern "C" int64_t tgcn_head_workspace(int Q, int I, int Hd) {
    if (Q < 1 || I < 1 || Hd < 1 || !bighead_applies(Q, I, Hd)) return 0;
    return (int64_t)bighead_workspace_chunks(I) * Q * Hd * (int64_t)sizeof(float);
}

extern "C" int tgcn_head_fused_update_supported(int Q, int I, int Hd) { return bighead_applies(Q, I, Hd) ? 1 : 0; }

extern "C" int tgcn_head_fwd(const float* x, const float* W1, const float* b1, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, float momentum, float eps, int training,
                             const float* W2, const float* b2, const tgcn_dropout_t* drop, float* act, float* xhat,
                             float* invstd, float* logp, void* workspace, int Q, int I, int Hd, int C, void* stream) {
    const float drop_p = (drop && drop->p > 0.f && training) ? drop->p : 0.f;
    TGCN_REQUIRE(drop_p < 1.f, "tgcn_head_fwd: dropout p = %g must be < 1", (double)drop_p);
    TGCN_REQUIRE(Q >= 1 && I >= 1 && Hd >= 1 && C >= 1, "tgcn_head_fwd: bad sizes");
    TGCN_REQUIRE(x && W1 && W2 && act && logp, "tgcn_head_fwd: null pointer");
    TGCN_SUPPORTED(C <= 32, "tgcn_head_fwd: at most 32 classes (one warp per sample), got %d", C);
    TGCN_REQUIRE(training || (running_mean && running_var), "tgcn_head_fwd: evaluation mode needs the running statistics");
    TGCN_REQUIRE(!training || Q > 1, "tgcn_head_fwd: batch statistics need more than one sample");
    const size_t smem = sizeof(float) * (size_t)(Q + 1) * kHeadFB;
    TGCN_SUPPORTED(smem <= 48 * 1024, "tgcn_head_fwd: batch %d too large for the fused head", Q);
    cudaStream_t st = as_stream(stream);
    if (bighead_applies(Q, I, Hd)) {
        // large fc1: the weight is streamed once by a dedicated kernel (bighead.cu), the rest of the head is unchanged
        TGCN_REQUIRE(workspace && aligned16(workspace), "tgcn_head_fwd: this shape needs tgcn_head_workspace(Q, I, Hd) bytes of workspace");
        TGCN_PROPAGATE(bighead_fwd(x, W1, b1, gamma, beta, running_mean, running_var, momentum, eps, training, drop_p,
                                   drop ? drop->seed : 0u, drop ? drop->step : nullptr, act, xhat, invstd,
                                   reinterpret_cast<float*>(workspace), Q, I, Hd, st));
        head_fwd2_kernel<<<(unsigned)ceil_div(Q, 2), 64, 0, st>>>(act, W2, b2, logp, Q, Hd, C);
        TGCN_LAUNCH_CHECK("head_fwd2");
        return TGCN_OK;
    }
    head_fwd1_kernel<<<(unsigned)ceil_div(Hd, kHeadFB), kHeadFwd1Threads, smem, st>>>(x, W1, b1, gamma, beta, running_mean, running_var,
                                                                                 momentum, eps, training, act, xhat, invstd, Q, I, Hd,
                                                                                 drop_p, drop ? drop->seed : 0u, drop ? drop->step : nullptr);
    TGCN_LAUNCH_CHECK("head_fwd1");
    head_fwd2_kernel<<<(unsigned)ceil_div(Q, 2), 64, 0, st>>>(act, W2, b2, logp, Q, Hd, C);
    TGCN_LAUNCH_CHECK("head_fwd2");
    return TGCN_OK;
}

extern "C" int tgcn_head_bwd(const float* dlogp, const float* logp, const float* act, const float* xhat, const float* invstd,
                             const float* x, float* W1, const float* gamma, const float* W2,
                             float* dx, float* dW1, float* db1, float* dgamma, float* dbeta, float* dW2, float* db2,
                             float* dh_scratch, const tgcn_dropout_t* drop, const tgcn_fc1_update_t* upd, int Q, int I, int Hd,
                             int C, void* stream) {
    TGCN_REQUIRE(Q >= 1 && I >= 1 && Hd >= 1 && C >= 1, "tgcn_head_bwd: bad sizes");
    const float drop_p = (drop && drop->p > 0.f) ? drop->p : 0.f;
    TGCN_REQUIRE(drop_p < 1.f, "tgcn_head_bwd: dropout p = %g must be < 1", (double)drop_p);
    TGCN_REQUIRE(dlogp && logp && act && xhat && invstd && x && W1 && W2 && (dW1 || upd) && dW2 && dh_scratch, "tgcn_head_bwd: null pointer");
    const bool big = bighead_applies(Q, I, Hd);
    TGCN_SUPPORTED(!upd || big, "tgcn_head_bwd: the fused fc1 update covers large heads only (tgcn_head_fused_update_supported)");
    const size_t smem1 = sizeof(float) * (size_t)Q * C;
    const int HdP = (Hd + 3) & ~3;
    const size_t smem2 = sizeof(float) * ((size_t)Q * HdP + (size_t)Q * kHeadCB + (size_t)HdP * kHeadCB);
    TGCN_SUPPORTED(smem1 <= 200 * 1024 && (big || smem2 <= 200 * 1024), "tgcn_head_bwd: Q=%d Hd=%d too large for the fused head", Q, Hd);
    cudaStream_t st = as_stream(stream);
    if (smem1 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(head_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_head_bwd: %s", cudaGetErrorString(e));
    }
    if (!big && smem2 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(head_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_head_bwd: %s", cudaGetErrorString(e));
    }
    head_bwd1_kernel<<<(unsigned)ceil_div(Hd, kHeadFB), kHeadFB * 32, smem1, st>>>(dlogp, logp, act, xhat, invstd, gamma, W2, dW2, db2, dgamma, dbeta, db1, dh_scratch, Q, Hd, C, 1.0f / (1.0f - drop_p));
    TGCN_LAUNCH_CHECK("head_bwd1");
    if (big) return bighead_bwd(dh_scratch, x, W1, dW1, dx, upd, Q, I, Hd, st);
    head_bwd2_kernel<<<(unsigned)ceil_div(I, kHeadCB), kHeadThreads, smem2, st>>>(dh_scratch, x, W1, dW1, dx, Q, I, Hd, HdP);
    TGCN_LAUNCH_CHECK("head_bwd2");
    return TGCN_OK;
}
