// Sample-resident fused layer kernels (TGCN_ENGINE_RESIDENT).
//
// The columns of the [N, Q*D] slab are independent through the whole recursion, so when one
// sample's slab [N, D] fits in shared memory the complete layer runs out of one CTA per sample:
//
//   forward : P_0 = x_q;  P_j = L~ P_{j-1}  (CSR gather, shared -> shared, ping-pong);
//             out += P_j * W'_j  after every step (accumulators live in registers for all K steps);
//             epilogue adds the bias and optionally applies ReLU + the permuted max-pool, so the
//             un-pooled activation never reaches HBM.  Every P_j is written to HBM exactly once
//             (coalesced, fire-and-forget) for the backward.
//   backward: dW'_j = P_j^T dOut  (P_j streamed back from HBM/L2, dOut resident in shared memory),
//             dx = sum_j (L~^T)^j (dOut W'_j^T) by Horner's rule on a resident [N, D] accumulator.
//             Per-sample partials are reduced over the batch by a second, deterministic kernel
//             that also applies the transposed weight mix and produces the bias gradient.
//
// The reference computes the same quantities with K-1 dense `bmm` launches plus einsum/permute
// copies per layer (tgcn/nn/gcn.py:108-154, :189-237) and autograd's stored intermediates.
// Work is fp32 FFMA + shared-memory gathers: at these sizes (Q*N*K*D*G <= ~0.3 GFLOP) the
// contraction is far below one SM-microsecond of tensor-core work and staging hi/lo TF32 operand
// images would cost more shared-memory traffic than the FFMAs it replaces; the tcgen05 engine
// (contract_tc*.cu) serves the large-graph configs where the stack streams from HBM.
#include "common.cuh"

namespace tgcn {

constexpr int kResThreads = 512;
constexpr size_t kResSmemLimit = 227 * 1024;
constexpr int kResMaxK = 32;

__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }

// W'_j[e] of the reference recursion (see mix_weights_kernel in contract.cu), fp64 accumulate.
__device__ __forceinline__ float mixed_weight(const float* __restrict__ W, int K, int64_t inner, int j, int64_t e,
                                              int recursion) {
    if (recursion == TGCN_RECURSION_CHEBYSHEV) return __ldg(W + (int64_t)j * inner + e);
    const double c = j < 2 ? 1.0 : 2.0;
    double s = 0.0, sign = 1.0;
    for (int k = j; k < K; k += 2, sign = -sign) s += sign * c * (double)__ldg(W + (int64_t)k * inner + e);
    return (float)s;
}

__device__ __forceinline__ void fma4s(float4& acc, float w, const float4& x) {
    acc.x = fmaf(w, x.x, acc.x);
    acc.y = fmaf(w, x.y, acc.y);
    acc.z = fmaf(w, x.z, acc.z);
    acc.w = fmaf(w, x.w, acc.w);
}

// out[v] += sum_e val[e] * in[col[e]][v] for one row, one float4 column group; CSR pairs packed as
// int2 (col, float bits) in shared memory (kCsrSmem) or read from global memory.
template <bool kCsrSmem>
__device__ __forceinline__ float4 gather_row(const int2* __restrict__ csr_s, const int* __restrict__ col,
                                             const float* __restrict__ val, int e, const int e1,
                                             const float4* __restrict__ in4, const int V, const int v) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (; e + 4 <= e1; e += 4) {
        int c0, c1, c2, c3;
        float w0, w1, w2, w3;
        if (kCsrSmem) {
            const int2 p0 = csr_s[e], p1 = csr_s[e + 1], p2 = csr_s[e + 2], p3 = csr_s[e + 3];
            c0 = p0.x; c1 = p1.x; c2 = p2.x; c3 = p3.x;
            w0 = __int_as_float(p0.y); w1 = __int_as_float(p1.y); w2 = __int_as_float(p2.y); w3 = __int_as_float(p3.y);
        } else {
            c0 = __ldg(col + e); c1 = __ldg(col + e + 1); c2 = __ldg(col + e + 2); c3 = __ldg(col + e + 3);
            w0 = __ldg(val + e); w1 = __ldg(val + e + 1); w2 = __ldg(val + e + 2); w3 = __ldg(val + e + 3);
        }
        const float4 x0 = in4[c0 * V + v], x1 = in4[c1 * V + v], x2 = in4[c2 * V + v], x3 = in4[c3 * V + v];
        fma4s(a0, w0, x0);
        fma4s(a1, w1, x1);
        fma4s(a0, w2, x2);
        fma4s(a1, w3, x3);
    }
    for (; e < e1; ++e) {
        int c0;
        float w0;
        if (kCsrSmem) { const int2 p0 = csr_s[e]; c0 = p0.x; w0 = __int_as_float(p0.y); }
        else { c0 = __ldg(col + e); w0 = __ldg(val + e); }
        fma4s(a0, w0, in4[c0 * V + v]);
    }
    return make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
}

// torch.max(dim) rule (pool.cu): first maximal element wins, NaN beats any number.
__device__ __forceinline__ void take_max_r(float cand, int s, float& best, int& arg) {
    const bool better = (cand > best) || (cand != cand && best == best);
    if (better) { best = cand; arg = s; }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct ResFwdParams {
    const int* rowptr; const int* col; const float* val;
    const float* x;        // [Q,N,D]
    const float* W;        // [K,D,G] raw layer weights
    const float* bias;     // [N,G] / [G] / null
    float* out;            // [Q,N,G] or null
    float* y;              // [Q,N/p,G] or null (fused pool)
    uint8_t* idx;          // [Q,N/p,G]
    float* stack;          // [Q][K][NP][DP] or null (inference)
    int N, NP, nnz, D, DP, V, G, GP, GG, K;
    int bias_mode, recursion, pool_p, relu;
};

struct ResSmemFwd { size_t rowptr, csr, P, W, total; };

static ResSmemFwd res_fwd_smem(int N, int nnz, int DP, int GP, bool csr_smem) {
    ResSmemFwd s{};
    const int NP = round_up4(N);
    size_t o = 0;
    s.P = o;      o += sizeof(float) * 2 * (size_t)NP * DP;
    s.W = o;      o += sizeof(float) * 2 * (size_t)DP * GP;
    s.csr = o;    o += csr_smem ? sizeof(int2) * (size_t)nnz : 0;
    s.rowptr = o; o += sizeof(int) * (size_t)(N + 1);
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

template <int TPT, bool kCsrSmem>
__global__ void __launch_bounds__(kResThreads, 1)
resident_fwd_kernel(const ResFwdParams p, const ResSmemFwd lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* Pbuf = reinterpret_cast<float*>(smem + lay.P);
    float* Wsm = reinterpret_cast<float*>(smem + lay.W);
    int2* csr_s = reinterpret_cast<int2*>(smem + lay.csr);
    int* rowptr_s = reinterpret_cast<int*>(smem + lay.rowptr);

    const int tid = threadIdx.x, T = blockDim.x;
    const int q = blockIdx.x;
    const int N = p.N, NP = p.NP, D = p.D, DP = p.DP, V = p.V, G = p.G, GP = p.GP, GG = p.GG, K = p.K;
    const int slab = NP * DP;                               // floats per P buffer / per stack slab

    for (int i = tid; i <= N; i += T) rowptr_s[i] = __ldg(p.rowptr + i);
    if (kCsrSmem)
        for (int e = tid; e < p.nnz; e += T) csr_s[e] = make_int2(__ldg(p.col + e), __float_as_int(__ldg(p.val + e)));
    {   // P_0 = x_q (zero pad columns / rows), also the first slab of the saved stack
        const float* xq = p.x + (int64_t)q * N * D;
        float* st0 = p.stack ? p.stack + (int64_t)q * K * slab : nullptr;
        for (int i = tid; i < slab; i += T) {
            const int n = i / DP, d = i - n * DP;
            const float v = (n < N && d < D) ? __ldg(xq + (int64_t)n * D + d) : 0.f;
            Pbuf[i] = v;
            if (st0) st0[i] = v;
        }
    }
    const int64_t inner = (int64_t)D * G;
    auto stage_w = [&](int j) {
        float* dst = Wsm + (j & 1) * DP * GP;
        for (int i = tid; i < DP * GP; i += T) {
            const int d = i / GP, g = i - d * GP;
            dst[i] = (d < D && g < G) ? mixed_weight(p.W, K, inner, j, (int64_t)d * G + g, p.recursion) : 0.f;
        }
    };
    stage_w(0);
    __syncthreads();

    const int ntiles = (NP / 4) * GG;
    float4 acc[TPT][4];
#pragma unroll
    for (int s = 0; s < TPT; ++s)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[s][r] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int j = 0; j < K; ++j) {
        float* Pcur = Pbuf + (j & 1) * slab;
        if (j > 0) {
            const float4* in4 = reinterpret_cast<const float4*>(Pbuf + ((j - 1) & 1) * slab);
            float4* out4 = reinterpret_cast<float4*>(Pcur);
            float4* st4 = p.stack ? reinterpret_cast<float4*>(p.stack + ((int64_t)q * K + j) * slab) : nullptr;
            const bool cheb = (p.recursion == TGCN_RECURSION_CHEBYSHEV) && j >= 2;
            for (int i = tid; i < N * V; i += T) {
                const int n = i / V, v = i - n * V;
                float4 r = gather_row<kCsrSmem>(csr_s, p.col, p.val, rowptr_s[n], rowptr_s[n + 1], in4, V, v);
                if (cheb) {   // T_j = 2 L~ T_{j-1} - T_{j-2}; T_{j-2} is what the output buffer still holds
                    const float4 o = out4[i];
                    r.x = fmaf(-1.f, o.x, 2.f * r.x); r.y = fmaf(-1.f, o.y, 2.f * r.y);
                    r.z = fmaf(-1.f, o.z, 2.f * r.z); r.w = fmaf(-1.f, o.w, 2.f * r.w);
                }
                out4[i] = r;
                if (st4) st4[i] = r;
            }
            stage_w(j);
            __syncthreads();
        }
        // contraction step: acc[tile] += P_j[4 rows][DP] * W'_j[DP][4 g]
        const float4* P4 = reinterpret_cast<const float4*>(Pcur);
        const float4* W4 = reinterpret_cast<const float4*>(Wsm + (j & 1) * DP * GP);
#pragma unroll
        for (int s = 0; s < TPT; ++s) {
            const int t = tid + s * T;
            if (t >= ntiles) break;
            const int rg = t / GG, gg = t - rg * GG;
            for (int v = 0; v < V; ++v) {
                const float4 w0 = W4[(4 * v + 0) * GG + gg], w1 = W4[(4 * v + 1) * GG + gg];
                const float4 w2 = W4[(4 * v + 2) * GG + gg], w3 = W4[(4 * v + 3) * GG + gg];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 a = P4[(rg * 4 + r) * V + v];
                    fma4s(acc[s][r], a.x, w0);
                    fma4s(acc[s][r], a.y, w1);
                    fma4s(acc[s][r], a.z, w2);
                    fma4s(acc[s][r], a.w, w3);
                }
            }
        }
    }

    // epilogue: bias, optional ReLU + max-pool over the tile's sibling rows
    const int pp = p.pool_p;
#pragma unroll
    for (int s = 0; s < TPT; ++s) {
        const int t = tid + s * T;
        if (t >= ntiles) break;
        const int rg = t / GG, gg = t - rg * GG;
        const int g0 = gg * 4;
        float o[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = rg * 4 + r;
            float b[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < N) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (g0 + c < G) {
                        if (p.bias_mode == TGCN_BIAS_PER_VERTEX) b[c] = __ldg(p.bias + (int64_t)n * G + g0 + c);
                        else if (p.bias_mode == TGCN_BIAS_PER_FILTER) b[c] = __ldg(p.bias + g0 + c);
                    }
                }
            }
            o[r][0] = acc[s][r].x + b[0]; o[r][1] = acc[s][r].y + b[1];
            o[r][2] = acc[s][r].z + b[2]; o[r][3] = acc[s][r].w + b[3];
        }
        if (p.out) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int n = rg * 4 + r;
                if (n >= N) continue;
                float* dst = p.out + ((int64_t)q * N + n) * G + g0;
                if ((G & 3) == 0) *reinterpret_cast<float4*>(dst) = make_float4(o[r][0], o[r][1], o[r][2], o[r][3]);
                else
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (g0 + c < G) dst[c] = o[r][c];
            }
        }
        if (p.y) {
            const int groups = 4 / pp;                      // pooled rows produced by this tile (1 or 2)
            for (int u = 0; u < groups; ++u) {
                const int m = (rg * 4) / pp + u;            // pooled row
                if (m * pp >= N) continue;
                float best[4];
                int arg[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    for (int s2 = 0; s2 < pp; ++s2) {
                        float vv = o[u * pp + s2][c];
                        if (p.relu) vv = (vv != vv) ? vv : fmaxf(vv, 0.f);
                        if (s2 == 0) { best[c] = vv; arg[c] = 0; }
                        else take_max_r(vv, s2, best[c], arg[c]);
                    }
                }
                const int64_t off = ((int64_t)q * (N / pp) + m) * G + g0;
                if ((G & 3) == 0) {
                    *reinterpret_cast<float4*>(p.y + off) = make_float4(best[0], best[1], best[2], best[3]);
                    *reinterpret_cast<uchar4*>(p.idx + off) =
                        make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (g0 + c < G) { p.y[off + c] = best[c]; p.idx[off + c] = (uint8_t)arg[c]; }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct ResBwdParams {
    const int* rowptrT; const int* colT; const float* valT;   // CSR of L~^T (only read when dx != null)
    const float* dout;     // [Q,N,G] or null
    const float* dy;       // [Q,N/p,G] (fused pool) or null
    const uint8_t* idx;    // [Q,N/p,G]
    const float* y;        // pooled forward output (ReLU mask) or null
    const float* stack;    // [Q][K][NP][DP]
    const float* W;        // [K,D,G] raw layer weights
    float* dWpart;         // [Q][K][DP][GP] per-sample gradient in the power basis
    float* dbpart;         // [Q][GP] per-sample column sums of dOut (per-filter bias) or null
    float* dx;             // [Q,N,D] or null
    int N, NP, nnz, D, DP, V, G, GP, GG, K;
    int recursion, pool_p, relu;
};

struct ResSmemBwd { size_t dOut, red, A, W, csr, rowptr, total; };

static ResSmemBwd res_bwd_smem(int N, int nnz, int DP, int GP, bool need_dx, bool csr_smem) {
    ResSmemBwd s{};
    const int NP = round_up4(N);
    size_t o = 0;
    s.dOut = o;   o += sizeof(float) * (size_t)NP * GP;
    s.red = o;    o += sizeof(float) * (size_t)kResThreads;
    s.A = o;      o += need_dx ? sizeof(float) * 2 * (size_t)NP * DP : 0;
    s.W = o;      o += need_dx ? sizeof(float) * 2 * (size_t)DP * GP : 0;
    s.csr = o;    o += (need_dx && csr_smem) ? sizeof(int2) * (size_t)nnz : 0;
    s.rowptr = o; o += need_dx ? sizeof(int) * (size_t)(N + 1) : 0;
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

template <bool kCsrSmem>
__global__ void __launch_bounds__(kResThreads, 1)
resident_bwd_kernel(const ResBwdParams p, const ResSmemBwd lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* dOut_s = reinterpret_cast<float*>(smem + lay.dOut);
    float* Abuf = reinterpret_cast<float*>(smem + lay.A);
    float* Wsm = reinterpret_cast<float*>(smem + lay.W);
    int2* csr_s = reinterpret_cast<int2*>(smem + lay.csr);
    int* rowptr_s = reinterpret_cast<int*>(smem + lay.rowptr);

    const int tid = threadIdx.x, T = blockDim.x;
    const int q = blockIdx.x;
    const int N = p.N, NP = p.NP, D = p.D, DP = p.DP, V = p.V, G = p.G, GP = p.GP, GG = p.GG, K = p.K;
    const int slab = NP * DP;
    const bool need_dx = p.dx != nullptr;

    // ---- dOut tile of this sample: plain copy, or the max-pool (+ReLU) gradient routed on the fly
    if (p.dout) {
        const float* src = p.dout + (int64_t)q * N * G;
        for (int i = tid; i < NP * GP; i += T) {
            const int n = i / GP, g = i - n * GP;
            dOut_s[i] = (n < N && g < G) ? __ldg(src + (int64_t)n * G + g) : 0.f;
        }
    } else {
        const int pp = p.pool_p;
        const int64_t base = (int64_t)q * (N / pp) * G;
        for (int i = tid; i < NP * GP; i += T) {
            const int n = i / GP, g = i - n * GP;
            float v = 0.f;
            if (n < N && g < G) {
                const int m = n / pp, s = n - m * pp;
                const int64_t off = base + (int64_t)m * G + g;
                if ((int)p.idx[off] == s) {
                    v = __ldg(p.dy + off);
                    if (p.relu) {
                        const float yy = __ldg(p.y + off);      // relu(x_argmax): > 0 <=> x_argmax > 0, NaN <=> NaN
                        v = (yy > 0.f || yy != yy) ? v : 0.f;
                    }
                }
            }
            dOut_s[i] = v;
        }
    }
    if (need_dx) {
        for (int i = tid; i <= N; i += T) rowptr_s[i] = __ldg(p.rowptrT + i);
        if (kCsrSmem)
            for (int e = tid; e < p.nnz; e += T) csr_s[e] = make_int2(__ldg(p.colT + e), __float_as_int(__ldg(p.valT + e)));
    }
    __syncthreads();

    // ---- per-filter bias gradient of this sample: column sums of dOut in a fixed order
    float* red_s = reinterpret_cast<float*>(smem + lay.red);
    const int RL = T / GP;                                   // row lanes per column
    if (p.dbpart) {
        const int g = tid % GP, lr = tid / GP;
        float s = 0.f;
        if (lr < RL)
            for (int n = lr; n < N; n += RL) s += dOut_s[n * GP + g];
        red_s[tid] = s;
    }

    // ---- dW'_j[d][g] = sum_n P_j[n][d] dOut[n][g]: one (j, 4 d, 4 g) tile per thread at a time
    {
        const float4* dO4 = reinterpret_cast<const float4*>(dOut_s);
        const int per_j = V * GG;
        const int ntiles = K * per_j;
        for (int t = tid; t < ntiles; t += T) {
            const int j = t / per_j, rem = t - j * per_j;
            const int v = rem / GG, gg = rem - v * GG;
            const float4* st4 = reinterpret_cast<const float4*>(p.stack + ((int64_t)q * K + j) * slab) + v;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
            int n = 0;
            for (; n + 4 <= N; n += 4) {
                float4 pv[4], dv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pv[u] = __ldg(st4 + (int64_t)(n + u) * V);
#pragma unroll
                for (int u = 0; u < 4; ++u) dv[u] = dO4[(n + u) * GG + gg];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    fma4s(a0, pv[u].x, dv[u]);
                    fma4s(a1, pv[u].y, dv[u]);
                    fma4s(a2, pv[u].z, dv[u]);
                    fma4s(a3, pv[u].w, dv[u]);
                }
            }
            for (; n < N; ++n) {
                const float4 pv = __ldg(st4 + (int64_t)n * V);
                const float4 dv = dO4[n * GG + gg];
                fma4s(a0, pv.x, dv); fma4s(a1, pv.y, dv); fma4s(a2, pv.z, dv); fma4s(a3, pv.w, dv);
            }
            float4* dst = reinterpret_cast<float4*>(p.dWpart + (((int64_t)q * K + j) * DP + 4 * v) * GP) + gg;
            dst[0] = a0; dst[GG] = a1; dst[2 * GG] = a2; dst[3 * GG] = a3;
        }
    }
    if (p.dbpart) {
        __syncthreads();
        if (tid < GP) {
            float s = 0.f;
            for (int u = 0; u < RL; ++u) s += red_s[u * GP + tid];
            p.dbpart[(int64_t)q * GP + tid] = s;
        }
    }
    if (!need_dx) return;

    // ---- dx = sum_j (L~^T)^j (dOut W'_j^T): Horner, A_j = dOut W'_j^T + L~^T A_{j+1}; textbook
    // recursion: Clenshaw, B_j = dOut W_j^T + 2 L~^T B_{j+1} - B_{j+2}, dx = dOut W_0^T + L~^T B_1 - B_2.
    const int64_t inner = (int64_t)D * G;
    auto stage_wt = [&](int j, int buf) {     // transposed image Wt[g][d] so the 4 d of a tile are one float4
        float* dst = Wsm + buf * DP * GP;
        for (int i = tid; i < DP * GP; i += T) {
            const int g = i / DP, d = i - g * DP;
            dst[i] = (d < D && g < G) ? mixed_weight(p.W, K, inner, j, (int64_t)d * G + g, p.recursion) : 0.f;
        }
    };
    stage_wt(K - 1, (K - 1) & 1);
    __syncthreads();
    const bool cheb = p.recursion == TGCN_RECURSION_CHEBYSHEV;
    const int NR = NP / 4;                                   // tile rows are n, n+NR, n+2NR, n+3NR
    const int ntiles = NR * V;
    for (int j = K - 1; j >= 0; --j) {
        const float4* Wt4 = reinterpret_cast<const float4*>(Wsm + (j & 1) * DP * GP);
        const float4* dO4 = reinterpret_cast<const float4*>(dOut_s);
        // A_{j+1} lives in buffer (j+1)&1; A_j goes to buffer j&1 (which still holds A_{j+2})
        const float4* Ain4 = reinterpret_cast<const float4*>(Abuf + ((j + 1) & 1) * slab);
        float4* Aout4 = reinterpret_cast<float4*>(Abuf + (j & 1) * slab);
        for (int t = tid; t < ntiles; t += T) {
            const int r0 = t / V, v = t - r0 * V;
            float4 acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int gg = 0; gg < GG; ++gg) {
                const float4 w0 = Wt4[(4 * gg + 0) * V + v], w1 = Wt4[(4 * gg + 1) * V + v];
                const float4 w2 = Wt4[(4 * gg + 2) * V + v], w3 = Wt4[(4 * gg + 3) * V + v];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 b = dO4[(r0 + r * NR) * GG + gg];
                    fma4s(acc[r], b.x, w0);
                    fma4s(acc[r], b.y, w1);
                    fma4s(acc[r], b.z, w2);
                    fma4s(acc[r], b.w, w3);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int n = r0 + r * NR;
                if (n >= N) continue;
                if (j < K - 1) {
                    const float4 s = gather_row<kCsrSmem>(csr_s, p.colT, p.valT, rowptr_s[n], rowptr_s[n + 1], Ain4, V, v);
                    const float sc = (cheb && j >= 1) ? 2.f : 1.f;
                    acc[r].x = fmaf(sc, s.x, acc[r].x); acc[r].y = fmaf(sc, s.y, acc[r].y);
                    acc[r].z = fmaf(sc, s.z, acc[r].z); acc[r].w = fmaf(sc, s.w, acc[r].w);
                    if (cheb && j < K - 2) {
                        const float4 o = Aout4[n * V + v];
                        acc[r].x -= o.x; acc[r].y -= o.y; acc[r].z -= o.z; acc[r].w -= o.w;
                    }
                }
                if (j > 0) {
                    Aout4[n * V + v] = acc[r];
                } else {
                    float* dst = p.dx + ((int64_t)q * N + n) * D + 4 * v;
                    if ((D & 3) == 0) *reinterpret_cast<float4*>(dst) = acc[r];
                    else {
                        const float o[4] = {acc[r].x, acc[r].y, acc[r].z, acc[r].w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) if (4 * v + c < D) dst[c] = o[c];
                    }
                }
            }
        }
        if (j > 0) stage_wt(j - 1, (j - 1) & 1);
        __syncthreads();
    }
}

// dW[k] = sum_j M[k,j] sum_q dWpart[q][j]   (transposed mix, fp64, fixed order => deterministic), and
// the bias gradient from dout or from the routed pool gradient.
struct ResReduceParams {
    const float* dWpart; const float* dbpart; float* dW;
    const float* dout; const float* dy; const uint8_t* idx; const float* y;
    float* db;
    int Q, N, D, DP, G, GP, K, recursion, bias_mode, pool_p, relu;
    int w_blocks;          // blocks [0, w_blocks) reduce weights, the rest the bias
};

constexpr int kRedElems = 32;    // (d,g) elements per weight block
constexpr int kRedMaxK = 32;

__device__ __forceinline__ float routed_dout(const ResReduceParams& p, int q, int n, int g) {
    if (p.dout) return __ldg(p.dout + ((int64_t)q * p.N + n) * p.G + g);
    const int pp = p.pool_p;
    const int m = n / pp, s = n - m * pp;
    const int64_t off = ((int64_t)q * (p.N / pp) + m) * p.G + g;
    if ((int)p.idx[off] != s) return 0.f;
    float v = __ldg(p.dy + off);
    if (p.relu) {
        const float yy = __ldg(p.y + off);
        v = (yy > 0.f || yy != yy) ? v : 0.f;
    }
    return v;
}

__global__ void __launch_bounds__(kRedElems * 16)
resident_reduce_kernel(const ResReduceParams p) {
    __shared__ double red[kRedMaxK][kRedElems];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < p.w_blocks) {
        // thread (j, e): s_j[e] = sum_q part[q][j][e]; then dW[k][e] = sum_j M[k,j] s_j[e]
        const int e_local = tid % kRedElems, j = tid / kRedElems;          // j in [0,16)
        const int e = blockIdx.x * kRedElems + e_local;                    // element of the D x G matrix
        const int DG = p.D * p.G;
        const int d = e < DG ? e / p.G : 0, g = e < DG ? e - d * p.G : 0;
        for (int jj = j; jj < p.K; jj += 16) {
            double s = 0.0;
            if (e < DG) {
                const float* src = p.dWpart + ((int64_t)jj * p.DP + d) * p.GP + g;
                const int64_t qs = (int64_t)p.K * p.DP * p.GP;
                float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
                int q = 0;
                for (; q + 4 <= p.Q; q += 4) {
                    f0 += __ldg(src + (q + 0) * qs); f1 += __ldg(src + (q + 1) * qs);
                    f2 += __ldg(src + (q + 2) * qs); f3 += __ldg(src + (q + 3) * qs);
                }
                for (; q < p.Q; ++q) f0 += __ldg(src + q * qs);
                s = ((double)f0 + (double)f1) + ((double)f2 + (double)f3);
            }
            red[jj][e_local] = s;
        }
        __syncthreads();
        if (e < DG) {
            for (int k = j; k < p.K; k += 16) {
                double r = 0.0;
                if (p.recursion == TGCN_RECURSION_CHEBYSHEV) {
                    r = red[k][e_local];
                } else {
                    double sign = 1.0;
                    for (int j2 = k; j2 >= 0; j2 -= 2, sign = -sign) {
                        const double c = j2 < 2 ? 1.0 : 2.0;
                        r += sign * c * red[j2][e_local];
                    }
                }
                p.dW[(int64_t)k * DG + e] = (float)r;
            }
        }
        return;
    }
    // ---- bias gradient
    const int64_t i = (int64_t)(blockIdx.x - p.w_blocks) * blockDim.x + tid;
    if (p.bias_mode == TGCN_BIAS_PER_VERTEX) {
        if (i >= (int64_t)p.N * p.G) return;
        const int n = (int)(i / p.G), g = (int)(i - (int64_t)n * p.G);
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < p.Q; ++q) s += routed_dout(p, q, n, g);
        p.db[i] = s;
    } else if (p.bias_mode == TGCN_BIAS_PER_FILTER) {
        if (i >= p.G) return;
        float s = 0.f;
#pragma unroll 8
        for (int q = 0; q < p.Q; ++q) s += __ldg(p.dbpart + (int64_t)q * p.GP + i);
        p.db[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ResPlan { bool ok, csr_smem; int NP, DP, V, GP, GG, tpt; size_t smem; };

static ResPlan res_plan_fwd(int N, int D, int G, int K, int64_t nnz) {
    ResPlan pl{};
    if (N < 1 || D < 1 || G < 1 || K < 1 || K > kResMaxK || nnz < 0 || nnz > (int64_t)INT32_MAX / 2) return pl;
    pl.NP = round_up4(N); pl.DP = round_up4(D); pl.V = pl.DP / 4; pl.GP = round_up4(G); pl.GG = pl.GP / 4;
    if ((int64_t)pl.NP * pl.DP > (1 << 20)) return pl;
    const int ntiles = (pl.NP / 4) * pl.GG;
    pl.tpt = (ntiles + kResThreads - 1) / kResThreads;
    if (pl.tpt > 4) return pl;
    ResSmemFwd s = res_fwd_smem(N, (int)nnz, pl.DP, pl.GP, true);
    pl.csr_smem = s.total <= kResSmemLimit;
    if (!pl.csr_smem) s = res_fwd_smem(N, (int)nnz, pl.DP, pl.GP, false);
    pl.smem = s.total;
    pl.ok = s.total <= kResSmemLimit;
    return pl;
}

static ResPlan res_plan_bwd(int N, int D, int G, int K, int64_t nnz, bool need_dx) {
    ResPlan pl{};
    if (N < 1 || D < 1 || G < 1 || K < 1 || K > kResMaxK || nnz < 0 || nnz > (int64_t)INT32_MAX / 2) return pl;
    pl.NP = round_up4(N); pl.DP = round_up4(D); pl.V = pl.DP / 4; pl.GP = round_up4(G); pl.GG = pl.GP / 4;
    if ((int64_t)pl.NP * pl.DP > (1 << 20)) return pl;
    ResSmemBwd s = res_bwd_smem(N, (int)nnz, pl.DP, pl.GP, need_dx, true);
    pl.csr_smem = s.total <= kResSmemLimit;
    if (!pl.csr_smem) s = res_bwd_smem(N, (int)nnz, pl.DP, pl.GP, need_dx, false);
    pl.smem = s.total;
    pl.ok = s.total <= kResSmemLimit;
    return pl;
}

template <typename Kern>
static int res_set_smem(Kern kern, size_t bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResSmemLimit);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu): %s", name, bytes, cudaGetErrorString(e));
    return TGCN_OK;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_resident_supported(int N, int D, int G, int K, int64_t nnz) {
    return (res_plan_fwd(N, D, G, K, nnz).ok && res_plan_bwd(N, D, G, K, nnz, true).ok) ? 1 : 0;
}

extern "C" int64_t tgcn_resident_stack_bytes(int Q, int N, int D, int K) {
    if (Q < 0 || N < 0 || D < 1 || K < 1) return 0;
    return (int64_t)sizeof(float) * Q * K * round_up4(N) * round_up4(D);
}

extern "C" int64_t tgcn_resident_bwd_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return (int64_t)sizeof(float) * Q * ((int64_t)K * round_up4(D) * round_up4(G) + round_up4(G));
}

extern "C" int tgcn_resident_layer_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int N, int64_t nnz,
                                       const float* x, const float* W, const float* bias, int bias_mode,
                                       float* out, float* y, uint8_t* idx, int pool_p, int relu, float* stack,
                                       int Q, int D, int G, int K, int recursion, void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && G >= 1 && K >= 1, "tgcn_resident_layer_fwd: bad sizes");
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE || recursion == TGCN_RECURSION_CHEBYSHEV,
                 "tgcn_resident_layer_fwd: unknown recursion %d", recursion);
    if (Q == 0 || N == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptr && x && W, "tgcn_resident_layer_fwd: null pointer");
    TGCN_REQUIRE(out || y, "tgcn_resident_layer_fwd: neither out nor pooled output requested");
    TGCN_REQUIRE(bias_mode == TGCN_BIAS_NONE || bias, "tgcn_resident_layer_fwd: bias_mode %d without bias", bias_mode);
    if (y) {
        TGCN_SUPPORTED(pool_p == 2 || pool_p == 4, "tgcn_resident_layer_fwd: pool size %d", pool_p);
        TGCN_REQUIRE(N % pool_p == 0, "tgcn_resident_layer_fwd: vertex count %d not divisible by pool size %d", N, pool_p);
        TGCN_REQUIRE(idx, "tgcn_resident_layer_fwd: pooled output without idx");
    }
    const ResPlan pl = res_plan_fwd(N, D, G, K, nnz);
    TGCN_SUPPORTED(pl.ok, "tgcn_resident_layer_fwd: N=%d D=%d G=%d nnz=%lld does not fit shared memory", N, D, G, (long long)nnz);
    ResFwdParams p{};
    p.rowptr = rowptr; p.col = col; p.val = val; p.x = x; p.W = W; p.bias = bias; p.out = out; p.y = y; p.idx = idx;
    p.stack = stack; p.N = N; p.NP = pl.NP; p.nnz = (int)nnz; p.D = D; p.DP = pl.DP; p.V = pl.V; p.G = G; p.GP = pl.GP;
    p.GG = pl.GG; p.K = K; p.bias_mode = bias_mode; p.recursion = recursion; p.pool_p = y ? pool_p : 4; p.relu = relu;
    const ResSmemFwd lay = res_fwd_smem(N, (int)nnz, pl.DP, pl.GP, pl.csr_smem);
    cudaStream_t st = as_stream(stream);
#define TGCN_RES_FWD(TPT, CS)                                                                              \
    do {                                                                                                   \
        TGCN_PROPAGATE(res_set_smem(resident_fwd_kernel<TPT, CS>, lay.total, "resident_fwd"));             \
        resident_fwd_kernel<TPT, CS><<<(unsigned)Q, kResThreads, lay.total, st>>>(p, lay);                 \
    } while (0)
    if (pl.csr_smem) {
        switch (pl.tpt) { case 1: TGCN_RES_FWD(1, true); break; case 2: TGCN_RES_FWD(2, true); break;
                          case 3: TGCN_RES_FWD(3, true); break; default: TGCN_RES_FWD(4, true); break; }
    } else {
        switch (pl.tpt) { case 1: TGCN_RES_FWD(1, false); break; case 2: TGCN_RES_FWD(2, false); break;
                          case 3: TGCN_RES_FWD(3, false); break; default: TGCN_RES_FWD(4, false); break; }
    }
#undef TGCN_RES_FWD
    TGCN_LAUNCH_CHECK("resident_layer_fwd");
    return TGCN_OK;
}

extern "C" int tgcn_resident_layer_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N, int64_t nnz,
                                       const float* dout, const float* dy, const uint8_t* idx, const float* y,
                                       int pool_p, int relu, const float* stack, const float* W,
                                       float* dW, float* db, int bias_mode, float* dx, void* workspace,
                                       int Q, int D, int G, int K, int recursion, void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && G >= 1 && K >= 1, "tgcn_resident_layer_bwd: bad sizes");
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE || recursion == TGCN_RECURSION_CHEBYSHEV,
                 "tgcn_resident_layer_bwd: unknown recursion %d", recursion);
    TGCN_REQUIRE(dW, "tgcn_resident_layer_bwd: null dW");
    cudaStream_t st = as_stream(stream);
    if (Q == 0 || N == 0) {
        cudaMemsetAsync(dW, 0, sizeof(float) * K * D * G, st);
        if (db && bias_mode == TGCN_BIAS_PER_VERTEX) cudaMemsetAsync(db, 0, sizeof(float) * N * G, st);
        if (db && bias_mode == TGCN_BIAS_PER_FILTER) cudaMemsetAsync(db, 0, sizeof(float) * G, st);
        return TGCN_OK;
    }
    TGCN_REQUIRE(stack && W && workspace, "tgcn_resident_layer_bwd: null pointer");
    TGCN_REQUIRE((dout != nullptr) != (dy != nullptr), "tgcn_resident_layer_bwd: pass exactly one of dout / dy");
    if (dy) {
        TGCN_SUPPORTED(pool_p == 2 || pool_p == 4, "tgcn_resident_layer_bwd: pool size %d", pool_p);
        TGCN_REQUIRE(N % pool_p == 0 && idx, "tgcn_resident_layer_bwd: bad pooled gradient arguments");
        TGCN_REQUIRE(!relu || y, "tgcn_resident_layer_bwd: relu backward needs the pooled forward output");
    }
    TGCN_REQUIRE(bias_mode == TGCN_BIAS_NONE || db, "tgcn_resident_layer_bwd: bias_mode %d without db", bias_mode);
    TGCN_REQUIRE(!dx || rowptrT, "tgcn_resident_layer_bwd: dx requested without the CSR of L^T");
    const ResPlan pl = res_plan_bwd(N, D, G, K, nnz, dx != nullptr);
    TGCN_SUPPORTED(pl.ok, "tgcn_resident_layer_bwd: N=%d D=%d G=%d nnz=%lld does not fit shared memory", N, D, G, (long long)nnz);
    ResBwdParams p{};
    p.rowptrT = rowptrT; p.colT = colT; p.valT = valT; p.dout = dout; p.dy = dy; p.idx = idx; p.y = y; p.stack = stack;
    p.W = W; p.dWpart = reinterpret_cast<float*>(workspace); p.dx = dx;
    p.dbpart = bias_mode == TGCN_BIAS_PER_FILTER ? p.dWpart + (int64_t)Q * K * pl.DP * pl.GP : nullptr; p.N = N; p.NP = pl.NP; p.nnz = (int)nnz; p.D = D;
    p.DP = pl.DP; p.V = pl.V; p.G = G; p.GP = pl.GP; p.GG = pl.GG; p.K = K; p.recursion = recursion;
    p.pool_p = dy ? pool_p : 4; p.relu = relu;
    const ResSmemBwd lay = res_bwd_smem(N, (int)nnz, pl.DP, pl.GP, dx != nullptr, pl.csr_smem);
    if (pl.csr_smem) {
        TGCN_PROPAGATE(res_set_smem(resident_bwd_kernel<true>, lay.total, "resident_bwd"));
        resident_bwd_kernel<true><<<(unsigned)Q, kResThreads, lay.total, st>>>(p, lay);
    } else {
        TGCN_PROPAGATE(res_set_smem(resident_bwd_kernel<false>, lay.total, "resident_bwd"));
        resident_bwd_kernel<false><<<(unsigned)Q, kResThreads, lay.total, st>>>(p, lay);
    }
    TGCN_LAUNCH_CHECK("resident_layer_bwd");

    ResReduceParams r{};
    r.dWpart = p.dWpart; r.dbpart = p.dbpart; r.dW = dW; r.dout = dout; r.dy = dy; r.idx = idx; r.y = y; r.db = db;
    r.Q = Q; r.N = N; r.D = D; r.DP = pl.DP; r.G = G; r.GP = pl.GP; r.K = K; r.recursion = recursion;
    r.bias_mode = bias_mode; r.pool_p = p.pool_p; r.relu = relu;
    r.w_blocks = (int)ceil_div((int64_t)D * G, kRedElems);
    int b_blocks = 0;
    if (bias_mode == TGCN_BIAS_PER_VERTEX) b_blocks = (int)ceil_div((int64_t)N * G, kRedElems * 16);
    else if (bias_mode == TGCN_BIAS_PER_FILTER) b_blocks = (int)ceil_div(G, kRedElems * 16);
    resident_reduce_kernel<<<(unsigned)(r.w_blocks + b_blocks), kRedElems * 16, 0, st>>>(r);
    TGCN_LAUNCH_CHECK("resident_reduce");
    return TGCN_OK;
}
