// Large dense head: fc1 of the cortical-mesh model is a [Hd=200] x [I=167 424] weight (134 MB) applied to a batch
// of 8 samples (pytorch_hcp_tgcn.py:143-155 with n2 = 10 464, g2 = 64).  At that shape fc1 is pure weight
// streaming -- 2 flop per weight byte -- and the reference's autograd + optim.SGD touch the weight seven times per
// step (forward, dx, dW write, dW read, momentum read/write, weight read/write).  Here:
//
//   bighead_fc1_kernel   forward: W1 streamed ONCE (coalesced 16-byte loads, the 8 x-rows of a column block in
//                        registers), per-column-block partial sums [P][Q][Hd], reduced in a fixed order by
//   bighead_bn_kernel    which also applies the bias, BatchNorm1d (batch statistics), ReLU and the dropout;
//   bighead_bwd_kernel   backward + optimizer in one pass over W1: dx = dh W1 (old weights), the rank-Q gradient
//                        dW1 = dh^T x formed on the fly from x and dh held in shared memory, and -- when an update
//                        descriptor is given -- buf = momentum * buf + dW1 / world; W1 -= lr * buf applied in place:
//                        W1 and the momentum buffer are read once and written once, dW1 never exists in memory.
//                        Data parallel (world > 1): the gradient of a linear layer is low rank, so instead of an
//                        allreduce of the 134 MB gradient the ranks exchange the ACTIVATIONS: every rank publishes
//                        its x [Q, I] (5.4 MB) and dh [Q, Hd] in an IPC-mapped region (peer.cu's pack kernel + step
//                        flags) and every rank's update kernel reads all ranks' x / dh column blocks over NVLink
//                        (P2P loads) while it streams its own W1 -- the same sum in the same order on every rank, so
//                        the replicas stay bit-identical.  NVLink bytes per rank and step: (world-1) * 5.4 MB instead
//                        of 2 * (world-1)/world * 134 MB.
// fp32, fixed summation orders (deterministic).
#include <cstdlib>
#include "common.cuh"

namespace tgcn {

constexpr int kBhThreads = 256;
constexpr int kBhFwdCols = 1024;      // columns per CTA of the forward (8 warps x 32 lanes x float4)
constexpr int kBhFwdRows = 20;        // W1 rows per CTA of the forward (Hd = 200 -> 10 row groups: 164 x 10 CTAs = 3.7 waves of 3 CTAs/SM)
constexpr int kBhCols = 128;          // columns per CTA of the backward (32 lanes x float4)

// Sum each of the 8 per-lane values over the 32 lanes of the warp with 9 shuffles (recursive halving): afterwards
// every lane holds the total of value index ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1) in v[0].
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
    const unsigned full = 0xffffffffu;
    {
        const bool hi = (lane & 16) != 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float send = hi ? v[j] : v[j + 4];
            const float keep = hi ? v[j + 4] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    {
        const bool hi = (lane & 8) != 0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float send = hi ? v[j] : v[j + 2];
            const float keep = hi ? v[j + 2] : v[j];
            v[j] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    {
        const bool hi = (lane & 4) != 0;
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 4);
    }
    v[0] += __shfl_xor_sync(full, v[0], 2);
    v[0] += __shfl_xor_sync(full, v[0], 1);
    return v[0];
}

// grid (ceil(I / 1024), ceil(Hd / 20)); block 256.  Q <= 8; I % 4 == 0.
// The 20 weight rows of a CTA are read in 5 batches of 4 rows, double-buffered in registers: the loads of batch b+1 are
// issued before batch b is multiplied and reduced, so every warp keeps 2 KB of the weight stream in flight all the
// time (80 registers, 3 CTAs per SM).  (First version: 25 rows in batches of 8 -- the 4th batch held one real row and
// seven predicated ones -- loaded and then consumed in turn: 59.6 us for the 134 MB stream = 36 % of the measured HBM
// peak, stall reason long scoreboard, profiles/r02/ncu_bighead_fc1.txt.)
constexpr int kBhFwdBatch = 4;
static_assert(kBhFwdRows % kBhFwdBatch == 0, "row batches must tile the CTA's rows");

__global__ void __launch_bounds__(kBhThreads, 3)
bighead_fc1_kernel(const float* __restrict__ x, const float* __restrict__ W1, float* __restrict__ partial, int Q, int I,
                   int Hd) {
    __shared__ float red[8][kBhFwdRows][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * kBhFwdCols + warp * 128 + lane * 4;
    const bool ok = i < I;
    const int o0 = blockIdx.y * kBhFwdRows;
    const int rows = min(kBhFwdRows, Hd - o0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wcol = W1 + (int64_t)o0 * I + i;
    float4 w[2][kBhFwdBatch];
#pragma unroll
    for (int r = 0; r < kBhFwdBatch; ++r)
        w[0][r] = (ok && r < rows) ? __ldg(reinterpret_cast<const float4*>(wcol + (int64_t)r * I)) : zero4;
    float4 xq[8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
        xq[q] = (ok && q < Q) ? __ldg(reinterpret_cast<const float4*>(x + (int64_t)q * I + i)) : zero4;
    const int qsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
    for (int b = 0; b < kBhFwdRows / kBhFwdBatch; ++b) {
        if (b + 1 < kBhFwdRows / kBhFwdBatch) {
#pragma unroll
            for (int r = 0; r < kBhFwdBatch; ++r) {
                const int ol = (b + 1) * kBhFwdBatch + r;
                w[(b + 1) & 1][r] = (ok && ol < rows) ? __ldg(reinterpret_cast<const float4*>(wcol + (int64_t)ol * I)) : zero4;
            }
        }
#pragma unroll
        for (int r = 0; r < kBhFwdBatch; ++r) {
            const float4 wr = w[b & 1][r];
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                v[q] = fmaf(wr.x, xq[q].x, fmaf(wr.y, xq[q].y, fmaf(wr.z, xq[q].z, wr.w * xq[q].w)));
            const float tot = warp_reduce8(v, lane);
            if ((lane & 3) == 0 && b * kBhFwdBatch + r < rows) red[warp][b * kBhFwdBatch + r][qsel] = tot;
        }
    }
    __syncthreads();
    if (threadIdx.x < kBhFwdRows * 8) {
        const int ol = threadIdx.x >> 3, q = threadIdx.x & 7;
        if (ol < rows && q < Q) {
            float s = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) s += red[w8][ol][q];
            partial[((int64_t)q * Hd + o0 + ol) * gridDim.x + blockIdx.x] = s;   // [Q][Hd][P]: contiguous for the reduction
        }
    }
}

__device__ __forceinline__ float bh_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// h = b1 + sum_p partial[q][f][p] (fixed order: lane l sums p = l, l + 32, ...; then a butterfly), BatchNorm1d
// statistics, ReLU, dropout.  One CTA per hidden feature, one warp per sample (warps loop when Q > 8): the P partials
// of a (sample, feature) pair are contiguous, so the reduction is a handful of coalesced loads per lane instead of a
// chain of P dependent ones (the first version: 23 us at P = 164).  Q <= 64.
constexpr int kBhBnMaxQ = 64;
__global__ void __launch_bounds__(kBhThreads)
bighead_bn_kernel(const float* __restrict__ partial, int P, const float* __restrict__ b1, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* running_mean, float* running_var, float momentum, float eps,
                  int training, float* __restrict__ act, float* __restrict__ xhat, float* __restrict__ invstd_out, int Q,
                  int Hd, float drop_p, uint32_t drop_seed, const uint32_t* drop_step) {
    __shared__ float hs[kBhBnMaxQ];
    __shared__ float s_mean, s_inv;
    const DropCfg drop = drop_resolve(drop_p, drop_seed, drop_step);
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float bias = b1 ? __ldg(b1 + f) : 0.f;
    for (int q = warp; q < Q; q += kBhThreads / 32) {
        const float* src = partial + ((int64_t)q * Hd + f) * P;
        float s = 0.f;
        for (int p = lane; p < P; p += 32) s += __ldg(src + p);
        s = bh_warp_sum(s);
        if (lane == 0) hs[q] = s + bias;
    }
    __syncthreads();
    if (warp == 0) {
        float mean, inv;
        if (training) {
            float s = 0.f;
            for (int q = lane; q < Q; q += 32) s += hs[q];
            mean = bh_warp_sum(s) / (float)Q;
            float v = 0.f;
            for (int q = lane; q < Q; q += 32) { const float d = hs[q] - mean; v = fmaf(d, d, v); }
            const float var = bh_warp_sum(v) / (float)Q;
            inv = 1.0f / sqrtf(var + eps);
            if (lane == 0 && running_mean) {
                const float unb = Q > 1 ? var * (float)Q / (float)(Q - 1) : var;
                running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * mean;
                running_var[f] = (1.f - momentum) * running_var[f] + momentum * unb;
            }
        } else {
            mean = running_mean[f];
            inv = 1.0f / sqrtf(running_var[f] + eps);
        }
        if (lane == 0) { s_mean = mean; s_inv = inv; if (invstd_out) invstd_out[f] = inv; }
    }
    __syncthreads();
    for (int q = tid; q < Q; q += kBhThreads) {
        const float xh = (hs[q] - s_mean) * s_inv;
        const float y = fmaf(xh, gamma ? __ldg(gamma + f) : 1.f, beta ? __ldg(beta + f) : 0.f);
        if (xhat) xhat[(int64_t)q * Hd + f] = xh;
        float a = fmaxf(y, 0.f);
        if (drop.scale != 0.f) a = drop_apply(a, (uint64_t)((int64_t)q * Hd + f), drop);
        act[(int64_t)q * Hd + f] = a;
    }
}

// ---- backward (+ fused SGD-momentum update, + peer exchange of the activations) ---------------------------------
// world > 1: all ranks' [x | dh] of this step, pulled over NVLink into ONE local buffer by a plain, fully parallel copy
// kernel (every SM has dozens of 16-byte P2P loads in flight), so that the update kernel below reads local memory only:
// a first version read the peers' column blocks from inside the update kernel and paid an exposed ~3 us NVLink round
// trip at the start of each of its 1308 one-per-SM CTAs.
struct BhGatherParams {
    const float* flat[kPeerMaxWorld];   // each rank's region base ([flat0 | flat1 | flags])
    const unsigned int* flags;          // own flag line
    unsigned int* step_ctr;
    unsigned int* done_blocks;
    float* dst;                         // [world][n_flat]
    int64_t n_flat;
    int world, rank;
    unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(256)
bighead_gather_kernel(const BhGatherParams p) {
    const unsigned int step = *p.step_ctr;
    if ((int)threadIdx.x < p.world) peer_wait_flag(p.flags + threadIdx.x, step + 1u, p.timeout_ns);
    __syncthreads();
    const int64_t par = (int64_t)(step & 1u) * p.n_flat;
    const int64_t n4 = p.n_flat / 4, total = n4 * p.world;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / n4);
        const int64_t e = (i - (int64_t)r * n4) * 4;
        *reinterpret_cast<float4*>(p.dst + (int64_t)r * p.n_flat + e) = ld_volatile_f4(p.flat[r] + par + e);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(p.done_blocks, 1u);
        if (prev == gridDim.x - 1) { *p.done_blocks = 0u; *p.step_ctr = step + 1u; }
    }
}

struct BhBwdParams {
    const float* x[kPeerMaxWorld];      // rank r's x [Q, I]  (local memory: the own tensor, or the gathered copies)
    const float* dh[kPeerMaxWorld];     // rank r's dh [Q, Hd]
    const unsigned int* flags;          // unused (kept null): the exchange is complete before this kernel starts
    unsigned int* step_ctr;
    unsigned int* done_blocks;
    float* W1;                          // [Hd, I]; updated in place when `mom` != null
    float* mom;                         // [Hd, I] momentum buffer, or null (plain backward)
    float* dW1;                         // [Hd, I] gradient output of the plain backward, or null
    float* dx;                          // [Q, I] or null
    int world, rank, Q, I, Hd, HdP;
    int RW;                             // rows of W1 per warp (even; kBhBwdWarps * RW >= Hd)
    int64_t n_flat;                     // unused (kept for layout stability of the struct)
    float lr, momentum, gscale;
    unsigned long long timeout_ns;
};

// One CTA = 128 columns of W1, all Hd rows; lane = one float4 column, warp w = rows [w RW, (w + 1) RW) in tiles of
// OB = 10 rows.  The accumulators are packed over ROW PAIRS -- {g[2p][c], g[2p+1][c]} in one 64-bit register -- so the
// coefficient pairs {dh[s][2p], dh[s][2p+1]} come straight out of shared memory as the first operand of fma.rn.f32x2
// and only the four components of x need a duplicating move: 5 LDS.64 + 1 LDS.128 + 4 MOV + 20 FFMA2 per sample for
// 40 multiply-adds (a first version with scalar FFMA and an in-kernel repack spent 1.2 instructions per multiply-add;
// at 8 ranks = 64 samples per weight that arithmetic was +180 us per step).  dx uses the same trick over SAMPLE pairs.
// Shared memory: xs [SP][128] (SP = world * Q sample slots), dhs [SP][HdP] (row pairs contiguous), dho [HdP][QN] (this
// rank's dh, sample pairs contiguous); dhs is re-used for the cross-warp reduction of dx.
constexpr int kBhBwdWarps = 10;
constexpr int kBhBwdThreads = kBhBwdWarps * 32;
constexpr int kBhOB = 10;

__device__ __forceinline__ unsigned long long dup2(float v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <int QN, bool kUpdate>
__global__ void __launch_bounds__(kBhBwdThreads, 1)
bighead_bwd_kernel(const BhBwdParams p) {
    extern __shared__ __align__(16) float bsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Q = p.Q, I = p.I, Hd = p.Hd, HdP = p.HdP, S = p.world * Q;
    const int SP = (S + 3) & ~3;
    float* xs = bsm;                                    // [SP][128]
    float* dhs = bsm + (size_t)SP * kBhCols;            // [SP][HdP]
    float* dho = dhs + (size_t)SP * HdP;                // [HdP][QN]
    const int64_t i0 = (int64_t)blockIdx.x * kBhCols;
    const int64_t icol = i0 + lane * 4;
    const bool ok = icol < I;
    for (int t = tid; t < SP * 32; t += kBhBwdThreads) {
        const int sq = t >> 5, l = t & 31;
        const int r = sq / Q, q = sq - r * Q;
        const int64_t c = i0 + l * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sq < S && c < I) v = __ldg(reinterpret_cast<const float4*>(p.x[r] + (int64_t)q * I + c));
        reinterpret_cast<float4*>(xs)[t] = v;
    }
    for (int t = tid; t < SP * HdP; t += kBhBwdThreads) {
        const int sq = t / HdP, f = t - sq * HdP;
        const int r = sq / Q, q = sq - r * Q;
        dhs[t] = (sq < S && f < Hd) ? __ldg(p.dh[r] + (int64_t)q * Hd + f) : 0.f;
    }
    for (int t = tid; t < HdP * QN; t += kBhBwdThreads) {
        const int f = t / QN, q = t - f * QN;
        dho[t] = (q < Q && f < Hd) ? __ldg(p.dh[p.rank] + (int64_t)q * Hd + f) : 0.f;
    }
    __syncthreads();

    unsigned long long dxa[QN / 2][4];                  // {dx[2qp][c], dx[2qp+1][c]}
#pragma unroll
    for (int qp = 0; qp < QN / 2; ++qp)
#pragma unroll
        for (int c = 0; c < 4; ++c) dxa[qp][c] = 0ull;
    const float4* xs4 = reinterpret_cast<const float4*>(xs);
    const int RW = p.RW;                                // rows per warp (even)
    const int row_end = min(Hd, (warp + 1) * RW);
    for (int o0 = warp * RW; o0 < row_end; o0 += kBhOB) {
        float4 w[kBhOB], m[kBhOB];
#pragma unroll
        for (int j = 0; j < kBhOB; ++j) {
            const bool live = ok && o0 + j < row_end;
            w[j] = live ? *reinterpret_cast<const float4*>(p.W1 + (int64_t)(o0 + j) * I + icol) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (kUpdate) m[j] = live ? *reinterpret_cast<const float4*>(p.mom + (int64_t)(o0 + j) * I + icol) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        unsigned long long g[kBhOB / 2][4];
#pragma unroll
        for (int jp = 0; jp < kBhOB / 2; ++jp)
#pragma unroll
            for (int c = 0; c < 4; ++c) g[jp][c] = 0ull;
        for (int sq = 0; sq < S; ++sq) {                     // rank-major, then sample: the same order on every rank
            const float4 xv = xs4[sq * 32 + lane];
            const unsigned long long xx[4] = {dup2(xv.x), dup2(xv.y), dup2(xv.z), dup2(xv.w)};
            const unsigned long long* dr = reinterpret_cast<const unsigned long long*>(dhs + (size_t)sq * HdP + o0);   // o0 even, HdP even
#pragma unroll
            for (int jp = 0; jp < kBhOB / 2; ++jp) {
                const unsigned long long d2 = dr[jp];        // {dh[sq][o0 + 2 jp], dh[sq][o0 + 2 jp + 1]}: warp broadcast
#pragma unroll
                for (int c = 0; c < 4; ++c) ffma2(g[jp][c], d2, xx[c]);
            }
        }
        if (p.dx) {
#pragma unroll
            for (int j = 0; j < kBhOB; ++j) {
                const unsigned long long ww[4] = {dup2(w[j].x), dup2(w[j].y), dup2(w[j].z), dup2(w[j].w)};
                const unsigned long long* dq = reinterpret_cast<const unsigned long long*>(dho + (size_t)(o0 + j) * QN);
#pragma unroll
                for (int qp = 0; qp < QN / 2; ++qp) {
                    const unsigned long long d2 = dq[qp];    // {dh[2qp][o], dh[2qp+1][o]} of this rank
#pragma unroll
                    for (int c = 0; c < 4; ++c) ffma2(dxa[qp][c], d2, ww[c]);
                }
            }
        }
#pragma unroll
        for (int jp = 0; jp < kBhOB / 2; ++jp) {
            float ga[4], gb[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) unpack2(g[jp][c], ga[c], gb[c]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 2 * jp + h;
                if (!(ok && o0 + j < row_end)) continue;
                const float* gg = h ? gb : ga;
                const int64_t off = (int64_t)(o0 + j) * I + icol;
                if (kUpdate) {
                    float4 mm, wn;
                    mm.x = fmaf(p.momentum, m[j].x, gg[0] * p.gscale); mm.y = fmaf(p.momentum, m[j].y, gg[1] * p.gscale);
                    mm.z = fmaf(p.momentum, m[j].z, gg[2] * p.gscale); mm.w = fmaf(p.momentum, m[j].w, gg[3] * p.gscale);
                    wn.x = fmaf(-p.lr, mm.x, w[j].x); wn.y = fmaf(-p.lr, mm.y, w[j].y);
                    wn.z = fmaf(-p.lr, mm.z, w[j].z); wn.w = fmaf(-p.lr, mm.w, w[j].w);
                    *reinterpret_cast<float4*>(p.mom + off) = mm;
                    *reinterpret_cast<float4*>(p.W1 + off) = wn;
                } else {
                    *reinterpret_cast<float4*>(p.dW1 + off) = make_float4(gg[0], gg[1], gg[2], gg[3]);
                }
            }
        }
    }
    if (p.dx) {
        __syncthreads();                                    // dhs / dho are dead: re-use them for the cross-warp reduction of dx
        float4* red = reinterpret_cast<float4*>(dhs);       // [warps][Q][32 lanes]
#pragma unroll
        for (int qp = 0; qp < QN / 2; ++qp) {
            float a[4], b[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) unpack2(dxa[qp][c], a[c], b[c]);
            if (2 * qp < Q) red[(warp * Q + 2 * qp) * 32 + lane] = make_float4(a[0], a[1], a[2], a[3]);
            if (2 * qp + 1 < Q) red[(warp * Q + 2 * qp + 1) * 32 + lane] = make_float4(b[0], b[1], b[2], b[3]);
        }
        __syncthreads();
        for (int t = tid; t < Q * 32; t += kBhBwdThreads) {
            const int q = t >> 5, l = t & 31;
            float4 s4 = red[(0 * Q + q) * 32 + l];
#pragma unroll
            for (int w8 = 1; w8 < kBhBwdWarps; ++w8) {
                const float4 v = red[(w8 * Q + q) * 32 + l];
                s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            }
            const int64_t c = i0 + l * 4;
            if (c < I) *reinterpret_cast<float4*>(p.dx + (int64_t)q * I + c) = s4;
        }
    }
}

// peer.cu: copies this rank's tensors into flat[step & 1] of its region and publishes the step flag to every rank
int peer_pack_launch(void* const* regions_host, int world, int rank, const float* const* srcs_host, const int64_t* numels_host,
                     int nseg, unsigned int* state, cudaStream_t st);

int bighead_workspace_chunks(int I) { return (int)ceil_div(I, kBhFwdCols); }

bool bighead_applies(int Q, int I, int Hd) {
    return (int64_t)I * Hd > ((int64_t)1 << 21) && Q >= 1 && Q <= 8 && (I % 4) == 0;
}

int bighead_fwd(const float* x, const float* W1, const float* b1, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float momentum, float eps, int training, float drop_p, uint32_t drop_seed,
                const uint32_t* drop_step, float* act, float* xhat, float* invstd, float* partial, int Q, int I, int Hd,
                cudaStream_t st) {
    TGCN_REQUIRE(aligned16(x) && aligned16(W1), "tgcn_head_fwd: x / W1 must be 16-byte aligned");
    const int P = bighead_workspace_chunks(I);
    const dim3 grid((unsigned)P, (unsigned)ceil_div(Hd, kBhFwdRows));
    bighead_fc1_kernel<<<grid, kBhThreads, 0, st>>>(x, W1, partial, Q, I, Hd);
    TGCN_LAUNCH_CHECK("bighead_fc1");
    bighead_bn_kernel<<<(unsigned)Hd, kBhThreads, 0, st>>>(partial, P, b1, gamma, beta, running_mean, running_var,
                                                                            momentum, eps, training, act, xhat, invstd, Q, Hd,
                                                                            drop_p, drop_seed, drop_step);
    TGCN_LAUNCH_CHECK("bighead_bn");
    return TGCN_OK;
}

template <bool kUpdate>
static int bighead_bwd_launch(const BhBwdParams& p, size_t smem, cudaStream_t st) {
    auto kern = bighead_bwd_kernel<8, kUpdate>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_head_bwd: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    }
    kern<<<(unsigned)ceil_div(p.I, kBhCols), kBhBwdThreads, smem, st>>>(p);
    TGCN_LAUNCH_CHECK("bighead_bwd");
    return TGCN_OK;
}

// dh [Q, Hd] comes from head_bwd1_kernel.  upd == null: plain backward (dW1, dx).  upd != null: fused update.
int bighead_bwd(const float* dh, const float* x, float* W1, float* dW1, float* dx, const tgcn_fc1_update_t* upd, int Q, int I,
                int Hd, cudaStream_t st) {
    TGCN_REQUIRE(aligned16(x) && aligned16(W1) && (!dW1 || aligned16(dW1)) && (!dx || aligned16(dx)),
                 "tgcn_head_bwd: x / W1 / dW1 / dx must be 16-byte aligned");
    const int world = upd ? upd->world : 1, rank = upd ? upd->rank : 0;
    TGCN_SUPPORTED(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "tgcn_head_bwd: world %d rank %d", world, rank);
    BhBwdParams p{};
    p.world = world; p.rank = rank; p.Q = Q; p.I = I; p.Hd = Hd;
    p.RW = ((int)ceil_div(Hd, kBhBwdWarps) + 1) & ~1;                 // even: row pairs never straddle two warps
    p.HdP = kBhBwdWarps * p.RW + kBhOB;                               // a warp's last tile may read (zero) rows past its range
    p.W1 = W1; p.dW1 = dW1; p.dx = dx;
    p.timeout_ns = peer_timeout_ns();
    if (upd) {
        TGCN_REQUIRE(upd->mom && aligned16(upd->mom), "tgcn_head_bwd: update descriptor without a 16-byte aligned momentum buffer");
        p.mom = upd->mom; p.lr = upd->lr; p.momentum = upd->momentum; p.gscale = 1.0f / (float)world;
    } else {
        TGCN_REQUIRE(dW1, "tgcn_head_bwd: neither dW1 nor an update descriptor");
    }
    int S = world * Q;
    if (world == 1 && upd) { const char* e = getenv("TGCN_BH_FAKEWORLD"); const int f = e ? atoi(e) : 0; if (f > 1 && f <= kPeerMaxWorld) S = f * Q; }
    const int SP = (S + 3) & ~3;
    const size_t smem_stage = sizeof(float) * ((size_t)SP * kBhCols + (size_t)SP * p.HdP + (size_t)p.HdP * 8);
    const size_t smem_red = sizeof(float) * ((size_t)SP * kBhCols + (size_t)kBhBwdWarps * Q * kBhCols);
    const size_t smem = smem_stage > smem_red ? smem_stage : smem_red;
    TGCN_SUPPORTED(smem <= 200 * 1024, "tgcn_head_bwd: world %d x batch %d x Hd %d does not fit shared memory", world, Q, Hd);
    if (world == 1) {
        p.x[0] = x; p.dh[0] = dh;
        // timing experiment only (scripts/time_kernels.py): the arithmetic of an N-rank update on one GPU
        static const int fake = [] { const char* e = getenv("TGCN_BH_FAKEWORLD"); return e ? atoi(e) : 0; }();
        if (upd && fake > 1 && fake <= kPeerMaxWorld) {
            for (int r = 1; r < fake; ++r) { p.x[r] = x; p.dh[r] = dh; }
            p.world = fake; p.gscale = 1.0f / (float)fake;
        }
    } else {
        TGCN_REQUIRE(upd->regions && upd->state && upd->gather, "tgcn_head_bwd: world > 1 needs the peer regions, the state and the gather buffer");
        TGCN_REQUIRE(aligned16(upd->gather), "tgcn_head_bwd: gather buffer must be 16-byte aligned");
        // publish x and dh of this step in the own region, pull every rank's copy into the local gather buffer
        const float* srcs[2] = {x, dh};
        const int64_t numels[2] = {(int64_t)Q * I, (int64_t)Q * Hd};
        TGCN_PROPAGATE(peer_pack_launch(upd->regions, world, rank, srcs, numels, 2, upd->state, st));
        const int64_t nx = ((int64_t)Q * I + 3) & ~(int64_t)3, nd = ((int64_t)Q * Hd + 3) & ~(int64_t)3;
        BhGatherParams g{};
        g.n_flat = nx + nd; g.world = world; g.rank = rank; g.timeout_ns = peer_timeout_ns();
        for (int r = 0; r < world; ++r) g.flat[r] = reinterpret_cast<const float*>(upd->regions[r]);
        g.flags = reinterpret_cast<const unsigned int*>(reinterpret_cast<const char*>(upd->regions[rank]) + 2 * g.n_flat * sizeof(float));
        g.step_ctr = upd->state; g.done_blocks = upd->state + 2; g.dst = upd->gather;
        const int gblocks = (int)min64(ceil_div(g.n_flat / 4 * world, 256), (int64_t)kNumSMs * 8);
        bighead_gather_kernel<<<gblocks, 256, 0, st>>>(g);
        TGCN_LAUNCH_CHECK("bighead_gather");
        for (int r = 0; r < world; ++r) {
            p.x[r] = upd->gather + (int64_t)r * g.n_flat;
            p.dh[r] = upd->gather + (int64_t)r * g.n_flat + nx;
        }
    }
    if (upd) return bighead_bwd_launch<true>(p, smem, st);
    return bighead_bwd_launch<false>(p, smem, st);
}

}  // namespace tgcn
