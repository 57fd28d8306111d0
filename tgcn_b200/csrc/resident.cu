// Sample-resident fused layer kernels (TGCN_ENGINE_RESIDENT).
//
// The columns of the [N, Q*D] slab are independent through the whole recursion, so when one
// sample's slab [N, D] fits in shared memory the complete layer runs out of shared memory:
//
//   forward : P_0 = x_q;  P_j = L~ P_{j-1}  (packed-CSR gather, shared -> shared, ping-pong);
//             out += P_j * W'_j after every step (accumulators live in registers for all K steps);
//             the epilogue adds the bias and optionally applies ReLU + the permuted max-pool, so the
//             un-pooled activation never reaches HBM.  Every P_j is written to HBM exactly once
//             (fire-and-forget) for the backward.  Small batches split a sample by rows over the
//             2 or 4 CTAs of a thread-block cluster: each CTA computes its rows and pushes them into
//             the peers' copies of the slab with st.async (distributed shared memory, completion
//             counted on the receiver's mbarrier) -- no cluster-wide barrier inside the loop.
//   backward: dW'_j = P_j^T dOut with the saved basis streamed HBM -> shared memory by 1-D bulk
//             copies (cp.async.bulk, mbarrier ring) and dOut resident in shared memory (the max-pool /
//             ReLU gradient routing is applied while it is staged);
//             dx = sum_j (L~^T)^j (dOut W'_j^T) by Horner's rule on a resident accumulator.
//             Per-sample partials are reduced over the batch by a second, deterministic kernel
//             that also applies the transposed weight mix and produces the bias gradient.
//
// The reference computes the same quantities with K-1 dense `bmm` launches plus einsum/permute
// copies per layer (tgcn/nn/gcn.py:108-154, :189-237) and autograd's stored intermediates.
// Work is fp32 FFMA + shared-memory gathers: at these sizes (Q*N*K*D*G <= ~0.3 GFLOP) the
// contraction is far below one SM-microsecond of tensor-core work and staging hi/lo TF32 operand
// images would cost more shared-memory traffic than the FFMAs it replaces; the tcgen05 engine
// (contract_tc*.cu) serves the large-graph configs where the stack streams from HBM.
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace tgcn {
using namespace tc;

constexpr size_t kResSmemLimit = 227 * 1024;
constexpr int kResMaxK = 32;
constexpr int kBwdThreads = 512;
constexpr int kRingStages = 3;

__host__ __device__ inline int round_up4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline int round_up2(int v) { return (v + 1) & ~1; }
__host__ __device__ inline int round_up16(int v) { return (v + 15) & ~15; }

__device__ __forceinline__ void fma4s(float4& acc, float w, const float4& x) { fma4_packed(acc, w, x); }

// ------------------------------------------------------------------------------------------------
// weight images: Wm[j][d][g] ([K][DP][GP], zero padded) = W'_j of the reference recursion
// (W'_j = sum_k M[k,j] W_k, see mix_weights_kernel in contract.cu; plain copy for the textbook
// recursion) and its transpose Wt[j][g][d] ([K][GP][DP]) for the dx recursion.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resident_prep_kernel(const float* __restrict__ W, float* __restrict__ Wm, float* __restrict__ Wt, uint8_t* __restrict__ Wtc,
                     int K, int D, int G, int DP, int GP, int GP16, int recursion) {
    // one thread per (d, g) of the padded 32 x GP16 tile the tensor-core image covers
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int rows_d = DP > 32 ? DP : 32;
    if (i >= rows_d * GP16) return;
    const int d = i / GP16, g = i - d * GP16;
    const bool valid = d < D && g < G;
    float w[kResMaxK];
#pragma unroll
    for (int k = 0; k < kResMaxK; ++k) w[k] = (valid && k < K) ? __ldg(W + ((int64_t)k * D + d) * G + g) : 0.f;
#pragma unroll
    for (int j = 0; j < kResMaxK; ++j) {
        if (j >= K) break;
        float m;
        if (recursion == TGCN_RECURSION_CHEBYSHEV) {
            m = w[j];
        } else {
            const double c = j < 2 ? 1.0 : 2.0;
            double s = 0.0, sign = 1.0;
#pragma unroll
            for (int k = j; k < kResMaxK; k += 2) {
                if (k < K) s += sign * c * (double)w[k];
                sign = -sign;
            }
            m = (float)s;
        }
        if (d < DP && g < GP) {
            Wm[((int64_t)j * DP + d) * GP + g] = m;
            Wt[((int64_t)j * GP + g) * DP + d] = m;
        }
        if (Wtc && d < 32) {
            // K-major B operand of the contraction: rows = g, 32 fp32 of d per 128-byte row, SWIZZLE_128B; hi then lo
            uint8_t* base = Wtc + (int64_t)j * 2 * GP16 * kRowBytes;
            const uint32_t off = sw128_offset((uint32_t)g, (uint32_t)d);
            const float h = tf32_hi(m);
            *reinterpret_cast<float*>(base + off) = h;
            *reinterpret_cast<float*>(base + (int64_t)GP16 * kRowBytes + off) = m - h;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed CSR: rowinfo[n] = (start, len), start even so a row's (col, val-bits) pairs can be read two
// at a time as one 16-byte vector; within a row the entries are ordered so that the m rows that
// share a quarter-warp hit different shared-memory bank groups at the same loop position (the
// bank group of neighbour row c is c mod m when a CTA keeps 8/m float4 per row).
// ------------------------------------------------------------------------------------------------
template <bool kSmem>
__device__ __forceinline__ int4 ld_pair(const int2* p) {
    if (kSmem) return *reinterpret_cast<const int4*>(p);
    return __ldg(reinterpret_cast<const int4*>(p));
}
template <bool kSmem>
__device__ __forceinline__ int2 ld_one(const int2* p) {
    if (kSmem) return *p;
    return __ldg(p);
}

// a0 / a1 += val[e] * in[col[e]][v] over entries [k0, len) of one row (k0 even): even positions go to a0, odd
// positions to a1 -- the ONE summation order every resident variant follows, so results do not depend on the
// variant (or on the batch / cluster split that selects it)
template <bool kCsrSmem>
__device__ __forceinline__ void gather_accum(const int2* __restrict__ ent, const int2 info, int k,
                                             const float4* __restrict__ in4, const int V, const int v,
                                             float4& a0, float4& a1) {
    const int2* e = ent + info.x;
    for (; k + 4 <= info.y; k += 4) {
        const int4 p01 = ld_pair<kCsrSmem>(e + k), p23 = ld_pair<kCsrSmem>(e + k + 2);
        const float4 x0 = in4[p01.x * V + v], x1 = in4[p01.z * V + v], x2 = in4[p23.x * V + v], x3 = in4[p23.z * V + v];
        fma4s(a0, __int_as_float(p01.y), x0);
        fma4s(a1, __int_as_float(p01.w), x1);
        fma4s(a0, __int_as_float(p23.y), x2);
        fma4s(a1, __int_as_float(p23.w), x3);
    }
    if (k + 2 <= info.y) {
        const int4 p01 = ld_pair<kCsrSmem>(e + k);
        const float4 x0 = in4[p01.x * V + v], x1 = in4[p01.z * V + v];
        fma4s(a0, __int_as_float(p01.y), x0);
        fma4s(a1, __int_as_float(p01.w), x1);
        k += 2;
    }
    if (k < info.y) {
        const int2 p0 = ld_one<kCsrSmem>(e + k);
        fma4s(a0, __int_as_float(p0.y), in4[p0.x * V + v]);
    }
}

// sum_e val[e] * in[col[e]][v] for one row and one float4 column group (fixed summation order)
template <bool kCsrSmem>
__device__ __forceinline__ float4 gather_row(const int2* __restrict__ ent, const int2 info,
                                             const float4* __restrict__ in4, const int V, const int v) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    gather_accum<kCsrSmem>(ent, info, 0, in4, V, v, a0, a1);
    return make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
}

// torch.max(dim) rule (pool.cu): first maximal element wins, NaN beats any number.
__device__ __forceinline__ void take_max_r(float cand, int s, float& best, int& arg) {
    const bool better = (cand > best) || (cand != cand && best == best);
    if (better) { best = cand; arg = s; }
}

// ---- thread-block cluster helpers (a sample may be split by rows over CL CTAs of one cluster)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16-byte store into a peer CTA's shared memory; completes 16 transaction bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, const float4& v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar) : "memory");
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct ResFwdParams {
    const int2* rowinfo; const int2* entries;   // packed CSR of L~
    const float* x;        // [Q,N,D]
    const float* Wm;       // [K][DP][GP] mixed, padded weights (resident_prep_kernel)
    const uint8_t* Wtc;    // [K][hi|lo][GP16][128 B] tensor-core image of the same weights (kTC)
    const float* bias;     // [N,G] / [G] / null
    float* out;            // [Q,N,G] or null
    float* y;              // [Q,N/p,G] or null (fused pool)
    uint8_t* idx;          // [Q,N/p,G]
    float* stack;          // [Q][NP][K][DP] or null (inference)
    int N, NP, E, D, DP, V, G, GP, GG, K;
    int bias_mode, recursion, pool_p, relu;
    float drop_p; uint32_t drop_seed; const uint32_t* drop_step;   // dropout between the ReLU and the pool (common.cuh)
    int CL, rg_per;        // CTAs per sample (cluster size) and row groups (of 4 rows) per CTA
    int GP16, MT;          // kTC: padded filter count of the MMA and number of 128-row tiles of this CTA's rows
};

struct ResSmemFwd { size_t aimg, wimg, bars, P, W, csr, rowinfo, total; };

constexpr uint32_t kTcTile = 128 * 128;      // one 128-row x 32-fp32 operand tile

static ResSmemFwd res_fwd_smem(int N, int64_t E, int DP, int GP, int K, bool csr_smem, bool w_smem, int tc_tiles = 0,
                               int GP16 = 0) {
    ResSmemFwd s{};
    const int NP = round_up4(N);
    size_t o = 0;
    if (tc_tiles > 0) {       // 1024-byte aligned operand images first (the kernel aligns its window, +1024 slack)
        s.aimg = o;    o += (size_t)tc_tiles * 2 * kTcTile;
        s.wimg = o;    o += (size_t)2 * 2 * GP16 * kRowBytes;
        o += 1024;
    }
    s.bars = o;    o += 64;
    s.P = o;       o += sizeof(float) * 2 * (size_t)NP * DP;
    s.W = o;       o += w_smem ? sizeof(float) * (size_t)K * DP * GP : 0;
    s.csr = o;     o += csr_smem ? sizeof(int2) * (size_t)E : 0;
    s.rowinfo = o; o += sizeof(int2) * (size_t)round_up2(N);
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

// tcgen05 path of the contraction (kTC): the SpMM threads also write the hi/lo TF32 images of their new rows into a
// SWIZZLE_128B operand tile; one thread issues, per recursion step, 3 MMAs (lo*hi, hi*lo, hi*hi: 3xTF32) per
// 8-wide k-step and 128-row tile against the step's weight image (bulk-copied, double buffered); the layer output
// accumulates in TMEM across all K steps while the next SpMM step runs; the epilogue reads it back with tcgen05.ld.
__device__ __forceinline__ void store_image4(uint8_t* hi, uint8_t* lo, uint32_t off, const float4& r) {
    float4 h, l;
    h.x = tf32_hi(r.x); h.y = tf32_hi(r.y); h.z = tf32_hi(r.z); h.w = tf32_hi(r.w);
    l.x = tf32_hi(r.x - h.x); l.y = tf32_hi(r.y - h.y); l.z = tf32_hi(r.z - h.z); l.w = tf32_hi(r.w - h.w);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

// ENT > 0 (needs at most one (row, float4) item per thread): the first ENT packed entries of the thread's row live
// in registers for all K steps -- the CSR operand is read once per launch instead of once per step, and the
// unrolled gathers of a step are all in flight together.
template <int THREADS, int TPT, bool kCsrSmem, bool kWSmem, bool kTC, int ENT = 0>
__global__ void __launch_bounds__(THREADS, 1)
resident_fwd_kernel(const ResFwdParams p, const ResSmemFwd lay) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* smem = smem_raw;
    if (kTC) { const uint32_t a = smem_u32(smem_raw); smem = smem_raw + (((a + 1023u) & ~1023u) - a); }
    uint8_t* aimg = smem + lay.aimg;                                   // [MT][hi 16 KB | lo 16 KB]
    uint8_t* wimg = smem + lay.wimg;                                   // 2 stages x [hi | lo] x GP16 x 128 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);     // [0] prologue loads, [1],[2] peer rows,
                                                                       // [3],[4] weight image stages, [5] MMAs done
    float* Pbuf = reinterpret_cast<float*>(smem + lay.P);
    float* Wsm = reinterpret_cast<float*>(smem + lay.W);
    int2* csr_s = reinterpret_cast<int2*>(smem + lay.csr);
    int2* rowinfo_s = reinterpret_cast<int2*>(smem + lay.rowinfo);

    const int tid = threadIdx.x;
    constexpr int T = THREADS;
    const int CL = p.CL;
    const int q = blockIdx.x / CL, c = blockIdx.x - q * CL;    // 1-D clusters: c is the rank in the cluster
    const int N = p.N, NP = p.NP, D = p.D, DP = p.DP, V = p.V, G = p.G, GG = p.GG, K = p.K;
    const DropCfg drop = drop_resolve((p.relu && p.y) ? p.drop_p : 0.f, p.drop_seed, p.drop_step);
    const int slab = NP * DP;                                  // floats per P buffer
    const int rg0 = min(c * p.rg_per, NP / 4), rg1 = min(NP / 4, rg0 + p.rg_per);   // a trailing rank may own no rows
    const int row0 = rg0 * 4, row1 = min(N, rg1 * 4);          // rows this CTA computes
    const int wslab = DP * p.GP;

    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[6]);
    const uint32_t wstage = 2u * (uint32_t)p.GP16 * kRowBytes;        // one weight image: hi + lo
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    uint32_t tmem_acc = 0;
    if (kTC) {
        if (tid < 32) tmem_alloc(tmem_slot, tmem_cols_pow2((uint32_t)(p.MT * p.GP16)));
        tcgen05_fence_before();
    }
    __syncthreads();
    if (kTC) {
        tcgen05_fence_after();
        tmem_acc = *tmem_slot;
        // rows beyond this CTA's share and the k-padding columns of the A image must be finite: zero it once
        float4* z = reinterpret_cast<float4*>(aimg);
        for (int i = tid; i < p.MT * 2 * (int)(kTcTile / 16); i += T) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[3], wstage);
            bulk_g2s(wimg, p.Wtc, wstage, &bars[3]);
        }
    }
    if (tid == 0) {   // operands that are used as they are: bulk copies straight into shared memory
        const uint32_t b_ri = (uint32_t)sizeof(int2) * (uint32_t)round_up2(N);
        const uint32_t b_csr = kCsrSmem ? (uint32_t)sizeof(int2) * (uint32_t)p.E : 0u;
        const uint32_t b_w = kWSmem ? (uint32_t)sizeof(float) * (uint32_t)(K * wslab) : 0u;
        mbar_arrive_expect_tx(&bars[0], b_ri + b_csr + b_w);
        bulk_g2s(rowinfo_s, p.rowinfo, b_ri, &bars[0]);
        if (b_csr) bulk_g2s(csr_s, p.entries, b_csr, &bars[0]);
        if (b_w) bulk_g2s(Wsm, p.Wm, b_w, &bars[0]);
    }
    const int2* ent = kCsrSmem ? csr_s : p.entries;
    {   // P_0 = x_q (zero pad columns / rows): every CTA of the cluster keeps the full slab; the owner
        // of a row also writes it to the saved stack [q][n][j][DP]
        const float* xq = p.x + (int64_t)q * N * D;
        float* stq = p.stack ? p.stack + (int64_t)q * NP * K * DP : nullptr;
        if (DP != D || NP != N) {
            for (int i = tid; i < slab; i += T) {
                const int n = i / DP, d = i - n * DP;
                if (n >= N || d >= D) Pbuf[i] = 0.f;
            }
        }
        const int total = N * D;
        if ((total & 3) == 0 && aligned16(p.x)) {
            const float4* x4 = reinterpret_cast<const float4*>(xq);
            for (int i4 = tid; i4 < total / 4; i4 += T) {
                const float4 v = __ldg(x4 + i4);
                const float vv[4] = {v.x, v.y, v.z, v.w};
                if (DP == D) {
                    reinterpret_cast<float4*>(Pbuf)[i4] = v;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = 4 * i4 + u, n = e / D, d = e - n * D;
                        Pbuf[n * DP + d] = vv[u];
                    }
                }
            }
        } else {
            for (int e = tid; e < total; e += T) {
                const int n = e / D, d = e - n * D;
                Pbuf[n * DP + d] = __ldg(xq + e);
            }
        }
        __syncthreads();
        if (stq || kTC) {
            const float4* P4 = reinterpret_cast<const float4*>(Pbuf);
            float4* st4 = reinterpret_cast<float4*>(stq);
            for (int i = row0 * V + tid; i < rg1 * 4 * V; i += T) {
                const int n = i / V, v = i - n * V;
                const float4 r = P4[i];
                if (stq) st4[(n * K) * V + v] = r;
                if (kTC) {
                    const uint32_t rl = (uint32_t)(n - row0);
                    uint8_t* tile = aimg + (size_t)(rl >> 7) * 2 * kTcTile;
                    store_image4(tile, tile + kTcTile, sw128_offset(rl & 127u, (uint32_t)(4 * v)), r);
                }
            }
        }
    }
    uint32_t peer_data[3] = {0u, 0u, 0u}, peer_bar[3] = {0u, 0u, 0u};
    if (CL > 1) {
        const uint32_t mine = smem_u32(Pbuf), mybar = smem_u32(&bars[1]);
        int k = 0;
        for (int r = 0; r < CL; ++r)
            if (r != c) { peer_data[k] = map_to_rank(mine, (uint32_t)r); peer_bar[k] = map_to_rank(mybar, (uint32_t)r); ++k; }
    }
    mbar_wait(&bars[0], 0);
    if (CL > 1) cluster_sync_all();   // every CTA of the cluster runs and has initialised its barriers
    else __syncthreads();

    int ec[ENT > 0 ? ENT : 1];
    float ew[ENT > 0 ? ENT : 1];
    if (ENT > 0) {
        const int i = row0 * V + tid;
#pragma unroll
        for (int k = 0; k < ENT; ++k) { ec[k] = 0; ew[k] = 0.f; }
        if (i < row1 * V) {
            const int n = i / V, v = i - n * V;
            const int2 info = rowinfo_s[n];
#pragma unroll
            for (int k = 0; k < ENT; ++k)
                if (k < info.y) {
                    const int2 pe = ld_one<kCsrSmem>(ent + info.x + k);
                    ec[k] = pe.x * V + v;            // float4 index of the gathered element, fixed for all steps
                    ew[k] = __int_as_float(pe.y);
                }
        }
    }
    const uint32_t rx_bytes = (uint32_t)(N - (row1 - row0)) * (uint32_t)DP * 4u;   // rows the peers send per step
    const int ntiles = (rg1 - rg0) * GG;
    float4 acc[TPT][4];
#pragma unroll
    for (int s = 0; s < TPT; ++s)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[s][r] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto issue_mma = [&](int j) {      // one thread: this step's 3xTF32 MMAs for every 128-row tile, then the commit
        mbar_wait(&bars[3 + (j & 1)], (uint32_t)((j >> 1) & 1));      // weight image of step j has landed
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc_tf32(128u, (uint32_t)p.GP16, 0u, 0u);
        const uint8_t* wst = wimg + (size_t)(j & 1) * wstage;
        const uint64_t dbh = make_desc_kmajor(smem_u32(wst)), dbl = make_desc_kmajor(smem_u32(wst + (size_t)p.GP16 * kRowBytes));
        const int nks = (DP + 7) >> 3;
        for (int mt = 0; mt < p.MT; ++mt) {
            const uint8_t* tile = aimg + (size_t)mt * 2 * kTcTile;
            const uint64_t dah = make_desc_kmajor(smem_u32(tile)), dal = make_desc_kmajor(smem_u32(tile + kTcTile));
            const uint32_t acc_col = tmem_acc + (uint32_t)(mt * p.GP16);
            for (int ks = 0; ks < nks; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);          // 8 fp32 = 32 bytes along K inside the swizzle row
                umma_tf32(acc_col, dal + adv, dbh + adv, idesc, (j | ks) ? 1u : 0u);
                umma_tf32(acc_col, dah + adv, dbl + adv, idesc, 1u);
                umma_tf32(acc_col, dah + adv, dbh + adv, idesc, 1u);
            }
        }
        umma_commit(&bars[5]);
        if (j + 1 < K) {   // next step's weight image into the other stage (its last reader, MMA j-1, has completed)
            mbar_arrive_expect_tx(&bars[3 + ((j + 1) & 1)], wstage);
            bulk_g2s(wimg + (size_t)((j + 1) & 1) * wstage, p.Wtc + (size_t)(j + 1) * wstage, wstage, &bars[3 + ((j + 1) & 1)]);
        }
    };
    if (kTC) {
        fence_proxy_async_smem();          // the image stores above (generic proxy) -> visible to the tensor core
        __syncthreads();
    }

    for (int j = 0; j < K; ++j) {
        float* Pcur = Pbuf + (j & 1) * slab;
        if (j > 0) {
            const int b = j & 1;                                  // peer rows of step j are counted on bars[1 + b]
            if (CL > 1) {
                if (j > 1) mbar_wait(&bars[1 + (b ^ 1)], (uint32_t)(((j - 2) >> 1) & 1));   // peers' rows of P_{j-1}
                if (tid == 0) mbar_arrive_expect_tx(&bars[1 + b], rx_bytes);
            }
            const float4* in4 = reinterpret_cast<const float4*>(Pbuf + ((j - 1) & 1) * slab);
            float4* out4 = reinterpret_cast<float4*>(Pcur);
            float4* st4 = p.stack ? reinterpret_cast<float4*>(p.stack + (int64_t)q * NP * K * DP) + j * V : nullptr;
            const bool cheb = (p.recursion == TGCN_RECURSION_CHEBYSHEV) && j >= 2;
            const uint32_t boff = (uint32_t)(b * slab) * 4u;
            bool img_free = false;
            // kTC: the last warp only issues MMAs, so the issue never delays a warp that also owns rows
            const int TW = kTC ? T - 32 : T;
            for (int i = row0 * V + tid; i < row1 * V && tid < TW; i += TW) {
                const int n = i / V, v = i - n * V;
                float4 r;
                if (ENT > 0) {
                    const int2 info = rowinfo_s[n];
                    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
                    for (int k = 0; k < ENT; k += 2) {
                        if (k < info.y) fma4s(a0, ew[k], in4[ec[k]]);
                        if (k + 1 < info.y) fma4s(a1, ew[k + 1], in4[ec[k + 1]]);
                    }
                    if (info.y > ENT) gather_accum<kCsrSmem>(ent, info, ENT, in4, V, v, a0, a1);
                    r = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
                } else {
                    r = gather_row<kCsrSmem>(ent, rowinfo_s[n], in4, V, v);
                }
                if (cheb) {   // T_j = 2 L~ T_{j-1} - T_{j-2}; T_{j-2} is what the output buffer still holds
                    const float4 o = out4[i];
                    r.x = fmaf(-1.f, o.x, 2.f * r.x); r.y = fmaf(-1.f, o.y, 2.f * r.y);
                    r.z = fmaf(-1.f, o.z, 2.f * r.z); r.w = fmaf(-1.f, o.w, 2.f * r.w);
                }
                out4[i] = r;
                if (CL > 1) {       // replicate the new rows into the peers' copies of the slab
                    const uint32_t off = boff + (uint32_t)i * 16u;
                    st_async_f4(peer_data[0] + off, r, peer_bar[0] + 8u * b);
                    if (CL > 2) {
                        st_async_f4(peer_data[1] + off, r, peer_bar[1] + 8u * b);
                        st_async_f4(peer_data[2] + off, r, peer_bar[2] + 8u * b);
                    }
                }
                if (st4) st4[(n * K) * V + v] = r;
                if (kTC) {
                    if (!img_free) {   // MMAs of step j-1 have drained the image: one poller per warp
                        const unsigned grp = __activemask();          // lanes that arrived together
                        if ((tid & 31) == __ffs(grp) - 1) mbar_wait(&bars[5], (uint32_t)((j - 1) & 1));
                        __syncwarp(grp);
                        img_free = true;
                    }
                    const uint32_t rl = (uint32_t)(n - row0);
                    uint8_t* tile = aimg + (size_t)(rl >> 7) * 2 * kTcTile;
                    store_image4(tile, tile + kTcTile, sw128_offset(rl & 127u, (uint32_t)(4 * v)), r);
                }
            }
            if (kTC) fence_proxy_async_smem();
            __syncthreads();
        }
        if (kTC) {
            if (tid == T - 1) issue_mma(j);
            continue;
        }
        // contraction step: acc[tile] += P_j[4 rows][DP] * W'_j[DP][4 g]
        const float4* P4 = reinterpret_cast<const float4*>(Pcur);
        const float4* W4 = reinterpret_cast<const float4*>(kWSmem ? Wsm + j * wslab : p.Wm + (int64_t)j * wslab);
#pragma unroll
        for (int s = 0; s < TPT; ++s) {
            const int t = tid + s * T;
            if (t >= ntiles) break;
            const int rgl = t / GG, gg = t - rgl * GG;
            const int rg = rg0 + rgl;
            for (int v = 0; v < V; ++v) {
                float4 w0, w1, w2, w3;
                if (kWSmem) {
                    w0 = W4[(4 * v + 0) * GG + gg]; w1 = W4[(4 * v + 1) * GG + gg];
                    w2 = W4[(4 * v + 2) * GG + gg]; w3 = W4[(4 * v + 3) * GG + gg];
                } else {
                    w0 = __ldg(W4 + (4 * v + 0) * GG + gg); w1 = __ldg(W4 + (4 * v + 1) * GG + gg);
                    w2 = __ldg(W4 + (4 * v + 2) * GG + gg); w3 = __ldg(W4 + (4 * v + 3) * GG + gg);
                }
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
    // the peers' rows of the last step must have landed before this CTA may leave (remote stores into a dead CTA)
    if (CL > 1 && K > 1) mbar_wait(&bars[1 + ((K - 1) & 1)], (uint32_t)(((K - 2) >> 1) & 1));

    if (kTC) {
        // ---- epilogue from TMEM: warp w reads lanes [32 (w % 4), +32) of tile w / 4; lane = one output row
        mbar_wait(&bars[5], (uint32_t)((K - 1) & 1));
        tcgen05_fence_after();
        const int warp = tid >> 5, lane = tid & 31;
        const int pp = p.pool_p;
        if (warp < 4 * p.MT) {
            const int mt = warp >> 2;
            const int rl = mt * 128 + (warp & 3) * 32 + lane;
            const int n = row0 + rl;
            const bool live = n < row1;
            for (int cb = 0; cb < p.GP16; cb += 16) {
                float v[16];
                tmem_ld16(tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * p.GP16 + cb), v);
                if (cb >= G) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int g = cb + i;
                    if (live && g < G) {
                        if (p.bias_mode == TGCN_BIAS_PER_VERTEX) v[i] += __ldg(p.bias + (int64_t)n * G + g);
                        else if (p.bias_mode == TGCN_BIAS_PER_FILTER) v[i] += __ldg(p.bias + g);
                    }
                }
                if (p.out && live) {
                    float* dst = p.out + ((int64_t)q * N + n) * G + cb;
                    if ((G & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (cb + i < G) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) if (cb + i < G) dst[i] = v[i];
                    }
                }
                if (p.y) {      // siblings are adjacent lanes: the leader of each group of pp lanes takes the first maximum
                    float best[16];
                    unsigned arg[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0 = v[i];
                        if (p.relu) x0 = (x0 != x0) ? x0 : fmaxf(x0, 0.f);
                        if (drop.scale != 0.f) x0 = drop_apply(x0, (uint64_t)(((int64_t)q * N + n) * G + cb + i), drop);
                        float b0 = x0;
                        int a0 = 0;
                        for (int s2 = 1; s2 < pp; ++s2) {
                            const float xs = __shfl_down_sync(0xffffffffu, x0, s2);
                            take_max_r(xs, s2, b0, a0);
                        }
                        best[i] = b0; arg[i] = (unsigned)a0;
                    }
                    if (live && (lane % pp) == 0) {
                        const int64_t off = ((int64_t)q * (N / pp) + n / pp) * G + cb;
                        if ((G & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                if (cb + i < G) {
                                    *reinterpret_cast<float4*>(p.y + off + i) = make_float4(best[i], best[i + 1], best[i + 2], best[i + 3]);
                                    *reinterpret_cast<uchar4*>(p.idx + off + i) =
                                        make_uchar4((unsigned char)arg[i], (unsigned char)arg[i + 1], (unsigned char)arg[i + 2],
                                                    (unsigned char)arg[i + 3]);
                                }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (cb + i < G) { p.y[off + i] = best[i]; p.idx[off + i] = (uint8_t)arg[i]; }
                        }
                    }
                }
            }
        }
        tcgen05_fence_before();
        __syncthreads();
        if (tid < 32) tmem_dealloc(tmem_acc, tmem_cols_pow2((uint32_t)(p.MT * p.GP16)));
        return;
    }

    // epilogue: bias, optional ReLU + max-pool over the tile's sibling rows
    const int pp = p.pool_p;
#pragma unroll
    for (int s = 0; s < TPT; ++s) {
        const int t = tid + s * T;
        if (t >= ntiles) break;
        const int rgl = t / GG, gg = t - rgl * GG;
        const int rg = rg0 + rgl;
        const int g0 = gg * 4;
        float o[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = rg * 4 + r;
            float b[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < N) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (g0 + cc < G) {
                        if (p.bias_mode == TGCN_BIAS_PER_VERTEX) b[cc] = __ldg(p.bias + (int64_t)n * G + g0 + cc);
                        else if (p.bias_mode == TGCN_BIAS_PER_FILTER) b[cc] = __ldg(p.bias + g0 + cc);
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
                    for (int cc = 0; cc < 4; ++cc) if (g0 + cc < G) dst[cc] = o[r][cc];
            }
        }
        if (p.y) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u * pp >= 4) break;                     // pooled rows produced by this tile: 4 / pp
                const int m = (rg * 4) / pp + u;            // pooled row
                if (m * pp >= N) continue;
                float best[4];
                int arg[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) {
                        if (s2 >= pp) break;
                        const int rsel = (pp == 4) ? s2 : ((u * 2 + s2) & 3);
                        float vv = o[rsel][cc];
                        if (p.relu) vv = (vv != vv) ? vv : fmaxf(vv, 0.f);
                        if (drop.scale != 0.f)
                            vv = drop_apply(vv, (uint64_t)(((int64_t)q * N + rg * 4 + rsel) * G + g0 + cc), drop);
                        if (s2 == 0) { best[cc] = vv; arg[cc] = 0; }
                        else take_max_r(vv, s2, best[cc], arg[cc]);
                    }
                }
                const int64_t off = ((int64_t)q * (N / pp) + m) * G + g0;
                if ((G & 3) == 0) {
                    *reinterpret_cast<float4*>(p.y + off) = make_float4(best[0], best[1], best[2], best[3]);
                    *reinterpret_cast<uchar4*>(p.idx + off) =
                        make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
                } else {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        if (g0 + cc < G) { p.y[off + cc] = best[cc]; p.idx[off + cc] = (uint8_t)arg[cc]; }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct ResBwdParams {
    const int2* rowinfoT; const int2* entriesT;               // packed CSR of L~^T (only read when dx != null)
    const float* dout;     // [Q,N,G] or null
    const float* dy;       // [Q,N/p,G] (fused pool) or null
    const uint8_t* idx;    // [Q,N/p,G]
    const float* y;        // pooled forward output (ReLU mask) or null
    const float* stack;    // [Q][NP][K][DP]
    const float* Wt;       // [K][GP][DP] transposed mixed weights (resident_prep_kernel)
    float* dWpart;         // [Q][K][DP][GP] per-sample gradient in the power basis
    float* dbpart;         // [Q][GP] per-sample column sums of dOut (per-filter bias) or null
    float* dx;             // [Q,N,D] or null
    int N, NP, E, D, DP, V, G, GP, GG, K;
    int recursion, pool_p, relu;
    float drop_scale;      // 1 / (1 - p) of the dropout fused behind the ReLU (1 when off)
    int S, Vh;             // CTAs per sample (gridDim.y) and float4 column groups of dx per CTA
    int R;                 // rows of the saved basis per ring stage
};

struct ResSmemBwd { size_t bars, dOut, red, ring, A, W, csr, rowinfo, total; };

static ResSmemBwd res_bwd_smem(int N, int64_t E, int DP, int Vh, int GP, int K, int R, bool need_dx, bool csr_smem,
                               bool w_smem) {
    ResSmemBwd s{};
    const int NP = round_up4(N);
    size_t o = 0;
    s.bars = o;    o += 64;
    s.dOut = o;    o += sizeof(float) * (size_t)NP * (GP + 4);
    s.red = o;     o += sizeof(float) * (size_t)kBwdThreads;
    s.ring = o;    o += sizeof(float) * (size_t)kRingStages * R * K * DP;
    s.A = o;       o += need_dx ? sizeof(float) * 2 * (size_t)NP * 4 * Vh : 0;
    s.W = o;       o += (need_dx && w_smem) ? sizeof(float) * (size_t)K * 4 * Vh * GP : 0;
    s.csr = o;     o += (need_dx && csr_smem) ? sizeof(int2) * (size_t)E : 0;
    s.rowinfo = o; o += need_dx ? sizeof(int2) * (size_t)round_up2(N) : 0;
    s.total = (o + 15) & ~(size_t)15;
    return s;
}

template <int DWT, bool kCsrSmem, bool kWSmem>
__global__ void __launch_bounds__(kBwdThreads, 1)
resident_bwd_kernel(const ResBwdParams p, const ResSmemBwd lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);     // [0..2] ring stages, [3] dx operands
    float* dOut_s = reinterpret_cast<float*>(smem + lay.dOut);
    float* ring = reinterpret_cast<float*>(smem + lay.ring);
    float* Abuf = reinterpret_cast<float*>(smem + lay.A);
    float* Wsm = reinterpret_cast<float*>(smem + lay.W);
    int2* csr_s = reinterpret_cast<int2*>(smem + lay.csr);
    int2* rowinfo_s = reinterpret_cast<int2*>(smem + lay.rowinfo);

    const int tid = threadIdx.x;
    constexpr int T = kBwdThreads;
    const int q = blockIdx.x, h = blockIdx.y, S = p.S;
    const int N = p.N, NP = p.NP, D = p.D, DP = p.DP, V = p.V, G = p.G, GP = p.GP, GG = p.GG, K = p.K;
    const int GS = GG + 1;                                    // dOut row stride in float4 (odd: 8 consecutive rows
    const int GSf = 4 * GS;                                   // hit 8 different bank groups)
    const bool need_dx = p.dx != nullptr;
    const int R = p.R;
    const int rowf = K * DP;                                  // floats of the saved basis per vertex
    const int nchunks = (N + R - 1) / R;
    const float* stq = p.stack + (int64_t)q * NP * rowf;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue_chunk = [&](int cidx) {   // one thread: rows [cidx*R, ...) of every order, one contiguous run
        const int r0 = cidx * R, rows = min(R, N - r0);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)rowf * 4u;
        uint64_t* bar = &bars[cidx % kRingStages];
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(ring + (size_t)(cidx % kRingStages) * R * rowf, stq + (size_t)r0 * rowf, bytes, bar);
    };
    if (tid == 0) {
        for (int cidx = 0; cidx < kRingStages - 1 && cidx < nchunks; ++cidx) issue_chunk(cidx);
        if (need_dx) {
            const uint32_t b_ri = (uint32_t)sizeof(int2) * (uint32_t)round_up2(N);
            const uint32_t b_csr = kCsrSmem ? (uint32_t)sizeof(int2) * (uint32_t)p.E : 0u;
            mbar_arrive_expect_tx(&bars[3], b_ri + b_csr);
            bulk_g2s(rowinfo_s, p.rowinfoT, b_ri, &bars[3]);
            if (b_csr) bulk_g2s(csr_s, p.entriesT, b_csr, &bars[3]);
        }
    }

    // ---- dOut tile of this sample: plain copy, or the max-pool (+ReLU) gradient routed on the fly
    if (p.dout) {
        const float* src = p.dout + (int64_t)q * N * G;
        if ((G & 3) == 0 && aligned16(p.dout)) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            for (int i = tid; i < NP * GG; i += T) {
                const int n = i / GG, gg = i - n * GG;
                reinterpret_cast<float4*>(dOut_s)[n * GS + gg] = n < N ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int i = tid; i < NP * GP; i += T) {
                const int n = i / GP, g = i - n * GP;
                dOut_s[n * GSf + g] = (n < N && g < G) ? __ldg(src + (int64_t)n * G + g) : 0.f;
            }
        }
    } else {
        const int pp = p.pool_p;
        const int M = N / pp;
        const int64_t base = (int64_t)q * M * G;
        if ((G & 3) == 0 && aligned16(p.dy) && (!p.relu || aligned16(p.y)) && (reinterpret_cast<uintptr_t>(p.idx) & 3u) == 0) {
            for (int i = tid; i < M * GG; i += T) {           // one pooled float4: three independent vector loads
                const int m = i / GG, gg = i - m * GG;
                const int64_t off = base + (int64_t)m * G + 4 * gg;
                const uchar4 a = *reinterpret_cast<const uchar4*>(p.idx + off);
                float4 v = __ldg(reinterpret_cast<const float4*>(p.dy + off));
                if (p.relu) {   // relu(x_argmax) = y: > 0 <=> x_argmax > 0, NaN <=> NaN (gradient passes)
                    const float4 yy = __ldg(reinterpret_cast<const float4*>(p.y + off));
                    const float ds = p.drop_scale;   // y > 0: the source passed the ReLU and was kept by the dropout
                    v.x = (yy.x > 0.f || yy.x != yy.x) ? v.x * ds : 0.f; v.y = (yy.y > 0.f || yy.y != yy.y) ? v.y * ds : 0.f;
                    v.z = (yy.z > 0.f || yy.z != yy.z) ? v.z * ds : 0.f; v.w = (yy.w > 0.f || yy.w != yy.w) ? v.w * ds : 0.f;
                }
                for (int s = 0; s < pp; ++s) {
                    const float4 o = make_float4(a.x == s ? v.x : 0.f, a.y == s ? v.y : 0.f, a.z == s ? v.z : 0.f,
                                                 a.w == s ? v.w : 0.f);
                    reinterpret_cast<float4*>(dOut_s)[(m * pp + s) * GS + gg] = o;
                }
            }
            for (int i = N * GG + tid; i < NP * GG; i += T)
                reinterpret_cast<float4*>(dOut_s)[(i / GG) * GS + (i % GG)] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = tid; i < NP * GP; i += T) {
                const int n = i / GP, g = i - n * GP;
                float v = 0.f;
                if (n < N && g < G) {
                    const int m = n / pp, s = n - m * pp;
                    const int64_t off = base + (int64_t)m * G + g;
                    if ((int)p.idx[off] == s) {
                        v = __ldg(p.dy + off);
                        if (p.relu) {
                            const float yy = __ldg(p.y + off);
                            v = (yy > 0.f || yy != yy) ? v * p.drop_scale : 0.f;
                        }
                    }
                }
                dOut_s[n * GSf + g] = v;
            }
        }
    }
    const int Vh = p.Vh, DPh = 4 * Vh;
    const int wslab = DPh * GP;      // this CTA's slice of one order of Wt: [GP][DPh]
    if (need_dx && kWSmem) {
        // Wt[j][g][h*DPh .. +DPh) -> shared [j][g][DPh]
        const float4* src = reinterpret_cast<const float4*>(p.Wt);
        float4* dst = reinterpret_cast<float4*>(Wsm);
        for (int i = tid; i < K * GP * Vh; i += T) {
            const int jg = i / Vh, vl = i - jg * Vh;
            dst[i] = __ldg(src + (int64_t)jg * V + h * Vh + vl);
        }
    }
    __syncthreads();

    // ---- per-filter bias gradient of this sample: column sums of dOut in a fixed order
    float* red_s = reinterpret_cast<float*>(smem + lay.red);
    const int RL = T / GP;                                   // row lanes per column
    const bool do_db = p.dbpart != nullptr && h == 0;
    if (do_db) {
        const int g = tid % GP, lr = tid / GP;
        float s = 0.f;
        if (lr < RL)
            for (int n = lr; n < N; n += RL) s += dOut_s[n * GSf + g];
        red_s[tid] = s;
    }

    // ---- dW'_j[d][g] = sum_n P_j[n][d] dOut[n][g]: DWT (j, 4 d, 4 g) tiles per thread, accumulated over
    // the row chunks as they arrive in the ring; the S CTAs of a sample take alternate warp-sized tile groups
    {
        const float4* dO4 = reinterpret_cast<const float4*>(dOut_s);
        const int per_j = V * GG;
        const int ntiles = K * per_j;
        int tj[DWT], tv[DWT], tg[DWT];
        bool live[DWT];
        float4 acc[DWT][4];
#pragma unroll
        for (int u = 0; u < DWT; ++u) {
            const int te = tid + u * T;
            const int t = (((te >> 5) * S + h) << 5) + (te & 31);
            live[u] = t < ntiles;
            const int tt = live[u] ? t : 0;
            tj[u] = tt / per_j;
            const int rem = tt - tj[u] * per_j;
            tv[u] = rem / GG; tg[u] = rem - tv[u] * GG;
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int cidx = 0; cidx < nchunks; ++cidx) {
            if (tid == 0 && cidx + kRingStages - 1 < nchunks) issue_chunk(cidx + kRingStages - 1);
            mbar_wait(&bars[cidx % kRingStages], (uint32_t)((cidx / kRingStages) & 1));
            const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)(cidx % kRingStages) * R * rowf);
            const int r0 = cidx * R, rows = min(R, N - r0);
#pragma unroll
            for (int u = 0; u < DWT; ++u) {
                if (!live[u]) continue;
                const float4* pcol = st4 + tj[u] * V + tv[u];
                const float4* dcol = dO4 + (size_t)r0 * GS + tg[u];
                int r = 0;
                for (; r + 4 <= rows; r += 4) {
                    float4 pv[4], dv[4];
#pragma unroll
                    for (int w = 0; w < 4; ++w) { pv[w] = pcol[(r + w) * K * V]; dv[w] = dcol[(r + w) * GS]; }
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        fma4s(acc[u][0], pv[w].x, dv[w]);
                        fma4s(acc[u][1], pv[w].y, dv[w]);
                        fma4s(acc[u][2], pv[w].z, dv[w]);
                        fma4s(acc[u][3], pv[w].w, dv[w]);
                    }
                }
                for (; r < rows; ++r) {
                    const float4 pv = pcol[r * K * V], dv = dcol[r * GS];
                    fma4s(acc[u][0], pv.x, dv); fma4s(acc[u][1], pv.y, dv);
                    fma4s(acc[u][2], pv.z, dv); fma4s(acc[u][3], pv.w, dv);
                }
            }
            __syncthreads();          // the stage may be refilled by the next iteration's bulk copy
        }
#pragma unroll
        for (int u = 0; u < DWT; ++u) {
            if (!live[u]) continue;
            float4* dst = reinterpret_cast<float4*>(p.dWpart + (((int64_t)q * K + tj[u]) * DP + 4 * tv[u]) * GP) + tg[u];
            dst[0] = acc[u][0]; dst[GG] = acc[u][1]; dst[2 * GG] = acc[u][2]; dst[3 * GG] = acc[u][3];
        }
    }
    if (do_db) {   // red_s was written before the chunk loop's barriers
        if (nchunks == 0) __syncthreads();
        if (tid < GP) {
            float s = 0.f;
            for (int u = 0; u < RL; ++u) s += red_s[u * GP + tid];
            p.dbpart[(int64_t)q * GP + tid] = s;
        }
    }
    if (!need_dx || h * Vh >= V) return;

    // ---- dx = sum_j (L~^T)^j (dOut W'_j^T): Horner, A_j = dOut W'_j^T + L~^T A_{j+1}; textbook
    // recursion: Clenshaw, B_j = dOut W_j^T + 2 L~^T B_{j+1} - B_{j+2}, dx = dOut W_0^T + L~^T B_1 - B_2.
    // This CTA owns the float4 column groups [h*Vh, (h+1)*Vh) of dx: columns never mix.
    mbar_wait(&bars[3], 0);
    const int2* ent = kCsrSmem ? csr_s : p.entriesT;
    const bool cheb = p.recursion == TGCN_RECURSION_CHEBYSHEV;
    const int aslab = NP * DPh;
    const int nitems = N * Vh;
    // The dense terms G_j = dOut W'_j^T of ALL orders first, with 4-row register tiles (0.5 shared loads per FMA
    // instead of 1.25 when each Horner step computed its own term), into the basis ring, which is free now.
    const bool use_g = (size_t)K * NP * DPh <= (size_t)kRingStages * R * rowf;
    float4* G4 = reinterpret_cast<float4*>(ring);
    if (use_g) {
        const float4* dO4 = reinterpret_cast<const float4*>(dOut_s);
        const int RGn = NP / 4;                               // tile rows: rg, rg + RGn, rg + 2 RGn, rg + 3 RGn
        for (int t = tid; t < K * RGn * Vh; t += T) {
            const int j = t / (RGn * Vh), rem = t - j * (RGn * Vh);
            const int rg = rem / Vh, vl = rem - rg * Vh;
            float4 acc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* wsm = reinterpret_cast<const float4*>(Wsm + j * wslab) + vl;
            const float4* wgl = reinterpret_cast<const float4*>(p.Wt + (int64_t)j * GP * DP) + h * Vh + vl;
            for (int gg = 0; gg < GG; ++gg) {
                float4 w0, w1, w2, w3;
                if (kWSmem) {
                    w0 = wsm[(4 * gg + 0) * Vh]; w1 = wsm[(4 * gg + 1) * Vh]; w2 = wsm[(4 * gg + 2) * Vh]; w3 = wsm[(4 * gg + 3) * Vh];
                } else {
                    w0 = __ldg(wgl + (4 * gg + 0) * V); w1 = __ldg(wgl + (4 * gg + 1) * V);
                    w2 = __ldg(wgl + (4 * gg + 2) * V); w3 = __ldg(wgl + (4 * gg + 3) * V);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 b = dO4[(rg + r * RGn) * GS + gg];
                    fma4s(acc[r], b.x, w0);
                    fma4s(acc[r], b.y, w1);
                    fma4s(acc[r], b.z, w2);
                    fma4s(acc[r], b.w, w3);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) G4[((size_t)j * NP + rg + r * RGn) * Vh + vl] = acc[r];
        }
        __syncthreads();
    }
    for (int j = K - 1; j >= 0; --j) {
        const float4* dO4 = reinterpret_cast<const float4*>(dOut_s);
        // A_{j+1} lives in buffer (j+1)&1; A_j goes to buffer j&1 (which still holds A_{j+2})
        const float4* Ain4 = reinterpret_cast<const float4*>(Abuf + ((j + 1) & 1) * aslab);
        float4* Aout4 = reinterpret_cast<float4*>(Abuf + (j & 1) * aslab);
        for (int i = tid; i < nitems; i += T) {
            const int n = i / Vh, vl = i - n * Vh;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* drow = dO4 + n * GS;
            if (use_g) {
                acc = G4[((size_t)j * NP + n) * Vh + vl];
            } else if (kWSmem) {
                const float4* wcol = reinterpret_cast<const float4*>(Wsm + j * wslab) + vl;
#pragma unroll 4
                for (int gg = 0; gg < GG; ++gg) {
                    const float4 b = drow[gg];
                    fma4s(acc, b.x, wcol[(4 * gg + 0) * Vh]);
                    fma4s(acc, b.y, wcol[(4 * gg + 1) * Vh]);
                    fma4s(acc, b.z, wcol[(4 * gg + 2) * Vh]);
                    fma4s(acc, b.w, wcol[(4 * gg + 3) * Vh]);
                }
            } else {
                const float4* wcol = reinterpret_cast<const float4*>(p.Wt + (int64_t)j * GP * DP) + h * Vh + vl;
#pragma unroll 4
                for (int gg = 0; gg < GG; ++gg) {
                    const float4 b = drow[gg];
                    fma4s(acc, b.x, __ldg(wcol + (4 * gg + 0) * V));
                    fma4s(acc, b.y, __ldg(wcol + (4 * gg + 1) * V));
                    fma4s(acc, b.z, __ldg(wcol + (4 * gg + 2) * V));
                    fma4s(acc, b.w, __ldg(wcol + (4 * gg + 3) * V));
                }
            }
            if (j < K - 1) {
                const float4 s = gather_row<kCsrSmem>(ent, rowinfo_s[n], Ain4, Vh, vl);
                const float sc = (cheb && j >= 1) ? 2.f : 1.f;
                acc.x = fmaf(sc, s.x, acc.x); acc.y = fmaf(sc, s.y, acc.y);
                acc.z = fmaf(sc, s.z, acc.z); acc.w = fmaf(sc, s.w, acc.w);
                if (cheb && j < K - 2) {
                    const float4 o = Aout4[i];
                    acc.x -= o.x; acc.y -= o.y; acc.z -= o.z; acc.w -= o.w;
                }
            }
            if (j > 0) {
                Aout4[i] = acc;
            } else {
                const int d0 = 4 * (h * Vh + vl);
                float* dst = p.dx + ((int64_t)q * N + n) * D + d0;
                if ((D & 3) == 0) *reinterpret_cast<float4*>(dst) = acc;
                else {
                    const float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) if (d0 + cc < D) dst[cc] = o[cc];
                }
            }
        }
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
    float drop_scale;
    int w_blocks;          // blocks [0, w_blocks) reduce weights, the rest the bias
};

constexpr int kRedElems = 32;    // (d,g) elements per weight block
constexpr int kRedMaxK = 32;

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
                // 32 independent loads in flight per round trip; fixed summation tree => deterministic
                float f[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) f[u] = 0.f;
                int q = 0;
                for (; q + 32 <= p.Q; q += 32) {
#pragma unroll
                    for (int u = 0; u < 32; ++u) f[u] += __ldg(src + (q + u) * qs);
                }
#pragma unroll
                for (int u = 0; u < 32; ++u)
                    if (q + u < p.Q) f[u] += __ldg(src + (q + u) * qs);
#pragma unroll
                for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
                    for (int u = 0; u < w; ++u) f[u] += f[u + w];
                s = (double)f[0];
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
        if (p.dout) {
            if (i >= (int64_t)p.N * p.G) return;
            const int64_t NG = (int64_t)p.N * p.G;
            float f[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) f[u] = 0.f;
            int q = 0;
            for (; q + 16 <= p.Q; q += 16) {
#pragma unroll
                for (int u = 0; u < 16; ++u) f[u] += __ldg(p.dout + (q + u) * NG + i);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (q + u < p.Q) f[u] += __ldg(p.dout + (q + u) * NG + i);
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                for (int u = 0; u < w; ++u) f[u] += f[u + w];
            p.db[i] = f[0];
        } else {
            // one pooled element per thread: its gradient goes to the sibling idx selected in each sample
            const int pp = p.pool_p, M = p.N / pp;
            if (i >= (int64_t)M * p.G) return;
            const int64_t MG = (int64_t)M * p.G;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 16
            for (int q = 0; q < p.Q; ++q) {
                const int64_t off = q * MG + i;
                const int a = p.idx[off];
                float v = __ldg(p.dy + off);
                if (p.relu) {
                    const float yy = __ldg(p.y + off);
                    v = (yy > 0.f || yy != yy) ? v * p.drop_scale : 0.f;
                }
                s0 += a == 0 ? v : 0.f; s1 += a == 1 ? v : 0.f; s2 += a == 2 ? v : 0.f; s3 += a == 3 ? v : 0.f;
            }
            const int m = (int)(i / p.G), g = (int)(i - (int64_t)m * p.G);
            const float s[4] = {s0, s1, s2, s3};
            for (int u = 0; u < pp; ++u) p.db[((int64_t)m * pp + u) * p.G + g] = s[u];
        }
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
struct ResPlan { bool ok, csr_smem, w_smem, tc; int ent; int NP, DP, V, GP, GG, GP16, MT, threads, tpt, split, rg_per, Vh, R, dwt; size_t smem; };

static bool res_dims_ok(int N, int D, int G, int K, int64_t E) {
    if (N < 1 || D < 1 || G < 1 || K < 1 || K > kResMaxK || E < 0 || E > (int64_t)INT32_MAX / 16) return false;
    return (int64_t)round_up4(N) * round_up4(D) * K <= (1 << 22) && round_up4(G) <= kBwdThreads;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// forward: CL CTAs (one cluster) per sample, rows split in groups of 4
static ResPlan res_plan_fwd(int Q, int N, int D, int G, int K, int64_t E) {
    ResPlan pl{};
    if (!res_dims_ok(N, D, G, K, E)) return pl;
    pl.NP = round_up4(N); pl.DP = round_up4(D); pl.V = pl.DP / 4; pl.GP = round_up4(G); pl.GG = pl.GP / 4;
    const int RG = pl.NP / 4;
    int CL = 1;
    if ((int64_t)Q * 2 <= kNumSMs && RG >= 2) CL = 2;
    if ((int64_t)Q * 4 <= kNumSMs && RG >= 4) CL = 4;
    const int forced = env_int("TGCN_RES_CL", 0);
    if ((forced == 1 || forced == 2 || forced == 4) && RG >= forced) CL = forced;
    pl.split = CL;
    pl.rg_per = (RG + CL - 1) / CL;
    const int ntiles = pl.rg_per * pl.GG;
    pl.GP16 = round_up16(G);
    pl.MT = (pl.rg_per * 4 + 127) / 128;
    // tensor-core contraction: one 32-wide k-block, accumulators of all row tiles in TMEM, images in shared memory
    if (tuning_value(kTuneResTc) != 0 && pl.DP <= 32 && pl.GP16 <= 256 && pl.MT * pl.GP16 <= 512 && pl.MT <= 4) {
        for (int cs = 1; cs >= 0 && !pl.ok; --cs) {
            pl.csr_smem = cs != 0; pl.w_smem = false;
            pl.smem = res_fwd_smem(N, E, pl.DP, pl.GP, K, pl.csr_smem, false, pl.MT, pl.GP16).total;
            pl.ok = pl.smem <= kResSmemLimit;
        }
        if (pl.ok) { pl.tc = true; pl.threads = 1024; pl.tpt = 1; return pl; }
    }
    pl.threads = (ntiles <= 1024 && env_int("TGCN_RES_T", 1024) == 1024) ? 1024 : 512;
    // at most one (row, float4) item and one tile per thread with 768 threads: cache the row's entries in registers
    if (tuning_value(kTuneResEnt) != 0 && pl.rg_per * 4 * pl.V <= 768 && ntiles <= 768) { pl.threads = 768; pl.ent = 16; }
    pl.tpt = (ntiles + pl.threads - 1) / pl.threads;
    if (pl.tpt > 4) return pl;
    for (int opt = 0; opt < 4 && !pl.ok; ++opt) {
        pl.csr_smem = !(opt & 2); pl.w_smem = !(opt & 1);
        pl.smem = res_fwd_smem(N, E, pl.DP, pl.GP, K, pl.csr_smem, pl.w_smem).total;
        pl.ok = pl.smem <= kResSmemLimit;
    }
    return pl;
}

// backward: S independent CTAs per sample (alternate dW tile groups; disjoint dx column groups)
static ResPlan res_plan_bwd(int Q, int N, int D, int G, int K, int64_t E, bool need_dx) {
    ResPlan pl{};
    if (!res_dims_ok(N, D, G, K, E)) return pl;
    pl.NP = round_up4(N); pl.DP = round_up4(D); pl.V = pl.DP / 4; pl.GP = round_up4(G); pl.GG = pl.GP / 4;
    int S = 1;
    for (int c = 2; c <= 4; c *= 2)
        if ((int64_t)Q * c <= kNumSMs && (!need_dx || pl.V % c == 0)) S = c;
    const int forced = env_int("TGCN_RES_S", 0);
    if ((forced == 1 || forced == 2 || forced == 4) && (!need_dx || pl.V % forced == 0)) S = forced;
    pl.split = S;
    pl.Vh = need_dx ? pl.V / S : pl.V;
    const int ntiles = K * pl.V * pl.GG;
    const int groups = (ntiles + 31) / 32;                    // warp-sized tile groups, dealt round-robin to the S CTAs
    const int my_tiles = ((groups + S - 1) / S) * 32;
    pl.dwt = (my_tiles + kBwdThreads - 1) / kBwdThreads;
    if (pl.dwt > 4) return pl;
    const int rowbytes = K * pl.DP * 4;
    int R = 24 * 1024 / rowbytes;                             // ~24 KB per ring stage
    if (R > N) R = N;
    if (R < 1) R = 1;
    while (!pl.ok && R >= 1) {
        pl.R = R;
        for (int opt = 0; opt < 4 && !pl.ok; ++opt) {
            pl.csr_smem = !(opt & 2); pl.w_smem = !(opt & 1);
            pl.smem = res_bwd_smem(N, E, pl.DP, pl.Vh, pl.GP, K, R, need_dx, pl.csr_smem, pl.w_smem).total;
            pl.ok = pl.smem <= kResSmemLimit;
        }
        R = R > 8 ? R / 2 : R - 1;
    }
    return pl;
}

static int bank_classes(int v_per_cta) { return v_per_cta == 1 ? 8 : v_per_cta == 2 ? 4 : v_per_cta == 4 ? 2 : 1; }

template <typename Kern, typename... Args>
static int res_launch(Kern kern, dim3 grid, int threads, int cluster, size_t smem, cudaStream_t st, const char* name,
                      Args... args) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResSmemLimit);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
    return TGCN_OK;
}

}  // namespace tgcn

using namespace tgcn;

// Host-side packing of a CSR triplet for the resident kernels (see the comment at gather_row).
// rowinfo_host: 2*ceil2(N) int32 (start, len; the padding row is zero); entries_host: 2*E int32
// (col, float bits) or NULL for a size query.  `classes` (1, 2, 4 or 8): tgcn_resident_pack_classes(...).
// Returns E (even), or -1 on bad arguments.
extern "C" int64_t tgcn_pack_csr_host(const int32_t* rowptr_host, const int32_t* col_host, const float* val_host, int N,
                                      int classes, int32_t* rowinfo_host, int32_t* entries_host) {
    if (N < 0 || !rowptr_host || (classes != 1 && classes != 2 && classes != 4 && classes != 8)) return -1;
    int64_t E = 0;
    for (int n = 0; n < N; ++n) {
        const int len = rowptr_host[n + 1] - rowptr_host[n];
        if (len < 0) return -1;
        if (rowinfo_host) { rowinfo_host[2 * n] = (int32_t)E; rowinfo_host[2 * n + 1] = len; }
        if (entries_host) {
            // stable counting sort by ((col - n) mod classes): class n mod m first, then the next one, ...
            int64_t o = E;
            for (int cls = 0; cls < classes; ++cls) {
                for (int e = rowptr_host[n]; e < rowptr_host[n + 1]; ++e) {
                    const int key = ((col_host[e] - n) % classes + classes) % classes;
                    if (key != cls) continue;
                    entries_host[2 * o] = col_host[e];
                    union { float f; int32_t i; } u; u.f = val_host[e];
                    entries_host[2 * o + 1] = u.i;
                    ++o;
                }
            }
            if (len & 1) { entries_host[2 * o] = 0; entries_host[2 * o + 1] = 0; }   // alignment slot, never read as an entry
        }
        E += (len + 1) & ~1;
    }
    if (rowinfo_host && (N & 1)) { rowinfo_host[2 * N] = 0; rowinfo_host[2 * N + 1] = 0; }
    return E;
}

// Bank classes the packing should use for a layer with D features per vertex and batch Q:
// backward == 0: the forward kernel (all V = ceil(D/4) float4 of a row in one CTA);
// backward != 0: the dx recursion of the backward kernel (V / S float4 per CTA).
extern "C" int tgcn_resident_pack_classes(int Q, int N, int D, int backward) {
    if (D < 1 || N < 1 || Q < 0) return 1;
    const int V = round_up4(D) / 4;
    if (!backward) return bank_classes(V);
    const ResPlan pl = res_plan_bwd(Q, N, D, 4, 1, 0, true);
    return bank_classes(pl.Vh > 0 ? pl.Vh : V);
}

extern "C" int tgcn_resident_supported(int N, int D, int G, int K, int64_t nnz) {
    const int64_t E = nnz + N;    // upper bound of the packed entry count
    // worst case over the batch-dependent splits: a single CTA per sample holds the most
    return (res_plan_fwd(kNumSMs, N, D, G, K, E).ok && res_plan_bwd(kNumSMs, N, D, G, K, E, true).ok) ? 1 : 0;
}

extern "C" int64_t tgcn_resident_stack_bytes(int Q, int N, int D, int K) {
    if (Q < 0 || N < 0 || D < 1 || K < 1) return 0;
    return (int64_t)sizeof(float) * Q * K * round_up4(N) * round_up4(D);
}

// weight images written by the forward and read by the backward: [Wm | Wt], K*DP*GP floats each
extern "C" int64_t tgcn_resident_weights_bytes(int D, int G, int K) {
    if (D < 1 || G < 1 || K < 1) return 0;
    // [Wm | Wt] fp32 images, then the tensor-core image [K][hi|lo][G^16][128 B] (16-byte aligned)
    return (int64_t)sizeof(float) * 2 * K * round_up4(D) * round_up4(G) + (int64_t)K * 2 * round_up16(G) * 128;
}

extern "C" int64_t tgcn_resident_bwd_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return (int64_t)sizeof(float) * Q * ((int64_t)K * round_up4(D) * round_up4(G) + round_up4(G));
}

extern "C" int tgcn_resident_layer_fwd(const int32_t* rowinfo, const int32_t* entries, int N, int64_t E,
                                       const float* x, const float* W, const float* bias, int bias_mode,
                                       float* out, float* y, uint8_t* idx, int pool_p, int relu,
                                       const tgcn_dropout_t* drop, float* stack,
                                       float* wimages, int Q, int D, int G, int K, int recursion, void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && G >= 1 && K >= 1, "tgcn_resident_layer_fwd: bad sizes");
    const float drop_p = (drop && drop->p > 0.f) ? drop->p : 0.f;
    TGCN_REQUIRE(drop_p < 1.f, "tgcn_resident_layer_fwd: dropout p = %g must be < 1", (double)drop_p);
    TGCN_REQUIRE(drop_p == 0.f || (relu && y), "tgcn_resident_layer_fwd: dropout is fused between the ReLU and the pool only");
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE || recursion == TGCN_RECURSION_CHEBYSHEV,
                 "tgcn_resident_layer_fwd: unknown recursion %d", recursion);
    if (Q == 0 || N == 0) return TGCN_OK;
    TGCN_REQUIRE(rowinfo && x && W && wimages && (entries || E == 0), "tgcn_resident_layer_fwd: null pointer");
    TGCN_REQUIRE((E & 1) == 0 && aligned16(entries) && aligned16(rowinfo) && aligned16(wimages),
                 "tgcn_resident_layer_fwd: packed CSR / weight images must be 16-byte aligned with an even entry count");
    TGCN_REQUIRE(!stack || aligned16(stack), "tgcn_resident_layer_fwd: stack must be 16-byte aligned");
    TGCN_REQUIRE(out || y, "tgcn_resident_layer_fwd: neither out nor pooled output requested");
    TGCN_REQUIRE(bias_mode == TGCN_BIAS_NONE || bias, "tgcn_resident_layer_fwd: bias_mode %d without bias", bias_mode);
    if (y) {
        TGCN_SUPPORTED(pool_p == 2 || pool_p == 4, "tgcn_resident_layer_fwd: pool size %d", pool_p);
        TGCN_REQUIRE(N % pool_p == 0, "tgcn_resident_layer_fwd: vertex count %d not divisible by pool size %d", N, pool_p);
        TGCN_REQUIRE(idx, "tgcn_resident_layer_fwd: pooled output without idx");
    }
    const ResPlan pl = res_plan_fwd(Q, N, D, G, K, E);
    TGCN_SUPPORTED(pl.ok, "tgcn_resident_layer_fwd: N=%d D=%d G=%d K=%d E=%lld does not fit shared memory", N, D, G, K, (long long)E);
    cudaStream_t st = as_stream(stream);
    float* Wm = wimages;
    float* Wt = wimages + (int64_t)K * pl.DP * pl.GP;
    uint8_t* Wtc = reinterpret_cast<uint8_t*>(wimages + (int64_t)2 * K * pl.DP * pl.GP);
    {
        const int rows_d = pl.DP > 32 ? pl.DP : 32;
        resident_prep_kernel<<<(unsigned)ceil_div(rows_d * pl.GP16, 256), 256, 0, st>>>(W, Wm, Wt, pl.tc ? Wtc : nullptr, K, D, G,
                                                                                      pl.DP, pl.GP, pl.GP16, recursion);
    }
    TGCN_LAUNCH_CHECK("resident_prep");
    ResFwdParams p{};
    p.rowinfo = reinterpret_cast<const int2*>(rowinfo); p.entries = reinterpret_cast<const int2*>(entries);
    p.x = x; p.Wm = Wm; p.Wtc = Wtc; p.GP16 = pl.GP16; p.MT = pl.MT; p.bias = bias; p.out = out; p.y = y; p.idx = idx;
    p.stack = stack; p.N = N; p.NP = pl.NP; p.E = (int)E; p.D = D; p.DP = pl.DP; p.V = pl.V; p.G = G; p.GP = pl.GP;
    p.GG = pl.GG; p.K = K; p.bias_mode = bias_mode; p.recursion = recursion; p.pool_p = y ? pool_p : 4; p.relu = relu;
    p.CL = pl.split; p.rg_per = pl.rg_per;
    p.drop_p = drop_p; p.drop_seed = drop ? drop->seed : 0u; p.drop_step = drop ? drop->step : nullptr;
    const ResSmemFwd lay = res_fwd_smem(N, E, pl.DP, pl.GP, K, pl.csr_smem, pl.w_smem, pl.tc ? pl.MT : 0, pl.GP16);
    const dim3 grid((unsigned)(Q * pl.split));
    if (pl.tc) {
        if (pl.csr_smem)
            TGCN_PROPAGATE(res_launch(resident_fwd_kernel<1024, 1, true, false, true>, grid, 1024, pl.split, lay.total, st,
                                      "resident_layer_fwd", p, lay));
        else
            TGCN_PROPAGATE(res_launch(resident_fwd_kernel<1024, 1, false, false, true>, grid, 1024, pl.split, lay.total, st,
                                      "resident_layer_fwd", p, lay));
        TGCN_LAUNCH_CHECK("resident_layer_fwd");
        return TGCN_OK;
    }
#define TGCN_RES_FWD2(TH, TPT, CS, WS) \
    TGCN_PROPAGATE(res_launch(resident_fwd_kernel<TH, TPT, CS, WS, false>, grid, TH, pl.split, lay.total, st, "resident_layer_fwd", p, lay))
#define TGCN_RES_FWD1(TH, TPT)                                                                                       \
    do {                                                                                                             \
        if (pl.csr_smem) { if (pl.w_smem) TGCN_RES_FWD2(TH, TPT, true, true); else TGCN_RES_FWD2(TH, TPT, true, false); }   \
        else             { if (pl.w_smem) TGCN_RES_FWD2(TH, TPT, false, true); else TGCN_RES_FWD2(TH, TPT, false, false); } \
    } while (0)
    if (pl.ent > 0) {
#define TGCN_RES_FWD_E(CS, WS) \
    TGCN_PROPAGATE(res_launch(resident_fwd_kernel<768, 1, CS, WS, false, 16>, grid, 768, pl.split, lay.total, st, "resident_layer_fwd", p, lay))
        if (pl.csr_smem) { if (pl.w_smem) TGCN_RES_FWD_E(true, true); else TGCN_RES_FWD_E(true, false); }
        else             { if (pl.w_smem) TGCN_RES_FWD_E(false, true); else TGCN_RES_FWD_E(false, false); }
#undef TGCN_RES_FWD_E
    } else if (pl.threads == 1024) {
        TGCN_RES_FWD1(1024, 1);
    } else {
        switch (pl.tpt) { case 1: TGCN_RES_FWD1(512, 1); break; case 2: TGCN_RES_FWD1(512, 2); break;
                          case 3: TGCN_RES_FWD1(512, 3); break; default: TGCN_RES_FWD1(512, 4); break; }
    }
#undef TGCN_RES_FWD1
#undef TGCN_RES_FWD2
    TGCN_LAUNCH_CHECK("resident_layer_fwd");
    return TGCN_OK;
}

extern "C" int tgcn_resident_layer_bwd(const int32_t* rowinfoT, const int32_t* entriesT, int N, int64_t E,
                                       const float* dout, const float* dy, const uint8_t* idx, const float* y,
                                       int pool_p, int relu, const tgcn_dropout_t* drop, const float* stack,
                                       const float* wimages,
                                       float* dW, float* db, int bias_mode, float* dx, void* workspace,
                                       int Q, int D, int G, int K, int recursion, void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && G >= 1 && K >= 1, "tgcn_resident_layer_bwd: bad sizes");
    const float drop_p = (drop && drop->p > 0.f) ? drop->p : 0.f;
    TGCN_REQUIRE(drop_p < 1.f, "tgcn_resident_layer_bwd: dropout p = %g must be < 1", (double)drop_p);
    TGCN_REQUIRE(drop_p == 0.f || (relu && dy), "tgcn_resident_layer_bwd: dropout is fused between the ReLU and the pool only");
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
    TGCN_REQUIRE(stack && wimages && workspace, "tgcn_resident_layer_bwd: null pointer");
    TGCN_REQUIRE(aligned16(stack) && aligned16(wimages) && aligned16(workspace), "tgcn_resident_layer_bwd: misaligned buffer");
    TGCN_REQUIRE((dout != nullptr) != (dy != nullptr), "tgcn_resident_layer_bwd: pass exactly one of dout / dy");
    if (dy) {
        TGCN_SUPPORTED(pool_p == 2 || pool_p == 4, "tgcn_resident_layer_bwd: pool size %d", pool_p);
        TGCN_REQUIRE(N % pool_p == 0 && idx, "tgcn_resident_layer_bwd: bad pooled gradient arguments");
        TGCN_REQUIRE(!relu || y, "tgcn_resident_layer_bwd: relu backward needs the pooled forward output");
    }
    TGCN_REQUIRE(bias_mode == TGCN_BIAS_NONE || db, "tgcn_resident_layer_bwd: bias_mode %d without db", bias_mode);
    if (dx) {
        TGCN_REQUIRE(rowinfoT && (entriesT || E == 0), "tgcn_resident_layer_bwd: dx requested without the packed CSR of L^T");
        TGCN_REQUIRE((E & 1) == 0 && aligned16(entriesT) && aligned16(rowinfoT),
                     "tgcn_resident_layer_bwd: packed CSR must be 16-byte aligned with an even entry count");
    }
    const ResPlan pl = res_plan_bwd(Q, N, D, G, K, E, dx != nullptr);
    TGCN_SUPPORTED(pl.ok, "tgcn_resident_layer_bwd: N=%d D=%d G=%d K=%d E=%lld does not fit shared memory", N, D, G, K, (long long)E);
    ResBwdParams p{};
    p.rowinfoT = reinterpret_cast<const int2*>(rowinfoT); p.entriesT = reinterpret_cast<const int2*>(entriesT);
    p.dout = dout; p.dy = dy; p.idx = idx; p.y = y; p.stack = stack;
    p.Wt = wimages + (int64_t)K * pl.DP * pl.GP; p.dWpart = reinterpret_cast<float*>(workspace); p.dx = dx;
    p.dbpart = bias_mode == TGCN_BIAS_PER_FILTER ? p.dWpart + (int64_t)Q * K * pl.DP * pl.GP : nullptr;
    p.N = N; p.NP = pl.NP; p.E = (int)E; p.D = D;
    p.DP = pl.DP; p.V = pl.V; p.G = G; p.GP = pl.GP; p.GG = pl.GG; p.K = K; p.recursion = recursion;
    p.pool_p = dy ? pool_p : 4; p.relu = relu; p.S = pl.split; p.Vh = pl.Vh; p.R = pl.R;
    p.drop_scale = 1.0f / (1.0f - drop_p);
    const ResSmemBwd lay = res_bwd_smem(N, E, pl.DP, pl.Vh, pl.GP, K, pl.R, dx != nullptr, pl.csr_smem, pl.w_smem);
    const dim3 grid((unsigned)Q, (unsigned)pl.split);
#define TGCN_RES_BWD2(DWT, CS, WS) \
    TGCN_PROPAGATE(res_launch(resident_bwd_kernel<DWT, CS, WS>, grid, kBwdThreads, 1, lay.total, st, "resident_layer_bwd", p, lay))
#define TGCN_RES_BWD1(DWT)                                                                                     \
    do {                                                                                                       \
        if (pl.csr_smem) { if (pl.w_smem) TGCN_RES_BWD2(DWT, true, true); else TGCN_RES_BWD2(DWT, true, false); }   \
        else             { if (pl.w_smem) TGCN_RES_BWD2(DWT, false, true); else TGCN_RES_BWD2(DWT, false, false); } \
    } while (0)
    switch (pl.dwt) { case 1: TGCN_RES_BWD1(1); break; case 2: TGCN_RES_BWD1(2); break;
                      case 3: TGCN_RES_BWD1(3); break; default: TGCN_RES_BWD1(4); break; }
#undef TGCN_RES_BWD1
#undef TGCN_RES_BWD2
    TGCN_LAUNCH_CHECK("resident_layer_bwd");

    ResReduceParams r{};
    r.dWpart = p.dWpart; r.dbpart = p.dbpart; r.dW = dW; r.dout = dout; r.dy = dy; r.idx = idx; r.y = y; r.db = db;
    r.Q = Q; r.N = N; r.D = D; r.DP = pl.DP; r.G = G; r.GP = pl.GP; r.K = K; r.recursion = recursion;
    r.bias_mode = bias_mode; r.pool_p = p.pool_p; r.relu = relu; r.drop_scale = p.drop_scale;
    r.w_blocks = (int)ceil_div((int64_t)D * G, kRedElems);
    int b_blocks = 0;
    if (bias_mode == TGCN_BIAS_PER_VERTEX)
        b_blocks = (int)ceil_div(dout ? (int64_t)N * G : (int64_t)(N / p.pool_p) * G, kRedElems * 16);
    else if (bias_mode == TGCN_BIAS_PER_FILTER) b_blocks = (int)ceil_div(G, kRedElems * 16);
    resident_reduce_kernel<<<(unsigned)(r.w_blocks + b_blocks), kRedElems * 16, 0, st>>>(r);
    TGCN_LAUNCH_CHECK("resident_reduce");
    return TGCN_OK;
}
