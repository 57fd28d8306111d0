// K2 / K4 on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with a 3xTF32 split,
// fp32 accumulators in TMEM, read back with tcgen05.ld for the fused epilogues.
//
//   fwd   : out[m, g]      = sum_{j,d} P_j[m, d] * Wmix[j][d][g] + bias     (A streams, one accumulator)
//   bwd_x : G_j[m, d]      = sum_g dOut[m, g] * Wmix[j][d][g]                (A resident, B streams per j)
//   bwd_w : dWmix[(j,d),g] = sum_m P_j[m, d] * dOut[m, g]                    (both operands MN-major, split over m)
//
// m = n*Q + q enumerates (vertex, sample) pairs: rows of the vertex-major slab.  All three are
// memory-bound (arithmetic intensity ~G/2 flop/B << the tf32 ridge), so the structure is: 128
// threads stage fp32 tiles global -> registers -> (hi, lo) swizzled shared-memory tiles, one
// thread issues the MMAs asynchronously, and several CTAs per SM overlap each other's phases.
#include "common.cuh"
#include "tc_common.cuh"

namespace tgcn {
using namespace tc;

constexpr int kTcThreads = 128;
constexpr int kTileM = 128;                 // rows (pairs) per CTA tile = UMMA M
constexpr uint32_t kTileBytes = 128 * 128;  // one [128 x 32 fp32] operand tile

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// ------------------------------------------------------------------------------------------------
// weight images: the exact shared-memory byte image (hi part then lo part) of each B tile, built
// once per call by a tiny kernel so that the GEMM kernels copy them with plain 16-byte moves.
//   fwd  image unit (j, kb): rows n = g (GP rows), cols k = d in [32 kb, 32 kb + 32)   value W[j][d][g]
//   bwdx image unit (j, gb): rows n = d (DP rows), cols k = g in [32 gb, 32 gb + 32)   value W[j][d][g]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_wimg_kernel(const float* __restrict__ W, uint8_t* __restrict__ img, int K, int D, int G, int rowsP, int KB,
                 int rows_are_g) {
    const int64_t per_unit = (int64_t)rowsP * 32;
    const int64_t total = (int64_t)K * KB * per_unit;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / per_unit;
        const int rem = (int)(i - u * per_unit);
        const int n = rem >> 5, c = rem & 31;
        const int j = (int)(u / KB), kb = (int)(u - (int64_t)j * KB);
        float v = 0.f;
        if (rows_are_g) {
            const int d = kb * 32 + c;
            if (d < D && n < G) v = __ldg(W + ((int64_t)j * D + d) * G + n);
        } else {
            const int g = kb * 32 + c;
            if (g < G && n < D) v = __ldg(W + ((int64_t)j * D + n) * G + g);
        }
        const float h = tf32_hi(v);
        uint8_t* base = img + u * 2 * (int64_t)rowsP * kRowBytes;
        const uint32_t off = sw128_offset((uint32_t)n, (uint32_t)c);
        *reinterpret_cast<float*>(base + off) = h;
        *reinterpret_cast<float*>(base + (int64_t)rowsP * kRowBytes + off) = v - h;
    }
}

// ------------------------------------------------------------------------------------------------
// forward contraction
// ------------------------------------------------------------------------------------------------
struct FwdTcParams {
    const float* stack; int64_t S;
    const uint8_t* wimg;
    const float* bias; int bias_mode;
    float* out;
    int M, Q, N, D, G, GP, K, KB;
};

__global__ void __launch_bounds__(kTcThreads)
contract_fwd_tc_kernel(const FwdTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const uint32_t wtile = (uint32_t)p.GP * kRowBytes;
    uint8_t* a_hi[2] = {smem, smem + kTileBytes};
    uint8_t* a_lo[2] = {smem + 2 * kTileBytes, smem + 3 * kTileBytes};
    uint8_t* w_st[2] = {smem + 4 * kTileBytes, smem + 4 * kTileBytes + 2 * wtile};   // [hi | lo] per stage
    __shared__ __align__(8) uint64_t bar_free[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kTileM;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)p.GP);

    if (tid == 0) {
        mbar_init(&bar_free[0], 1);
        mbar_init(&bar_free[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(kTileM, (uint32_t)p.GP, 0, 0);

    const int units = p.K * p.KB;
    float cur[32], nxt[32];

    auto load_unit = [&](int u, float* v) {
        const int j = u / p.KB, kb = u - j * p.KB;
        const int c = kb * 32 + lane;
        const float* base = p.stack + (int64_t)j * p.S + (int64_t)m0 * p.D + c;
        const bool cok = c < p.D;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int r = warp + 4 * e;
            v[e] = (cok && m0 + r < p.M) ? __ldg(base + (int64_t)r * p.D) : 0.f;
        }
    };

    load_unit(0, cur);
    for (int u = 0; u < units; ++u) {
        const int s = u & 1;
        if (u + 1 < units) load_unit(u + 1, nxt);
        if (u >= 2) mbar_wait(&bar_free[s], (uint32_t)(((u >> 1) - 1) & 1));   // MMAs of unit u-2 have drained stage s
#pragma unroll
        for (int e = 0; e < 32; ++e)
            store_split(a_hi[s], a_lo[s], sw128_offset((uint32_t)(warp + 4 * e), (uint32_t)lane), cur[e]);
        {   // weight image of this unit: 2 * GP * 128 bytes, 16-byte moves
            const float4* src = reinterpret_cast<const float4*>(p.wimg + (int64_t)u * 2 * wtile);
            float4* dst = reinterpret_cast<float4*>(w_st[s]);
            for (int i = tid; i < (int)(2 * wtile / 16); i += kTcThreads) dst[i] = __ldg(src + i);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tcgen05_fence_after();
            const int kb = u % p.KB;
            const int cw = min(32, p.D - kb * 32);
            const int nks = (cw + 7) >> 3;
            const uint64_t dah = make_desc_kmajor(smem_u32(a_hi[s])), dal = make_desc_kmajor(smem_u32(a_lo[s]));
            const uint64_t dbh = make_desc_kmajor(smem_u32(w_st[s])), dbl = make_desc_kmajor(smem_u32(w_st[s] + wtile));
            for (int ks = 0; ks < nks; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);   // 8 fp32 = 32 bytes = 2 x 16 B along K inside the swizzle row
                umma_tf32(tmem_acc, dal + adv, dbh + adv, idesc, (u | ks) ? 1u : 0u);
                umma_tf32(tmem_acc, dah + adv, dbl + adv, idesc, 1u);
                umma_tf32(tmem_acc, dah + adv, dbh + adv, idesc, 1u);
            }
            umma_commit(&bar_free[s]);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) cur[e] = nxt[e];
    }
    {   // all MMAs done: the last commit covers every earlier one
        const int ul = units - 1;
        mbar_wait(&bar_free[ul & 1], (uint32_t)((ul >> 1) & 1));
        tcgen05_fence_after();
    }
    // epilogue: thread = one output row
    const int m = m0 + tid;
    const bool live = m < p.M;
    int n = 0, q = 0;
    if (live) { n = m / p.Q; q = m - n * p.Q; }
    float* dst = p.out + ((int64_t)q * p.N + n) * p.G;
    const bool vec = (p.G % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0);
    for (int cb = 0; cb < p.GP; cb += 16) {
        float v[16];
        tmem_ld16(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
        if (!live) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int g = cb + i;
            if (g < p.G) {
                if (p.bias_mode == TGCN_BIAS_PER_VERTEX) v[i] += __ldg(p.bias + (int64_t)n * p.G + g);
                else if (p.bias_mode == TGCN_BIAS_PER_FILTER) v[i] += __ldg(p.bias + g);
            }
        }
        if (vec) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                if (cb + i < p.G) *reinterpret_cast<float4*>(dst + cb + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (cb + i < p.G) dst[cb + i] = v[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_acc, ncols);
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. the basis: gstack[j][m, d] = sum_g dOut[m, g] Wmix[j][d][g]
// ------------------------------------------------------------------------------------------------
struct BwdXTcParams {
    const float* dout;
    const uint8_t* wimg;     // units (j, gb): [hi | lo] x [DP rows x 128 B]
    float* gstack; int64_t S;
    int M, Q, N, D, G, DP, K, GB;
    int NT;                  // d columns handled per CTA (multiple of 16, <= 64); blockIdx.y selects the chunk
};

__global__ void __launch_bounds__(kTcThreads)
contract_bwd_x_tc_kernel(const BwdXTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    // A: GB blocks x (hi, lo); B: 2 stages x GB blocks x (hi, lo) of [NT rows x 128 B]
    uint8_t* a_hi = smem;
    uint8_t* a_lo = smem + (size_t)p.GB * kTileBytes;
    const uint32_t btile = (uint32_t)p.NT * kRowBytes;
    uint8_t* b_st[2];
    b_st[0] = smem + 2 * (size_t)p.GB * kTileBytes;
    b_st[1] = b_st[0] + 2 * (size_t)p.GB * btile;
    __shared__ __align__(8) uint64_t bar_done[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * kTileM;
    const int n0 = blockIdx.y * p.NT;                      // first d column of this CTA
    const uint32_t ncols = tmem_cols_pow2((uint32_t)(2 * p.NT));

    if (tid == 0) {
        mbar_init(&bar_done[0], 1);
        mbar_init(&bar_done[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);

    // stage the dOut tile once: row r <- API row (q*N + n), 32-column blocks over g
    for (int e = 0; e < 32; ++e) {
        const int r = warp + 4 * e;
        const int m = m0 + r;
        const float* src = nullptr;
        if (m < p.M) {
            const int n = m / p.Q, q = m - n * p.Q;
            src = p.dout + ((int64_t)q * p.N + n) * p.G;
        }
        for (int gb = 0; gb < p.GB; ++gb) {
            const int g = gb * 32 + lane;
            const float v = (src && g < p.G) ? __ldg(src + g) : 0.f;
            store_split(a_hi + (size_t)gb * kTileBytes, a_lo + (size_t)gb * kTileBytes,
                        sw128_offset((uint32_t)r, (uint32_t)lane), v);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(kTileM, (uint32_t)p.NT, 0, 0);
    const int m = m0 + tid;
    const bool live = m < p.M;

    auto epilogue = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bar_done[s], (uint32_t)((j >> 1) & 1));
        tcgen05_fence_after();
        float* dst = p.gstack + (int64_t)j * p.S + (int64_t)m * p.D + n0;
        for (int cb = 0; cb < p.NT; cb += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * p.NT + cb), v);
            if (!live) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (n0 + cb + i < p.D) dst[cb + i] = v[i];
        }
        tcgen05_fence_before();
    };

    for (int j = 0; j < p.K; ++j) {
        const int s = j & 1;
        // B[s] and accumulator s were released by epilogue(j-2), executed by every thread in iteration j-1
        for (int gb = 0; gb < p.GB; ++gb) {
            // image rows [n0, n0+NT) of unit (j, gb); hi and lo parts are DP*128 bytes apart in the image
            const uint8_t* unit = p.wimg + ((int64_t)j * p.GB + gb) * 2 * (int64_t)p.DP * kRowBytes;
            for (int part = 0; part < 2; ++part) {
                const float4* src = reinterpret_cast<const float4*>(unit + (int64_t)part * p.DP * kRowBytes + (int64_t)n0 * kRowBytes);
                float4* dstp = reinterpret_cast<float4*>(b_st[s] + ((size_t)gb * 2 + part) * btile);
                for (int i = tid; i < (int)(btile / 16); i += kTcThreads) dstp[i] = __ldg(src + i);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tcgen05_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(s * p.NT);
            for (int gb = 0; gb < p.GB; ++gb) {
                const int cw = min(32, p.G - gb * 32);
                const int nks = (cw + 7) >> 3;
                const uint64_t dah = make_desc_kmajor(smem_u32(a_hi + (size_t)gb * kTileBytes));
                const uint64_t dal = make_desc_kmajor(smem_u32(a_lo + (size_t)gb * kTileBytes));
                const uint64_t dbh = make_desc_kmajor(smem_u32(b_st[s] + ((size_t)gb * 2 + 0) * btile));
                const uint64_t dbl = make_desc_kmajor(smem_u32(b_st[s] + ((size_t)gb * 2 + 1) * btile));
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);
                    umma_tf32(acc, dal + adv, dbh + adv, idesc, (gb | ks) ? 1u : 0u);
                    umma_tf32(acc, dah + adv, dbl + adv, idesc, 1u);
                    umma_tf32(acc, dah + adv, dbh + adv, idesc, 1u);
                }
            }
            umma_commit(&bar_done[s]);
        }
        if (j >= 1) epilogue(j - 1);     // overlaps the MMAs of j
    }
    epilogue(p.K - 1);
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. the (mixed) weights: dW[(j,d), g] = sum_m P_j[m, d] dOut[m, g]
// A^T and dOut are consumed in their natural layouts as MN-major operands (reduction index m = rows).
// ------------------------------------------------------------------------------------------------
constexpr int kBwKT = 16;   // m rows per staged unit (two k-groups of 8)

struct BwdWTcParams {
    const float* stack; int64_t S;
    const float* dout;
    float* partial;          // [P][JD][G]
    int M, Q, N, D, G, GP, K, JD;
    int MT;                  // 128-row output tiles per CTA; blockIdx.y selects the group
    int units_per_cta;       // units (of kBwKT rows) each blockIdx.x walks, contiguous
};

__global__ void __launch_bounds__(kTcThreads)
contract_bwd_w_tc_kernel(const BwdWTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const uint32_t blk = kBwKT * kRowBytes;                 // one 32-wide MN block: KT rows x 128 B = 2 KB
    const uint32_t a_part = (uint32_t)p.MT * 4 * blk;       // MT tiles x 4 blocks
    const uint32_t nb_b = (uint32_t)(p.GP / 32);
    const uint32_t b_part = nb_b * blk;
    const uint32_t stage_bytes = 2 * a_part + 2 * b_part;   // [A hi | A lo | B hi | B lo]
    uint8_t* st[2] = {smem, smem + stage_bytes};
    __shared__ __align__(8) uint64_t bar_free[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)(p.MT * p.GP));
    const int mn0 = blockIdx.y * p.MT * 128;                 // first packed (j,d) index of this CTA
    const int mn_cnt = min(p.MT * 128, p.JD - mn0);          // live packed indices

    if (tid == 0) {
        mbar_init(&bar_free[0], 1);
        mbar_init(&bar_free[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    // zero both stages once: padding rows/columns are never written afterwards
    for (uint32_t i = tid; i < 2 * stage_bytes / 16; i += kTcThreads) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(128, (uint32_t)p.GP, 1, 1);

    const int total_units = (p.M + kBwKT - 1) / kBwKT;
    const int u_begin = blockIdx.x * p.units_per_cta;
    const int u_end = min(total_units, u_begin + p.units_per_cta);
    int it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
        const int s = it & 1;
        const int mbase = u * kBwKT;
        if (it >= 2) mbar_wait(&bar_free[s], (uint32_t)(((it >> 1) - 1) & 1));
        uint8_t* ah = st[s];
        uint8_t* al = st[s] + a_part;
        uint8_t* bh = st[s] + 2 * a_part;
        uint8_t* bl = bh + b_part;
        // A: element (k = m - mbase, mn) = stack[j][m][d], mn - mn0 = packed index inside this CTA
        for (int i = tid; i < kBwKT * mn_cnt; i += kTcThreads) {
            const int k = i / mn_cnt, mnl = i - k * mn_cnt;
            const int mn = mn0 + mnl;
            const int j = mn / p.D, d = mn - j * p.D;
            const int m = mbase + k;
            const float v = (m < p.M) ? __ldg(p.stack + (int64_t)j * p.S + (int64_t)m * p.D + d) : 0.f;
            const uint32_t off = (uint32_t)(mnl >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(mnl & 31));
            store_split(ah, al, off, v);
        }
        // B: element (k, g) = dout[api(m)][g]
        for (int i = tid; i < kBwKT * p.G; i += kTcThreads) {
            const int k = i / p.G, g = i - k * p.G;
            const int m = mbase + k;
            float v = 0.f;
            if (m < p.M) {
                const int n = m / p.Q, q = m - n * p.Q;
                v = __ldg(p.dout + ((int64_t)q * p.N + n) * p.G + g);
            }
            const uint32_t off = (uint32_t)(g >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(g & 31));
            store_split(bh, bl, off, v);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tcgen05_fence_after();
            for (int t = 0; t < p.MT; ++t) {
                if (t * 128 >= mn_cnt) break;
                const uint32_t acc = tmem_base + (uint32_t)(t * p.GP);
                const uint32_t aoff = (uint32_t)t * 4 * blk;
                for (int ks = 0; ks < kBwKT / 8; ++ks) {
                    const uint32_t adv = (uint32_t)ks * kAtomBytes;          // next group of 8 reduction rows
                    const uint64_t dah = make_desc_mnmajor(smem_u32(ah + aoff + adv), blk);
                    const uint64_t dal = make_desc_mnmajor(smem_u32(al + aoff + adv), blk);
                    const uint64_t dbh = make_desc_mnmajor(smem_u32(bh + adv), blk);
                    const uint64_t dbl = make_desc_mnmajor(smem_u32(bl + adv), blk);
                    umma_tf32(acc, dal, dbh, idesc, (it | ks) ? 1u : 0u);
                    umma_tf32(acc, dah, dbl, idesc, 1u);
                    umma_tf32(acc, dah, dbh, idesc, 1u);
                }
            }
            umma_commit(&bar_free[s]);
        }
    }
    float* dst_base = p.partial + (int64_t)blockIdx.x * p.JD * p.G;
    if (it > 0) {
        const int il = it - 1;
        mbar_wait(&bar_free[il & 1], (uint32_t)((il >> 1) & 1));
        tcgen05_fence_after();
        for (int t = 0; t < p.MT; ++t) {
            if (t * 128 >= mn_cnt) break;
            const int mn = mn0 + t * 128 + tid;
            for (int cb = 0; cb < p.GP; cb += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * p.GP + cb), v);
                if (mn < p.JD) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (cb + i < p.G) dst_base[(int64_t)mn * p.G + cb + i] = v[i];
                }
            }
        }
    } else {
        for (int i = tid; i < mn_cnt * p.G; i += kTcThreads) dst_base[(int64_t)mn0 * p.G + i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct TcPlan {
    bool ok;
    int GP, KB, DP, GB, NT, NY;          // fwd / bwd_x
    int GPw, MT, NYw, P, units_per_cta;  // bwd_w
    size_t smem_fwd, smem_bwdx, smem_bwdw;
    int64_t img_fwd_bytes, img_bwdx_bytes;
};

static TcPlan make_plan(int Q, int N, int D, int G, int K) {
    TcPlan t{};
    const int64_t M = (int64_t)Q * N;
    t.GP = round_up(G, 16);
    t.KB = (D + 31) / 32;
    t.DP = round_up(D, 16);
    t.GB = (G + 31) / 32;
    t.NT = t.DP <= 64 ? t.DP : 64;
    t.DP = round_up(t.DP, t.NT);
    t.NY = t.DP / t.NT;
    t.GPw = round_up(G, 32);
    const int JD = K * D;
    const int tiles = (JD + 127) / 128;
    int mt = 256 / t.GPw;                 // <= 256 TMEM columns per CTA so that two CTAs share an SM
    if (mt > 3) mt = 3;                   // keeps the two-stage ring at <= ~100 KB
    if (mt < 1) mt = 1;
    if (mt > tiles) mt = tiles;
    t.MT = mt;
    t.NYw = (tiles + mt - 1) / mt;
    const int64_t total_units = (M + kBwKT - 1) / kBwKT;
    int64_t want = (2 * (int64_t)kNumSMs) / t.NYw;
    if (want < 1) want = 1;
    if (want > total_units) want = total_units > 0 ? total_units : 1;
    t.units_per_cta = (int)((total_units + want - 1) / want);
    if (t.units_per_cta < 1) t.units_per_cta = 1;
    t.P = (int)((total_units + t.units_per_cta - 1) / t.units_per_cta);
    if (t.P < 1) t.P = 1;
    t.smem_fwd = 1024 + 4 * (size_t)kTileBytes + 4 * (size_t)t.GP * kRowBytes;
    t.smem_bwdx = 1024 + 2 * (size_t)t.GB * kTileBytes + 2 * 2 * (size_t)t.GB * t.NT * kRowBytes;
    const size_t blk = kBwKT * kRowBytes;
    t.smem_bwdw = 1024 + 2 * (2 * (size_t)t.MT * 4 * blk + 2 * (size_t)(t.GPw / 32) * blk);
    t.img_fwd_bytes = (int64_t)K * t.KB * 2 * t.GP * kRowBytes;
    t.img_bwdx_bytes = (int64_t)K * t.GB * 2 * t.DP * kRowBytes;
    const size_t lim = 200 * 1024;
    t.ok = G <= 256 && t.GP <= 256 && t.GB <= 4 && t.smem_fwd <= lim && t.smem_bwdx <= lim && t.smem_bwdw <= lim &&
           M < (int64_t)INT32_MAX - 256;
    return t;
}

int tc_supported(int Q, int N, int D, int G, int K) { return make_plan(Q, N, D, G, K).ok ? 1 : 0; }

// scratch the tensor-core engine needs: weight images (fwd + bwd_x) and the bwd_w partials
int64_t tc_fwd_scratch_bytes(int Q, int N, int D, int G, int K) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    return t.img_fwd_bytes + 1024;
}
int64_t tc_bwd_scratch_bytes(int Q, int N, int D, int G, int K) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    const int64_t partial = (int64_t)t.P * K * D * G * (int64_t)sizeof(float);
    return t.img_bwdx_bytes + partial + 2048;
}

template <typename Kern>
static int set_smem(Kern kern, size_t bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu): %s", name, bytes, cudaGetErrorString(e));
    return TGCN_OK;
}

static uint8_t* align_up(void* p, size_t a) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + a - 1) & ~(uintptr_t)(a - 1));
}

int contract_fwd_tc(const float* stack, const float* Wmix, const float* bias, int bias_mode, float* out, void* scratch,
                    int Q, int N, int D, int G, int K, cudaStream_t st) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    TGCN_SUPPORTED(t.ok, "contract_fwd_tc: shape D=%d G=%d outside the tcgen05 tiles", D, G);
    TGCN_REQUIRE(scratch, "contract_fwd_tc: null scratch");
    uint8_t* img = align_up(scratch, 1024);
    {
        const int64_t total = (int64_t)K * t.KB * t.GP * 32;
        prep_wimg_kernel<<<(unsigned)min64(ceil_div(total, 256), 2 * kNumSMs), 256, 0, st>>>(Wmix, img, K, D, G, t.GP, t.KB, 1);
        TGCN_LAUNCH_CHECK("prep_wimg(fwd)");
    }
    FwdTcParams p{};
    p.stack = stack; p.S = (int64_t)N * Q * D; p.wimg = img; p.bias = bias; p.bias_mode = bias_mode; p.out = out;
    p.M = Q * N; p.Q = Q; p.N = N; p.D = D; p.G = G; p.GP = t.GP; p.K = K; p.KB = t.KB;
    TGCN_PROPAGATE(set_smem(contract_fwd_tc_kernel, t.smem_fwd, "contract_fwd_tc"));
    contract_fwd_tc_kernel<<<(unsigned)ceil_div(p.M, kTileM), kTcThreads, t.smem_fwd, st>>>(p);
    TGCN_LAUNCH_CHECK("contract_fwd_tc");
    return TGCN_OK;
}

int contract_bwd_x_tc(const float* dout, const float* Wmix, float* gstack, void* scratch,
                      int Q, int N, int D, int G, int K, cudaStream_t st) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    TGCN_SUPPORTED(t.ok, "contract_bwd_x_tc: shape D=%d G=%d outside the tcgen05 tiles", D, G);
    TGCN_REQUIRE(scratch, "contract_bwd_x_tc: null scratch");
    uint8_t* img = align_up(scratch, 1024);
    {
        const int64_t total = (int64_t)K * t.GB * t.DP * 32;
        prep_wimg_kernel<<<(unsigned)min64(ceil_div(total, 256), 2 * kNumSMs), 256, 0, st>>>(Wmix, img, K, D, G, t.DP, t.GB, 0);
        TGCN_LAUNCH_CHECK("prep_wimg(bwd_x)");
    }
    BwdXTcParams p{};
    p.dout = dout; p.wimg = img; p.gstack = gstack; p.S = (int64_t)N * Q * D;
    p.M = Q * N; p.Q = Q; p.N = N; p.D = D; p.G = G; p.DP = t.DP; p.K = K; p.GB = t.GB; p.NT = t.NT;
    TGCN_PROPAGATE(set_smem(contract_bwd_x_tc_kernel, t.smem_bwdx, "contract_bwd_x_tc"));
    dim3 grid((unsigned)ceil_div(p.M, kTileM), (unsigned)t.NY);
    contract_bwd_x_tc_kernel<<<grid, kTcThreads, t.smem_bwdx, st>>>(p);
    TGCN_LAUNCH_CHECK("contract_bwd_x_tc");
    return TGCN_OK;
}

// partials land in `scratch` after the bwd_x image region; the caller reduces them
int contract_bwd_w_tc(const float* stack, const float* dout, float* partial, int* P_out,
                      int Q, int N, int D, int G, int K, cudaStream_t st) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    TGCN_SUPPORTED(t.ok, "contract_bwd_w_tc: shape D=%d G=%d outside the tcgen05 tiles", D, G);
    BwdWTcParams p{};
    p.stack = stack; p.S = (int64_t)N * Q * D; p.dout = dout; p.partial = partial;
    p.M = Q * N; p.Q = Q; p.N = N; p.D = D; p.G = G; p.GP = t.GPw; p.K = K; p.JD = K * D; p.MT = t.MT;
    p.units_per_cta = t.units_per_cta;
    TGCN_PROPAGATE(set_smem(contract_bwd_w_tc_kernel, t.smem_bwdw, "contract_bwd_w_tc"));
    dim3 grid((unsigned)t.P, (unsigned)t.NYw);
    contract_bwd_w_tc_kernel<<<grid, kTcThreads, t.smem_bwdw, st>>>(p);
    TGCN_LAUNCH_CHECK("contract_bwd_w_tc");
    *P_out = t.P;
    return TGCN_OK;
}

uint8_t* tc_bwd_partial_ptr(void* scratch, int Q, int N, int D, int G, int K) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    return align_up(align_up(scratch, 1024) + t.img_bwdx_bytes, 256);
}

}  // namespace tgcn
