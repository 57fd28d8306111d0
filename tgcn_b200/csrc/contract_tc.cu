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
#include <cstdlib>
#include "tc_common.cuh"

namespace tgcn {
using namespace tc;

constexpr int kTcThreads = 128;
constexpr int kTileM = 128;                 // rows (pairs) per CTA tile = UMMA M
constexpr uint32_t kTileBytes = 128 * 128;  // one [128 x 32 fp32] operand tile

// Align the dynamic shared-memory window to 1024 B with pointer arithmetic on the __shared__ array itself,
// so that the compiler keeps the shared address space (LDS/STS instead of generic LD/ST).
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    const uint32_t a = tc::smem_u32(p);
    return p + (((a + 1023u) & ~1023u) - a);
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
        // the lo part is rounded to tf32 as well: the tensor core would otherwise truncate it (a biased error)
        *reinterpret_cast<float*>(base + (int64_t)rowsP * kRowBytes + off) = tf32_hi(v - h);
    }
}

// ------------------------------------------------------------------------------------------------
// vector helpers: VEC in {1,2,4} fp32 per lane; a row of 32 columns is covered by 32/VEC lanes and
// one warp instruction covers VEC rows.
// ------------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ldg_vec(const float* p, float* v) {
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

// hi/lo split of VEC consecutive fp32 into the two operand tiles (same byte offset in both)
template <int VEC>
__device__ __forceinline__ void store_split_vec(uint8_t* hi, uint8_t* lo, uint32_t off, const float* x) {
    float h[VEC], l[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { h[i] = tf32_hi(x[i]); l[i] = tf32_hi(x[i] - h[i]); }
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(hi + off) = make_float2(h[0], h[1]);
        *reinterpret_cast<float2*>(lo + off) = make_float2(l[0], l[1]);
    } else {
        *reinterpret_cast<float*>(hi + off) = h[0];
        *reinterpret_cast<float*>(lo + off) = l[0];
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

template <int VEC>
__global__ void __launch_bounds__(kTcThreads)
contract_fwd_tc_kernel(const FwdTcParams p) {
    constexpr int LPR = 32 / VEC;   // lanes per 32-column row
    constexpr int NI = 32 / VEC;    // instructions per warp per unit (32 rows, VEC rows each)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const uint32_t wtile = (uint32_t)p.GP * kRowBytes;
    uint8_t* a_hi[2] = {smem, smem + kTileBytes};
    uint8_t* a_lo[2] = {smem + 2 * kTileBytes, smem + 3 * kTileBytes};
    uint8_t* w_st[2] = {smem + 4 * kTileBytes, smem + 4 * kTileBytes + 2 * wtile};   // [hi | lo] per stage
    __shared__ __align__(8) uint64_t bar_free[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane / LPR, c0 = (lane % LPR) * VEC;
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
    float buf[32];

    auto load_unit = [&](int u) {
        const int j = u / p.KB, kb = u - j * p.KB;
        const int c = kb * 32 + c0;
        const float* base = p.stack + (int64_t)j * p.S + (int64_t)m0 * p.D + c;
        const bool cok = c < p.D;                       // D % VEC == 0: vectors are all-or-nothing
#pragma unroll
        for (int e = 0; e < NI; ++e) {
            const int r = warp * 32 + e * VEC + sub;
            if (cok && m0 + r < p.M) {
                ldg_vec<VEC>(base + (int64_t)r * p.D, &buf[e * VEC]);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) buf[e * VEC + i] = 0.f;
            }
        }
    };

    load_unit(0);
    for (int u = 0; u < units; ++u) {
        const int s = u & 1;
        if (u >= 2) mbar_wait(&bar_free[s], (uint32_t)(((u >> 1) - 1) & 1));   // MMAs of unit u-2 have drained stage s
#pragma unroll
        for (int e = 0; e < NI; ++e)
            store_split_vec<VEC>(a_hi[s], a_lo[s], sw128_offset((uint32_t)(warp * 32 + e * VEC + sub), (uint32_t)c0),
                                 &buf[e * VEC]);
        if (u + 1 < units) load_unit(u + 1);            // in flight across the barrier and the MMA issue
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

template <int VEC>
__global__ void __launch_bounds__(kTcThreads)
contract_bwd_x_tc_kernel(const BwdXTcParams p) {
    constexpr int LPR = 32 / VEC;
    constexpr int NI = 32 / VEC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    // A: GB blocks x (hi, lo); B: 2 stages x GB blocks x (hi, lo) of [NT rows x 128 B]; then the
    // epilogue staging area [4 warps][32 rows x NT fp32]
    uint8_t* a_hi = smem;
    uint8_t* a_lo = smem + (size_t)p.GB * kTileBytes;
    const uint32_t btile = (uint32_t)p.NT * kRowBytes;
    uint8_t* b_st[2];
    b_st[0] = smem + 2 * (size_t)p.GB * kTileBytes;
    b_st[1] = b_st[0] + 2 * (size_t)p.GB * btile;
    float* epi = reinterpret_cast<float*>(b_st[1] + 2 * (size_t)p.GB * btile);
    __shared__ __align__(8) uint64_t bar_done[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane / LPR, c0 = (lane % LPR) * VEC;
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
#pragma unroll 4
    for (int e = 0; e < NI; ++e) {
        const int r = warp * 32 + e * VEC + sub;
        const int m = m0 + r;
        const float* src = nullptr;
        if (m < p.M) {
            const int n = m / p.Q, q = m - n * p.Q;
            src = p.dout + ((int64_t)q * p.N + n) * p.G;
        }
        for (int gb = 0; gb < p.GB; ++gb) {
            const int g = gb * 32 + c0;
            float v[VEC];
            if (src && g < p.G) {
                ldg_vec<VEC>(src + g, v);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[i] = 0.f;
            }
            store_split_vec<VEC>(a_hi + (size_t)gb * kTileBytes, a_lo + (size_t)gb * kTileBytes,
                                 sw128_offset((uint32_t)r, (uint32_t)c0), v);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(kTileM, (uint32_t)p.NT, 0, 0);
    const int m = m0 + tid;
    const bool live = m < p.M;
    // When this CTA owns every column of a row (one d-chunk) and D is not a multiple of 4, rows are
    // staged in shared memory so that each warp writes its 32 x D block of the slab as one contiguous,
    // coalesced run; otherwise each thread writes its own row with 16-byte stores.
    const bool linear = (gridDim.y == 1) && (p.D % 4 != 0);
    const int rows_live = min(32, p.M - (m0 + warp * 32));   // may be <= 0
    float* my_epi = epi + (size_t)warp * 32 * p.NT;

    auto epilogue = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bar_done[s], (uint32_t)((j >> 1) & 1));
        tcgen05_fence_after();
        float* dst = p.gstack + (int64_t)j * p.S + (int64_t)m * p.D + n0;
        for (int cb = 0; cb < p.NT; cb += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * p.NT + cb), v);
            if (linear) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (cb + i < p.D) my_epi[lane * p.D + cb + i] = v[i];
            } else if (live) {
                if (p.D % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        if (n0 + cb + i < p.D) *reinterpret_cast<float4*>(dst + cb + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + cb + i < p.D) dst[cb + i] = v[i];
                }
            }
        }
        if (linear) {
            __syncwarp();
            if (rows_live > 0) {
                float* wdst = p.gstack + (int64_t)j * p.S + (int64_t)(m0 + warp * 32) * p.D;
                const int cnt = rows_live * p.D;
                for (int t = lane; t < cnt; t += 32) wdst[t] = my_epi[t];
            }
            __syncwarp();
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
// A CTA owns the orders [j0, j0 + JC); its packed output index is mn = (j - j0) * DPAD + d.
// ------------------------------------------------------------------------------------------------
constexpr int kBwKT = 16;      // m rows per staged unit (two MMA k-steps of 8)
constexpr int kBwMaxFloats = 48;   // staged A floats per thread per unit (JC * CB * 4)

struct BwdWTcParams {
    const float* stack; int64_t S;
    const float* dout;
    float* partial;          // [P][K*D][G]
    int M, Q, N, D, G, GP, K;
    int DPAD, CB, JC, MT;    // padded D (multiple of 8), 32-column blocks per order, orders per CTA, output tiles
    int units_per_cta;       // units (of kBwKT rows) each blockIdx.x walks, contiguous
};

template <int VEC, int VECG>
__global__ void __launch_bounds__(kTcThreads)
contract_bwd_w_tc_kernel(const BwdWTcParams p) {
    constexpr int LPR = 32 / VEC;
    constexpr int RG = kBwKT / VEC;                          // row groups per (order, column block)
    constexpr int NSLOT = kBwMaxFloats / VEC;                // staged vectors per thread
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

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane / LPR, c0 = (lane % LPR) * VEC;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)(p.MT * p.GP));
    const int j0 = blockIdx.y * p.JC;
    const int jc = min(p.JC, p.K - j0);
    const int mn_cnt = jc * p.DPAD;                          // packed indices this CTA produces
    const int n_items = jc * p.CB * RG;                      // (order, column block, row group) triples
    const int GV = p.G / VECG;
    const int n_bitems = kBwKT * GV;

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

    float abuf[kBwMaxFloats];
    float bbuf[8 * VECG];                                    // up to 8 dOut vectors per thread (G <= 256)

    auto load_unit = [&](int u) {
        const int mbase = u * kBwKT;
#pragma unroll
        for (int e = 0; e < NSLOT; ++e) {
            const int item = warp + 4 * e;
            bool ok = item < n_items;
            int jl = 0, cbk = 0, rg = 0;
            if (ok) {
                rg = item % RG;
                const int t = item / RG;
                cbk = t % p.CB;
                jl = t / p.CB;
            }
            const int k = rg * VEC + sub;
            const int d = cbk * 32 + c0;
            ok = ok && d < p.D && mbase + k < p.M;
            if (ok) {
                ldg_vec<VEC>(p.stack + (int64_t)(j0 + jl) * p.S + (int64_t)(mbase + k) * p.D + d, &abuf[e * VEC]);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) abuf[e * VEC + i] = 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int item = tid + kTcThreads * e;
            if (item < n_bitems) {
                const int k = item / GV, gv = item - k * GV;
                const int m = mbase + k;
                if (m < p.M) {
                    const int n = m / p.Q, q = m - n * p.Q;
                    ldg_vec<VECG>(p.dout + ((int64_t)q * p.N + n) * p.G + gv * VECG, &bbuf[e * VECG]);
                } else {
#pragma unroll
                    for (int i = 0; i < VECG; ++i) bbuf[e * VECG + i] = 0.f;
                }
            }
        }
    };

    int it = 0;
    if (u_begin < u_end) load_unit(u_begin);
    for (int u = u_begin; u < u_end; ++u, ++it) {
        const int s = it & 1;
        if (it >= 2) mbar_wait(&bar_free[s], (uint32_t)(((it >> 1) - 1) & 1));
        uint8_t* ah = st[s];
        uint8_t* al = st[s] + a_part;
        uint8_t* bh = st[s] + 2 * a_part;
        uint8_t* bl = bh + b_part;
#pragma unroll
        for (int e = 0; e < NSLOT; ++e) {
            const int item = warp + 4 * e;
            if (item < n_items) {
                const int rg = item % RG;
                const int t = item / RG;
                const int cbk = t % p.CB, jl = t / p.CB;
                const int d = cbk * 32 + c0;
                if (d < p.D) {
                    const int k = rg * VEC + sub;
                    const int mn = jl * p.DPAD + d;
                    const uint32_t off = (uint32_t)(mn >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(mn & 31));
                    store_split_vec<VEC>(ah, al, off, &abuf[e * VEC]);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int item = tid + kTcThreads * e;
            if (item < n_bitems) {
                const int k = item / GV, gv = item - k * GV;
                const int g = gv * VECG;
                const uint32_t off = (uint32_t)(g >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(g & 31));
                store_split_vec<VECG>(bh, bl, off, &bbuf[e * VECG]);
            }
        }
        if (u + 1 < u_end) load_unit(u + 1);
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
    // partial[blockIdx.x][(j, d)][g]
    float* dst_base = p.partial + (int64_t)blockIdx.x * p.K * p.D * p.G;
    if (it > 0) {
        const int il = it - 1;
        mbar_wait(&bar_free[il & 1], (uint32_t)((il >> 1) & 1));
        tcgen05_fence_after();
        for (int t = 0; t < p.MT; ++t) {
            if (t * 128 >= mn_cnt) break;
            const int mn = t * 128 + tid;
            const int jl = mn / p.DPAD, d = mn - jl * p.DPAD;
            const bool ok = mn < mn_cnt && d < p.D;
            float* dst = dst_base + ((int64_t)(j0 + jl) * p.D + d) * p.G;
            for (int cb = 0; cb < p.GP; cb += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * p.GP + cb), v);
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (cb + i < p.G) dst[cb + i] = v[i];
                }
            }
        }
    } else {
        for (int i = tid; i < jc * p.D * p.G; i += kTcThreads) dst_base[(int64_t)j0 * p.D * p.G + i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// second-generation (persistent, TMA-fed) kernels, contract_tc2.cu
int contract_fwd_tc2(const float* stack, const uint8_t* wimg, const float* bias, int bias_mode, float* out,
                     int Q, int N, int D, int G, int GP, int K, cudaStream_t st, int* launched);
int contract_bwd_w_tc2(const float* stack, const float* dout, float* partial, int* P_out,
                       int Q, int N, int D, int G, int K, cudaStream_t st, int* launched);
int bwd_w2_partials(int Q, int N, int D, int G, int K);
// third-generation kernels (TMA tensor copies into the swizzled operand layout), contract_tc3.cu
int contract_fwd_tc3(const float* stack, const uint8_t* wimg, const float* bias, int bias_mode, float* out,
                     int Q, int N, int D, int G, int GP, int K, cudaStream_t st, int* launched);
int contract_bwd_x_tc3(const float* dout, const uint8_t* wimg, float* gstack, int Q, int N, int D, int DP, int G, int K,
                       cudaStream_t st, int* launched);
int contract_bwd_w_tc3(const float* stack, const float* dout, float* partial, int* P_out, int Q, int N, int D, int G, int K,
                       cudaStream_t st, int* launched);
int bwd_w3_partials(int Q, int N, int D, int G, int K);

static bool use_v2() {
    static const bool on = [] { const char* e = getenv("TGCN_TC_V1"); return !(e && e[0] == '1'); }();
    return on;
}

struct TcPlan {
    bool ok;
    int GP, KB, DP, GB, NT, NY;                          // fwd / bwd_x
    int GPw, DPAD, CB, JC, MT, NYw, P, units_per_cta;    // bwd_w
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
    // bwd_w
    t.GPw = round_up(G, 32);
    t.DPAD = round_up(D, 8);
    t.CB = (D + 31) / 32;
    int mt = t.GPw <= 256 ? 256 / t.GPw : 0;             // <= 256 TMEM columns per CTA: two CTAs share an SM
    if (mt > 3) mt = 3;                                  // keeps the two-stage ring near 100 KB
    int jc = mt > 0 ? (mt * 128) / t.DPAD : 0;           // orders whose packed outputs fit the CTA's tiles
    if (jc > kBwMaxFloats / 4 / t.CB) jc = kBwMaxFloats / 4 / t.CB;   // staged floats per thread: JC * CB * 4
    if (jc > K) jc = K;
    t.JC = jc;
    t.MT = jc > 0 ? (jc * t.DPAD + 127) / 128 : 0;
    t.NYw = jc > 0 ? (K + jc - 1) / jc : 0;
    const int64_t total_units = (M + kBwKT - 1) / kBwKT;
    int64_t want = t.NYw > 0 ? (2 * (int64_t)kNumSMs) / t.NYw : 1;
    if (want < 1) want = 1;
    if (want > total_units) want = total_units > 0 ? total_units : 1;
    t.units_per_cta = (int)((total_units + want - 1) / want);
    if (t.units_per_cta < 1) t.units_per_cta = 1;
    t.P = (int)((total_units + t.units_per_cta - 1) / t.units_per_cta);
    if (t.P < 1) t.P = 1;
    t.smem_fwd = 1024 + 4 * (size_t)kTileBytes + 4 * (size_t)t.GP * kRowBytes;
    t.smem_bwdx = 1024 + 2 * (size_t)t.GB * kTileBytes + 2 * 2 * (size_t)t.GB * t.NT * kRowBytes + 128 * (size_t)t.NT * 4;
    const size_t blk = kBwKT * kRowBytes;
    t.smem_bwdw = 1024 + 2 * (2 * (size_t)t.MT * 4 * blk + 2 * (size_t)(t.GPw / 32) * blk);
    t.img_fwd_bytes = (int64_t)K * t.KB * 2 * t.GP * kRowBytes;
    t.img_bwdx_bytes = (int64_t)K * t.GB * 2 * t.DP * kRowBytes;
    const size_t lim = 200 * 1024;
    t.ok = G <= 256 && t.GP <= 256 && t.GB <= 4 && jc >= 1 && t.smem_fwd <= lim && t.smem_bwdx <= lim &&
           t.smem_bwdw <= lim && M < (int64_t)INT32_MAX - 256;
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
    int p2 = bwd_w2_partials(Q, N, D, G, K);
    const int p3 = bwd_w3_partials(Q, N, D, G, K);
    if (p3 > p2) p2 = p3;
    const int64_t partial = (int64_t)(t.P > p2 ? t.P : p2) * K * D * G * (int64_t)sizeof(float);
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

// widest vector (in fp32) with which rows of `ld` floats starting at `base` can be read
static int vec_width(const void* base, int64_t ld, int64_t slab) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base);
    if (ld % 4 == 0 && slab % 4 == 0 && a % 16 == 0) return 4;
    if (ld % 2 == 0 && slab % 2 == 0 && a % 8 == 0) return 2;
    return 1;
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
    {
        int launched = 0;
        TGCN_PROPAGATE(contract_fwd_tc3(stack, img, bias, bias_mode, out, Q, N, D, G, t.GP, K, st, &launched));
        if (launched) return TGCN_OK;
    }
    if (use_v2()) {
        int launched = 0;
        TGCN_PROPAGATE(contract_fwd_tc2(stack, img, bias, bias_mode, out, Q, N, D, G, t.GP, K, st, &launched));
        if (launched) return TGCN_OK;
    }
    const unsigned grid = (unsigned)ceil_div(p.M, kTileM);
    const int vec = vec_width(stack, D, p.S);
    if (vec == 4) {
        TGCN_PROPAGATE(set_smem(contract_fwd_tc_kernel<4>, t.smem_fwd, "contract_fwd_tc"));
        contract_fwd_tc_kernel<4><<<grid, kTcThreads, t.smem_fwd, st>>>(p);
    } else if (vec == 2) {
        TGCN_PROPAGATE(set_smem(contract_fwd_tc_kernel<2>, t.smem_fwd, "contract_fwd_tc"));
        contract_fwd_tc_kernel<2><<<grid, kTcThreads, t.smem_fwd, st>>>(p);
    } else {
        TGCN_PROPAGATE(set_smem(contract_fwd_tc_kernel<1>, t.smem_fwd, "contract_fwd_tc"));
        contract_fwd_tc_kernel<1><<<grid, kTcThreads, t.smem_fwd, st>>>(p);
    }
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
    {
        int launched = 0;
        TGCN_PROPAGATE(contract_bwd_x_tc3(dout, img, gstack, Q, N, D, t.DP, G, K, st, &launched));
        if (launched) return TGCN_OK;
    }
    BwdXTcParams p{};
    p.dout = dout; p.wimg = img; p.gstack = gstack; p.S = (int64_t)N * Q * D;
    p.M = Q * N; p.Q = Q; p.N = N; p.D = D; p.G = G; p.DP = t.DP; p.K = K; p.GB = t.GB; p.NT = t.NT;
    dim3 grid((unsigned)ceil_div(p.M, kTileM), (unsigned)t.NY);
    const int vec = vec_width(dout, G, 4);
    // the 16-byte row stores of the epilogue need 16-byte aligned slab rows
    TGCN_SUPPORTED(D % 4 != 0 || (reinterpret_cast<uintptr_t>(gstack) % 16 == 0 && p.S % 4 == 0),
                   "contract_bwd_x_tc: gstack must be 16-byte aligned");
    if (vec == 4) {
        TGCN_PROPAGATE(set_smem(contract_bwd_x_tc_kernel<4>, t.smem_bwdx, "contract_bwd_x_tc"));
        contract_bwd_x_tc_kernel<4><<<grid, kTcThreads, t.smem_bwdx, st>>>(p);
    } else if (vec == 2) {
        TGCN_PROPAGATE(set_smem(contract_bwd_x_tc_kernel<2>, t.smem_bwdx, "contract_bwd_x_tc"));
        contract_bwd_x_tc_kernel<2><<<grid, kTcThreads, t.smem_bwdx, st>>>(p);
    } else {
        TGCN_PROPAGATE(set_smem(contract_bwd_x_tc_kernel<1>, t.smem_bwdx, "contract_bwd_x_tc"));
        contract_bwd_x_tc_kernel<1><<<grid, kTcThreads, t.smem_bwdx, st>>>(p);
    }
    TGCN_LAUNCH_CHECK("contract_bwd_x_tc");
    return TGCN_OK;
}

template <int VEC>
static int launch_bwd_w(const BwdWTcParams& p, const TcPlan& t, int vecg, cudaStream_t st) {
    dim3 grid((unsigned)t.P, (unsigned)t.NYw);
    if (vecg == 4) {
        TGCN_PROPAGATE(set_smem(contract_bwd_w_tc_kernel<VEC, 4>, t.smem_bwdw, "contract_bwd_w_tc"));
        contract_bwd_w_tc_kernel<VEC, 4><<<grid, kTcThreads, t.smem_bwdw, st>>>(p);
    } else {
        TGCN_PROPAGATE(set_smem(contract_bwd_w_tc_kernel<VEC, 1>, t.smem_bwdw, "contract_bwd_w_tc"));
        contract_bwd_w_tc_kernel<VEC, 1><<<grid, kTcThreads, t.smem_bwdw, st>>>(p);
    }
    return TGCN_OK;
}

// partials land in `scratch` after the bwd_x image region; the caller reduces them
int contract_bwd_w_tc(const float* stack, const float* dout, float* partial, int* P_out,
                      int Q, int N, int D, int G, int K, cudaStream_t st) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    TGCN_SUPPORTED(t.ok, "contract_bwd_w_tc: shape D=%d G=%d outside the tcgen05 tiles", D, G);
    {
        int launched = 0;
        TGCN_PROPAGATE(contract_bwd_w_tc3(stack, dout, partial, P_out, Q, N, D, G, K, st, &launched));
        if (launched) return TGCN_OK;
    }
    if (use_v2()) {
        int launched = 0;
        TGCN_PROPAGATE(contract_bwd_w_tc2(stack, dout, partial, P_out, Q, N, D, G, K, st, &launched));
        if (launched) return TGCN_OK;
    }
    BwdWTcParams p{};
    p.stack = stack; p.S = (int64_t)N * Q * D; p.dout = dout; p.partial = partial;
    p.M = Q * N; p.Q = Q; p.N = N; p.D = D; p.G = G; p.GP = t.GPw; p.K = K;
    p.DPAD = t.DPAD; p.CB = t.CB; p.JC = t.JC; p.MT = t.MT;
    p.units_per_cta = t.units_per_cta;
    const int vec = vec_width(stack, D, p.S);
    int vecg = vec_width(dout, G, 4) == 4 ? 4 : 1;
    if (vecg == 4 && (int64_t)kBwKT * (G / 4) > 8 * kTcThreads) vecg = 1;
    TGCN_SUPPORTED((int64_t)kBwKT * (G / vecg) <= 8 * kTcThreads, "contract_bwd_w_tc: G=%d too wide for the dOut staging registers", G);
    if (vec == 4) TGCN_PROPAGATE(launch_bwd_w<4>(p, t, vecg, st));
    else if (vec == 2) TGCN_PROPAGATE(launch_bwd_w<2>(p, t, vecg, st));
    else TGCN_PROPAGATE(launch_bwd_w<1>(p, t, vecg, st));
    TGCN_LAUNCH_CHECK("contract_bwd_w_tc");
    *P_out = t.P;
    return TGCN_OK;
}

uint8_t* tc_bwd_partial_ptr(void* scratch, int Q, int N, int D, int G, int K) {
    const TcPlan t = make_plan(Q, N, D, G, K);
    return align_up(align_up(scratch, 1024) + t.img_bwdx_bytes, 256);
}

}  // namespace tgcn
