// Third-generation tensor-core contraction kernels: operand tiles arrive by 2-D/3-D TMA tensor copies
// (cp.async.bulk.tensor, hardware SWIZZLE_128B) straight into the layout tcgen05.mma reads.
//
// What the second generation (contract_tc2.cu) left on the table (profiles/r01/time_tc_v2_mesh32k.txt: 1.9-2.2 TB/s
// of the K-slab stream, tensor pipe 7 % active):
//   * every 128 x D fp32 chunk went raw -> shared (bulk copy) -> registers -> TWO swizzled tiles (hi and lo): 5 bytes
//     of shared-memory traffic per operand byte before the tensor core read anything, plus the weight image of the
//     order copied shared -> shared once per unit;
//   * the eight transform warps also ran the epilogue, so the MMA pipe drained at every tile boundary (one TMEM
//     accumulator).
// Here (requires slab rows padded to a multiple of 32 floats -- `Dp`, the layer pads D = 30 -> 32); 17 warps per CTA:
//   warp 0       producer   one elected lane: a TMA tensor copy per (row tile, order, 32-column block) lands a
//                           [128 x 32 fp32] tile in SWIZZLE_128B layout.  kind::tf32 reads the upper 19 bits of each
//                           fp32, so this tile IS the "hi" operand -- no hi pass at all;
//   warps 9-16   transform  lo = rna_tf32(x - trunc_tf32(x)), element-wise at the SAME swizzled offsets (LDS.128 ->
//                           ALU -> STS.128, conflict-free, no address arithmetic), then a proxy fence;
//   warps 1-4    MMA        up to four ISSUER warps, each with its own TMEM accumulators and its own ring of stages
//                           (units are dealt round-robin; a waiter on an mbarrier must see every phase, so rings are
//                           never shared).  One issuer thread spent ~1600 cycles per unit on the descriptor / MMA /
//                           commit instruction stream of twelve N = 32 MMAs while the tensor pipe sat at 13 % -- the
//                           single issuer WAS the bottleneck (profiles/r02/contract_tc3_notes.txt), hence several of
//                           them, and the products hi*Wh and hi*Wl fused into ONE MMA of N = 2*GP against the
//                           side-by-side image [Wh | Wl] (the epilogue adds the two halves), plus lo*Wh: 8 MMAs per
//                           tile instead of 12.  Accumulators alternate between TWO TMEM buffers per issuer, so the
//                           next row tile's MMAs start while the epilogue drains the previous one;
//   warps 5-8    epilogue   tcgen05.ld (one TMEM lane quarter per warp) in 16-column chunks, bias, stores.
// The weight images (K-major SWIZZLE_128B, built once per call by prep_wimg_kernel) stay resident in shared memory
// when they fit (cortical mesh layer 1: 80 KB), else they are streamed with their tile (1-D bulk copy, L2 hits).
// bwd_x is the same pipeline with dOut as the TMA-fed operand (3-D box over [Q][N][G]) and W^T images in per-issuer
// slots; bwd_w feeds BOTH operands by TMA as MN-major tiles (SWIZZLE_128B_ATOM_32B tensor maps = the UMMA
// SWIZZLE_128B_BASE32B layout), deals the (D x G) output tiles to the issuers and writes per-CTA partials that
// reduce_partials_kernel sums in a fixed order.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace tgcn {
using namespace tc;

constexpr int kT3TransformWarps = 8;
constexpr int kT3MaxStages = 6;
constexpr uint32_t kT3Tile = 128 * 128;    // one [128 x 32 fp32] operand tile
constexpr size_t kT3SmemLimit = 224 * 1024;

// ---- TMA ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// fp32 tensor [d2][d1][d0] (d0 innermost, contiguous), box [b2][b1][b0], swizzle mode `sw`; out-of-range rows read as 0
static int make_tmap3(CUtensorMap* out, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return set_error(TGCN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    const cuuint32_t box[3] = {b0, b1, b2};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(TGCN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TGCN_OK;
}

__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

__device__ __forceinline__ uint8_t* align1024_3(uint8_t* p) {
    const uint32_t a = smem_u32(p);
    return p + (((a + 1023u) & ~1023u) - a);
}

// 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// lo part of a float4 of fp32 values under the tensor core's tf32 read (upper 19 bits): x - trunc(x) is exact in fp32
// (13 significant bits); it is then ROUNDED to tf32 here, because the tensor core would truncate it -- truncation errors
// all point toward zero and add up linearly over the 10^4..10^5 terms of a weight gradient, rounding errors do not
// (the whole-model training test caught the difference: tests/test_gpu_model.py).
__device__ __forceinline__ float tf32_lo1(float x) {
    return tf32_hi(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}
__device__ __forceinline__ float4 tf32_lo4(const float4& x) {
    return make_float4(tf32_lo1(x.x), tf32_lo1(x.y), tf32_lo1(x.z), tf32_lo1(x.w));
}

// ------------------------------------------------------------------------------------------------
// forward:  out[m, g] = sum_j sum_d P_j[m, d] W'_j[d, g] + bias
// unit u = (order j, column block db) of a row tile: one [128 x 32] TMA tile of the stack
//
// What the first version of this kernel taught (profiles/r02/contract_tc3_notes.txt): with G = 32 filters every
// tcgen05.mma is tiny (128 x 32 x 8: ~16 tensor-pipe cycles) and ONE issuing thread could not feed them -- its own
// instruction stream (descriptor arithmetic, two barrier polls, 12 MMAs and a commit per unit, ~1600 cycles) was the
// pace of the whole kernel (tensor pipe 13 % busy, DRAM 35 %, no stage-count sensitivity).  Hence:
//   * the hi part is multiplied by [W_hi ; W_lo] in ONE MMA of N = 2 GP (the weight image stores the lo rows right
//     behind the hi rows, so both are one K-major B tile): 2 MMAs per k-step instead of 3, columns [GP, 2 GP) of the
//     accumulator hold hi * W_lo and are added in the epilogue;
//   * NI issuer warps take the units round-robin, each with its own TMEM accumulator (no ordering between issuers is
//     needed; the epilogue adds the NI partial sums in a fixed order);
//   * the hot-loop barrier waits carry no clock reads.
// ------------------------------------------------------------------------------------------------
constexpr int kT3Issuers = 4;              // issuer warps 1..4 (NI <= 4 of them active)
constexpr int kT3EpiWarp0 = 1 + kT3Issuers;
constexpr int kT3TransformWarp0F = kT3EpiWarp0 + 4;
constexpr int kT3ThreadsF = (kT3TransformWarp0F + kT3TransformWarps) * 32;     // 17 warps = 544 threads

// bounded poll without clock reads (hot loops): ~2^28 polls, then trap
__device__ __forceinline__ void mbar_wait_hot(uint64_t* bar, uint32_t parity) {
    uint32_t n = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++n > (1u << 28)) __trap();
    }
}

struct Fwd3Params {
    const uint8_t* wimg;           // per unit: [hi | lo] x [GP rows x 128 B], K-major SWIZZLE_128B image
    const float* bias; int bias_mode;
    float* out;
    int M, Q, N, G, GP, K, DB;     // DB = Dp / 32 column blocks per order
    int ntiles, NS, w_resident;
    int NI;                        // active MMA issuer warps = TMEM accumulators per row-tile buffer
    int NSi;                       // stages owned by each issuer (NS = NI * NSi)
    uint32_t wunit;                // bytes of one weight image: 2 * GP * 128
};

// Stage of unit u of the CTA's ti-th row tile.  Every issuer owns its own NSi stages and barriers: a parity wait is
// only sound for a waiter that sees EVERY phase of its barrier, and an issuer only sees the units it multiplies
// (with one shared ring, an issuer that skips units would wait on a phase two ahead and be released by the
// intermediate one of the same parity).  c = running count of the issuer's units.
struct StagePos { uint32_t s, ph; };
__device__ __forceinline__ StagePos fwd3_stage(int ti, int u, int units, int NI, int NSi) {
    const int i = u % NI;
    const int per_tile = (units - i + NI - 1) / NI;          // units of issuer i in one row tile
    const uint32_t c = (uint32_t)(ti * per_tile + u / NI);
    StagePos r;
    r.s = (uint32_t)(i * NSi) + c % (uint32_t)NSi;
    r.ph = (c / (uint32_t)NSi) & 1u;
    return r;
}

__global__ void __launch_bounds__(kT3ThreadsF, 1)
contract_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const Fwd3Params p) {
    extern __shared__ uint8_t smem_raw3[];
    uint8_t* smem = align1024_3(smem_raw3);
    __shared__ __align__(8) uint64_t full[kT3MaxStages], lo_ready[kT3MaxStages], empty[kT3MaxStages], acc_full[2], acc_empty[2], w_full;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = p.NS, units = p.K * p.DB, NI = p.NI;
    const uint32_t stage_bytes = 2 * kT3Tile + (p.w_resident ? 0u : p.wunit);
    uint8_t* stage0 = smem;
    uint8_t* wres = smem + (size_t)NS * stage_bytes;                  // resident weight images (w_resident)
    const uint32_t accw = (uint32_t)(NI * 2 * p.GP);                  // TMEM columns of one row-tile buffer
    const uint32_t ncols = tmem_cols_pow2(2u * accw);

    if (tid == 0) {
        for (int i = 0; i < kT3MaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&lo_ready[i], kT3TransformWarps); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], (uint32_t)NI); mbar_init(&acc_empty[i], 4); }
        mbar_init(&w_full, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            if (p.w_resident) {
                const uint32_t wtotal = (uint32_t)units * p.wunit;
                mbar_arrive_expect_tx(&w_full, wtotal);
                for (uint32_t off = 0; off < wtotal; off += 32768u)
                    bulk_g2s(wres + off, p.wimg + off, min(32768u, wtotal - off), &w_full);
            }
            int ti = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
                const int m0 = tile * 128;
                for (int u = 0; u < units; ++u) {
                    const StagePos sp = fwd3_stage(ti, u, units, NI, p.NSi);
                    const uint32_t s = sp.s;
                    mbar_wait_hot(&empty[s], sp.ph ^ 1u);            // first use: parity 1 passes on a fresh barrier
                    uint8_t* st = stage0 + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], kT3Tile + (p.w_resident ? 0u : p.wunit));
                    const int j = u / p.DB, db = u - j * p.DB;
                    tma_load_3d(st, &tmA, db * 32, m0, j, &full[s]);
                    if (!p.w_resident) bulk_g2s(st + 2 * kT3Tile, p.wimg + (size_t)u * p.wunit, p.wunit, &full[s]);
                }
            }
        }
    } else if (warp < kT3EpiWarp0) {
        // ===================== MMA issuers: issuer i takes the units u = i, i + NI, ... of every row tile =====================
        const int i = warp - 1;
        if (lane == 0 && i < NI) {
            const uint32_t idesc2 = make_idesc_tf32(128, 2u * (uint32_t)p.GP, 0, 0);     // hi x [W_hi ; W_lo]
            const uint32_t idesc1 = make_idesc_tf32(128, (uint32_t)p.GP, 0, 0);          // lo x W_hi
            if (p.w_resident) mbar_wait(&w_full, 0);
            const uint64_t desc_stage0 = make_desc_kmajor(smem_u32(stage0));
            const uint64_t desc_w0 = make_desc_kmajor(smem_u32(p.w_resident ? wres : stage0 + 2 * kT3Tile));
            const uint64_t stage_step = (uint64_t)(stage_bytes >> 4), w_step = (uint64_t)(p.wunit >> 4);
            uint32_t ti = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
                const uint32_t a = ti & 1u;
                if (ti >= 2) mbar_wait_hot(&acc_empty[a], ((ti >> 1) - 1) & 1);          // the epilogue drained this buffer
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + a * accw + (uint32_t)(i * 2 * p.GP);
                uint32_t fresh = 0u;                                                      // accumulate flag of the next hi MMA
                for (int u = i; u < units; u += NI) {
                    const StagePos sp = fwd3_stage((int)ti, u, units, NI, p.NSi);
                    const uint32_t s = sp.s, ph = sp.ph;
                    mbar_wait_hot(&full[s], ph);
                    mbar_wait_hot(&lo_ready[s], ph);
                    tcgen05_fence_after();
                    const uint64_t dah = desc_stage0 + (uint64_t)s * stage_step;          // raw tile = hi operand
                    const uint64_t dal = dah + (uint64_t)(kT3Tile >> 4);
                    const uint64_t dbw = p.w_resident ? desc_w0 + (uint64_t)u * w_step : desc_w0 + (uint64_t)s * stage_step;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);                          // 32 bytes along K, in 16-byte units
                        umma_tf32(acc, dah + adv, dbw + adv, idesc2, fresh);
                        umma_tf32(acc, dal + adv, dbw + adv, idesc1, 1u);
                        fresh = 1u;
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[a]);
            }
        }
    } else if (warp < kT3TransformWarp0F) {
        // ===================== epilogue (TMEM lane quarter = warp % 4) =====================
        const int lq = warp & 3;
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
            const uint32_t a = ti & 1u;
            mbar_wait(&acc_full[a], (ti >> 1) & 1);
            tcgen05_fence_after();
            const int m = tile * 128 + lq * 32 + lane;
            const bool live = m < p.M;
            int n = 0, q = 0;
            if (live) { n = m / p.Q; q = m - n * p.Q; }
            float* dst = p.out + ((int64_t)q * p.N + n) * p.G;
            const bool vec = (p.G % 4 == 0) && aligned16(p.out);
            for (int cb = 0; cb < p.GP; cb += 16) {                                 // GP is a multiple of 16
                float v[16];
                const uint32_t tcol = tmem_base + ((uint32_t)(lq * 32) << 16) + a * accw + (uint32_t)cb;
                tmem_ld16(tcol, v);
                for (int k = 1; k < 2 * NI; ++k) {                                 // fixed order: issuer 0 (hi*Wh+lo*Wh, hi*Wl), issuer 1, ...
                    float w[16];
                    tmem_ld16(tcol + (uint32_t)(k * p.GP), w);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += w[e];
                }
                if (!live) continue;
                if (p.bias_mode == TGCN_BIAS_PER_VERTEX) {
                    const float* b = p.bias + (int64_t)n * p.G + cb;
                    if (vec) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            if (cb + e < p.G) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(b + e));
                                v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
                            }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) if (cb + e < p.G) v[e] += __ldg(b + e);
                    }
                } else if (p.bias_mode == TGCN_BIAS_PER_FILTER) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) if (cb + e < p.G) v[e] += __ldg(p.bias + cb + e);
                }
                if (vec) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        if (cb + e < p.G) *reinterpret_cast<float4*>(dst + cb + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) if (cb + e < p.G) dst[cb + e] = v[e];
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
    } else {
        // ===================== transform: lo tile from the raw (= hi) tile =====================
        const int t = tid - kT3TransformWarp0F * 32;                  // 0..255
        const uint32_t base_u32 = smem_u32(stage0);
        int ti = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
            for (int u = 0; u < units; ++u) {
                const StagePos sp = fwd3_stage(ti, u, units, NI, p.NSi);
                const uint32_t s = sp.s, ph = sp.ph;
                mbar_wait_hot(&full[s], ph);
                const uint32_t src = base_u32 + s * stage_bytes + (uint32_t)t * 16u;
                float4 x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[e].x), "=f"(x[e].y), "=f"(x[e].z), "=f"(x[e].w)
                                 : "r"(src + (uint32_t)e * 4096u));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 l = tf32_lo4(x[e]);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src + kT3Tile + (uint32_t)e * 4096u), "f"(l.x), "f"(l.y),
                                 "f"(l.z), "f"(l.w) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&lo_ready[s]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. the input:  G_j[m, d] = sum_g dOut[m, g] W'_j[d, g]   for every order j
// The dOut tile of 128 (vertex, sample) pairs is loaded ONCE (3-D TMA box over dOut[Q][N][G]: 128/Q vertices x Q
// samples x 32 filters, rows land sample-major), its lo part is formed once, and the K orders then stream their
// weight images through a small ring; the accumulator of order j drains (TMEM -> 16-byte stores of the slab rows)
// while order j + 1 multiplies.
// ------------------------------------------------------------------------------------------------
constexpr int kX3WRing = 4;

struct BwdX3Params {
    const uint8_t* wimg;           // per (j, gb): [hi | lo] x [DP rows (d) x 128 B (32 g)], K-major SWIZZLE_128B
    float* gstack;                 // [K][M][D]
    int M, Q, N, D, DP, G, GB, K, nt;      // nt = 128 / Q vertices per tile
    int ntiles, abufs, NI;
    uint32_t wunit;                // bytes of one order's images: GB * 2 * DP * 128
};

// Issuer i multiplies the orders j = i, i + NI, ... of every row tile, alternating between its own two TMEM
// accumulators [128 x 2 DP] (hi x [W_hi ; W_lo] in one MMA of N = 2 DP, then lo x W_hi into the first DP columns).
__global__ void __launch_bounds__(kT3ThreadsF, 1)
contract_bwd_x_tc3_kernel(const __grid_constant__ CUtensorMap tmD, const BwdX3Params p) {
    extern __shared__ uint8_t smem_raw3[];
    uint8_t* smem = align1024_3(smem_raw3);
    __shared__ __align__(8) uint64_t a_full[2], a_lo[2], a_empty[2], w_full[kX3WRing], w_empty[kX3WRing];
    __shared__ __align__(8) uint64_t acc_full[kT3Issuers][2], acc_empty[kT3Issuers][2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NI = p.NI;
    const uint32_t abytes = (uint32_t)p.GB * 2u * kT3Tile;            // one A buffer: GB x [raw | lo]
    uint8_t* abase = smem;
    uint8_t* wbase = smem + (size_t)p.abufs * abytes;
    const uint32_t accw = 2u * (uint32_t)p.DP;                        // columns of one accumulator
    const uint32_t ncols = tmem_cols_pow2(2u * (uint32_t)NI * accw);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_lo[i], kT3TransformWarps); mbar_init(&a_empty[i], (uint32_t)NI); }
        for (int i = 0; i < kT3Issuers; ++i)
            for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[i][b], 1); mbar_init(&acc_empty[i][b], 4); }
        for (int i = 0; i < kX3WRing; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tmD);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t AB = (uint32_t)p.abufs;

    if (warp == 0) {
        if (lane == 0) {
            auto load_a = [&](uint32_t ti, int tile) {
                const uint32_t ab = ti % AB;
                if (ti >= AB) mbar_wait_hot(&a_empty[ab], ((ti / AB) - 1) & 1);
                mbar_arrive_expect_tx(&a_full[ab], (uint32_t)p.GB * kT3Tile);
                for (int gb = 0; gb < p.GB; ++gb)
                    tma_load_3d(abase + (size_t)ab * abytes + (size_t)gb * 2 * kT3Tile, &tmD, gb * 32, tile * p.nt, 0, &a_full[ab]);
            };
            uint32_t ti = 0, gw = 0;
            if ((int)blockIdx.x < p.ntiles) load_a(0, blockIdx.x);
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
                if (AB > 1 && tile + (int)gridDim.x < p.ntiles) load_a(ti + 1, tile + gridDim.x);
                for (int j = 0; j < p.K; ++j, ++gw) {
                    // weight-image slots are owned per issuer (see fwd3_stage): slot and phase from the issuer's job count
                    const StagePos wp = fwd3_stage((int)ti, j, p.K, NI, kX3WRing / NI);
                    const uint32_t ws = wp.s;
                    mbar_wait_hot(&w_empty[ws], wp.ph ^ 1u);
                    mbar_arrive_expect_tx(&w_full[ws], p.wunit);
                    bulk_g2s(wbase + (size_t)ws * p.wunit, p.wimg + (size_t)j * p.wunit, p.wunit, &w_full[ws]);
                }
                if (AB == 1 && tile + (int)gridDim.x < p.ntiles) load_a(ti + 1, tile + gridDim.x);
            }
        }
    } else if (warp < kT3EpiWarp0) {
        const int i = warp - 1;
        if (lane == 0 && i < NI) {
            const uint32_t idesc2 = make_idesc_tf32(128, 2u * (uint32_t)p.DP, 0, 0);
            const uint32_t idesc1 = make_idesc_tf32(128, (uint32_t)p.DP, 0, 0);
            const uint64_t whalf = (uint64_t)(((uint32_t)p.DP * kRowBytes) >> 4);     // hi image of one (j, gb); lo follows
            const uint64_t desc_a0 = make_desc_kmajor(smem_u32(abase)), desc_w0 = make_desc_kmajor(smem_u32(wbase));
            uint32_t ti = 0, cnt = 0;                                                 // cnt: jobs this issuer has done
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
                const uint32_t ab = ti % AB, aph = (ti / AB) & 1;
                mbar_wait_hot(&a_full[ab], aph);
                mbar_wait_hot(&a_lo[ab], aph);
                const uint64_t dat = desc_a0 + (uint64_t)ab * (uint64_t)(abytes >> 4);
                for (int j = i; j < p.K; j += NI, ++cnt) {
                    const StagePos wp = fwd3_stage((int)ti, j, p.K, NI, kX3WRing / NI);
                    const uint32_t ws = wp.s, acb = cnt & 1u;
                    mbar_wait_hot(&w_full[ws], wp.ph);
                    if (cnt >= 2) mbar_wait_hot(&acc_empty[i][acb], ((cnt >> 1) - 1) & 1);
                    tcgen05_fence_after();
                    const uint32_t acc = tmem_base + ((uint32_t)i * 2u + acb) * accw;
                    const uint64_t dw = desc_w0 + (uint64_t)ws * (uint64_t)(p.wunit >> 4);
                    uint32_t fresh = 0u;
                    for (int gb = 0; gb < p.GB; ++gb) {
                        const uint64_t dah = dat + (uint64_t)gb * (uint64_t)((2 * kT3Tile) >> 4);
                        const uint64_t dal = dah + (uint64_t)(kT3Tile >> 4);
                        const uint64_t dbw = dw + (uint64_t)gb * 2u * whalf;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t adv = (uint64_t)(ks * 2);
                            umma_tf32(acc, dah + adv, dbw + adv, idesc2, fresh);
                            umma_tf32(acc, dal + adv, dbw + adv, idesc1, 1u);
                            fresh = 1u;
                        }
                    }
                    umma_commit(&w_empty[ws]);
                    umma_commit(&acc_full[i][acb]);
                }
                umma_commit(&a_empty[ab]);
            }
        }
    } else if (warp < kT3TransformWarp0F) {
        const int lq = warp & 3;
        const int r = lq * 32 + lane;                                  // row of the tile: sample-major (q, vertex)
        const int q = r / p.nt, nl = r - q * p.nt;
        const bool vec = (p.D % 4 == 0) && aligned16(p.gstack);
        uint32_t cnt[kT3Issuers] = {0u, 0u, 0u, 0u};
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int n = tile * p.nt + nl;
            const bool live = n < p.N;
            const int64_t m = (int64_t)n * p.Q + q;
            for (int j = 0; j < p.K; ++j) {
                const int i = j % NI;
                uint32_t c = 0;
#pragma unroll
                for (int k = 0; k < kT3Issuers; ++k) if (k == i) { c = cnt[k]; cnt[k] = c + 1; }
                const uint32_t acb = c & 1u;
                mbar_wait(&acc_full[i][acb], (c >> 1) & 1);
                tcgen05_fence_after();
                float* dst = p.gstack + ((int64_t)j * p.M + m) * p.D;
                const uint32_t tcol0 = tmem_base + ((uint32_t)(lq * 32) << 16) + ((uint32_t)i * 2u + acb) * accw;
                for (int cb = 0; cb < p.DP; cb += 16) {
                    float v[16], w[16];
                    tmem_ld16(tcol0 + (uint32_t)cb, v);
                    tmem_ld16(tcol0 + (uint32_t)(p.DP + cb), w);
                    if (!live) continue;
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += w[e];
                    if (vec) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            if (cb + e < p.D) *reinterpret_cast<float4*>(dst + cb + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) if (cb + e < p.D) dst[cb + e] = v[e];
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[i][acb]);
            }
        }
    } else {
        const int t = tid - kT3TransformWarp0F * 32;
        const uint32_t base_u32 = smem_u32(abase);
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
            const uint32_t ab = ti % AB;
            mbar_wait_hot(&a_full[ab], (ti / AB) & 1);
            for (int gb = 0; gb < p.GB; ++gb) {
                const uint32_t src = base_u32 + ab * abytes + (uint32_t)gb * 2u * kT3Tile + (uint32_t)t * 16u;
                float4 x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[e].x), "=f"(x[e].y), "=f"(x[e].z), "=f"(x[e].w)
                                 : "r"(src + (uint32_t)e * 4096u));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 l = tf32_lo4(x[e]);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src + kT3Tile + (uint32_t)e * 4096u), "f"(l.x), "f"(l.y),
                                 "f"(l.z), "f"(l.w) : "memory");
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_lo[ab]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. the mixed weights:  dW'[(j, d), g] = sum_m P_j[m, d] dOut[m, g]
// Both operands are MN-major (the reduction index m is the slow one in memory): a unit of RU pairs is, per order
// and 32-column block, a [RU x 32 fp32] TMA tile in SWIZZLE_128B_ATOM_32B layout -- the layout tcgen05 reads
// MN-major tf32 operands in (SWIZZLE_128B_BASE32B) -- and the dOut rows of the unit arrive through a 3-D box over
// dOut[Q][N][G] whose strides are given vertex-major, so its rows come out in the same (vertex, sample) order.
// Four 32-wide blocks form one 128-row output tile [128 x 2 GPw] (hi^T x [dOut_hi | dOut_lo] in one MMA, lo^T x
// dOut_hi into the first GPw columns); the output tiles are dealt to the issuer warps (tile t -> issuer t mod NI),
// accumulate in TMEM for the whole kernel and are written once, as a per-CTA partial that reduce_partials sums in
// a fixed order.
// ------------------------------------------------------------------------------------------------
struct BwdW3Params {
    float* partial;                // [P][K*D][G]
    int M, Q, N, D, G, GPw, K, DB;
    int NB;                        // 32-wide (order, column block) pairs handled per CTA (grid.y splits the rest)
    int MT;                        // output tiles per CTA = ceil(NB / 4)
    int RU, NS, NI, units_per_cta, total_units;
    int dpg;                       // > 0: blocks in column-block-major order, b = dbi * K + j for the dpg column blocks of this y group:
                                   // the K orders of a column block arrive by ONE tensor copy (box of K orders); 0: order-major
                                   // blocks (gbk = j * DB + db), one copy per block
};

__global__ void __launch_bounds__(kT3ThreadsF, 1)
contract_bwd_w_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD, const BwdW3Params p) {
    extern __shared__ uint8_t smem_raw3[];
    uint8_t* smem = align1024_3(smem_raw3);
    __shared__ __align__(8) uint64_t full[kT3MaxStages], lo_ready[kT3MaxStages], empty[kT3MaxStages], acc_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t blk = (uint32_t)p.RU * kRowBytes;                  // one 32-wide block of RU reduction rows
    const int GBk = p.GPw / 32;
    // hi (= raw) blocks of A.  The last output tile may own fewer than 4 blocks: its MMA then also reads the blocks that
    // follow in the stage (finite or not, they only reach accumulator rows that are never stored), so no padding blocks
    // are allocated.
    const uint32_t a_part = (uint32_t)p.NB * blk;
    const uint32_t b_part = (uint32_t)GBk * blk;
    const uint32_t stage_bytes = 2u * a_part + 2u * b_part;           // [A hi | A lo | B hi | B lo]
    const uint32_t accw = 2u * (uint32_t)p.GPw;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)p.MT * accw);
    const int blk0 = blockIdx.y * p.NB;                               // order-major: first (order, column block) pair of this CTA
    const int db0 = blockIdx.y * p.dpg;                               // column-block-major: first column block of this CTA
    const int nb = p.dpg > 0 ? min(p.dpg, p.DB - db0) * p.K : min(p.NB, p.K * p.DB - blk0);
    // first row of block b in the [K * D] x G output
    auto block_row0 = [&](int b) -> int {
        if (p.dpg > 0) { const int dbi = b / p.K, j = b - dbi * p.K; return j * p.D + (db0 + dbi) * 32; }
        return (blk0 + b) * 32;
    };
    const int u_begin = blockIdx.x * p.units_per_cta;
    const int u_end = min(p.total_units, u_begin + p.units_per_cta);
    const int NS = p.NS, NI = p.NI;

    if (tid == 0) {
        for (int i = 0; i < kT3MaxStages; ++i) { mbar_init(&full[i], 1); mbar_init(&lo_ready[i], kT3TransformWarps); mbar_init(&empty[i], (uint32_t)NI); }
        mbar_init(&acc_full, (uint32_t)NI);
        fence_mbar_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmD);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
    // zero the stages once (+ the slack a last, partly filled output tile reads behind its stage)
    for (uint32_t i = tid; i < ((uint32_t)NS * stage_bytes + 4u * blk) / 16; i += kT3ThreadsF) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int gblocks = (p.G + 31) / 32;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int u = u_begin; u < u_end; ++u, ++g) {
                const uint32_t s = g % (uint32_t)NS;
                if (g >= (uint32_t)NS) mbar_wait_hot(&empty[s], ((g / NS) - 1) & 1);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(nb + gblocks) * blk);
                const int m0 = u * p.RU;
                if (p.dpg > 0) {
                    for (int dbi = 0; dbi * p.K < nb; ++dbi)               // box [32 x RU x K orders] lands as K consecutive blocks
                        tma_load_3d(st + (size_t)dbi * p.K * blk, &tmA, (db0 + dbi) * 32, m0, 0, &full[s]);
                } else {
                    for (int b = 0; b < nb; ++b) {
                        const int gbk = blk0 + b, j = gbk / p.DB, db = gbk - j * p.DB;
                        tma_load_3d(st + (size_t)b * blk, &tmA, db * 32, m0, j, &full[s]);
                    }
                }
                for (int gb = 0; gb < gblocks; ++gb)
                    tma_load_3d(st + 2 * (size_t)a_part + (size_t)gb * blk, &tmD, gb * 32, 0, m0 / p.Q, &full[s]);
            }
        }
    } else if (warp < kT3EpiWarp0) {
        const int i = warp - 1;
        if (lane == 0 && i < NI) {
            const uint32_t idesc2 = make_idesc_tf32(128, 2u * (uint32_t)p.GPw, 1, 1);
            const uint32_t idesc1 = make_idesc_tf32(128, (uint32_t)p.GPw, 1, 1);
            const uint64_t desc0 = make_desc_mnmajor(smem_u32(smem), blk);
            uint32_t g = 0;
            for (int u = u_begin; u < u_end; ++u, ++g) {
                const uint32_t s = g % (uint32_t)NS, ph = (g / NS) & 1;
                mbar_wait_hot(&full[s], ph);
                mbar_wait_hot(&lo_ready[s], ph);
                tcgen05_fence_after();
                const uint64_t dst0 = desc0 + (uint64_t)s * (uint64_t)(stage_bytes >> 4);
                const uint64_t dbh = dst0 + (uint64_t)((2u * a_part) >> 4);              // [B hi blocks | B lo blocks]: N = 2 GPw
                for (int t = i; t < p.MT; t += NI) {
                    const uint32_t acc = tmem_base + (uint32_t)t * accw;
                    const uint64_t dah = dst0 + (uint64_t)(((uint32_t)t * 4u * blk) >> 4);
                    const uint64_t dal = dah + (uint64_t)(a_part >> 4);
                    for (int ks = 0; ks < p.RU / 8; ++ks) {
                        const uint64_t adv = (uint64_t)(((uint32_t)ks * kAtomBytes) >> 4);   // next 8 reduction rows
                        umma_tf32(acc, dah + adv, dbh + adv, idesc2, (g | (uint32_t)ks) ? 1u : 0u);
                        umma_tf32(acc, dal + adv, dbh + adv, idesc1, 1u);
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(&acc_full);
        }
    } else if (warp < kT3TransformWarp0F) {
        // ===================== epilogue: partial[blockIdx.x][(j, d)][g] =====================
        const int lq = warp & 3;
        float* dst_base = p.partial + (int64_t)blockIdx.x * p.K * p.D * p.G;
        if (u_begin < u_end) {
            mbar_wait(&acc_full, 0);
            tcgen05_fence_after();
            for (int t = 0; t < p.MT; ++t) {
                const int b = t * 4 + lq;                                  // block of this lane quarter
                const bool ok = b < nb;
                float* dst = dst_base + ((int64_t)block_row0(ok ? b : 0) + lane) * p.G;
                for (int cb = 0; cb < p.GPw; cb += 16) {
                    float v[16], w[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)t * accw + (uint32_t)cb, v);
                    tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)t * accw + (uint32_t)(p.GPw + cb), w);
                    if (ok) {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (cb + e < p.G) dst[cb + e] = v[e] + w[e];
                    }
                }
            }
        } else {
            const int t0 = tid - kT3EpiWarp0 * 32;
            const int per_blk = 32 * p.G;
            for (int e = t0; e < nb * per_blk; e += 128) {
                const int b = e / per_blk;
                dst_base[(int64_t)block_row0(b) * p.G + (e - b * per_blk)] = 0.f;
            }
        }
    } else {
        // ===================== transform: lo parts of the A blocks and of the dOut blocks =====================
        const int t = tid - kT3TransformWarp0F * 32;
        const uint32_t base_u32 = smem_u32(smem);
        const uint32_t a_f4 = (uint32_t)nb * blk / 16u, b_f4 = (uint32_t)gblocks * blk / 16u;
        uint32_t g = 0;
        for (int u = u_begin; u < u_end; ++u, ++g) {
            const uint32_t s = g % (uint32_t)NS;
            mbar_wait_hot(&full[s], (g / NS) & 1);
            const uint32_t st = base_u32 + s * stage_bytes;
            for (uint32_t i0 = 0; i0 < a_f4 + b_f4; i0 += 1024u) {
                float4 x[4];
                uint32_t off[4], dlt[4];
                bool ok[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t e = i0 + (uint32_t)k * 256u + (uint32_t)t;
                    ok[k] = e < a_f4 + b_f4;
                    const bool isb = e >= a_f4;
                    off[k] = isb ? 2u * a_part + (e - a_f4) * 16u : e * 16u;
                    dlt[k] = isb ? b_part : a_part;
                    if (ok[k])
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[k].x), "=f"(x[k].y), "=f"(x[k].z), "=f"(x[k].w)
                                     : "r"(st + off[k]));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!ok[k]) continue;
                    const float4 l = tf32_lo4(x[k]);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st + off[k] + dlt[k]), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w)
                                 : "memory");
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&lo_ready[s]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

template <typename Kern>
static int set_smem3(Kern kern, size_t bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu): %s", name, bytes, cudaGetErrorString(e));
    return TGCN_OK;
}

static bool use_v3() {
    static const bool on = [] { const char* e = getenv("TGCN_TC_V3"); return !(e && e[0] == '0'); }();
    return on;
}

// returns with *launched = 1 when the shape is covered (slab rows of D = 32 * DB floats, 16-byte aligned stack)
int contract_fwd_tc3(const float* stack, const uint8_t* wimg, const float* bias, int bias_mode, float* out,
                     int Q, int N, int D, int G, int GP, int K, cudaStream_t st, int* launched) {
    *launched = 0;
    const int64_t M = (int64_t)Q * N;
    if (!use_v3() || D % 32 != 0 || GP > 128 || GP % 16 != 0 || M <= 0 || M >= (int64_t)INT32_MAX - 256 ||
        (reinterpret_cast<uintptr_t>(stack) & 15u) != 0)
        return TGCN_OK;
    Fwd3Params p{};
    p.wimg = wimg; p.bias = bias; p.bias_mode = bias_mode; p.out = out;
    p.M = (int)M; p.Q = Q; p.N = N; p.G = G; p.GP = GP; p.K = K; p.DB = D / 32;
    p.ntiles = (int)ceil_div(M, 128);
    p.wunit = 2u * (uint32_t)GP * kRowBytes;
    const size_t wtotal = (size_t)K * p.DB * p.wunit;
    const size_t fixed = 1024 + 256;
    p.w_resident = (wtotal <= 96 * 1024) ? 1 : 0;
    if (const char* e = getenv("TGCN_T3_WRES")) p.w_resident = (atoi(e) != 0 && wtotal <= 96 * 1024) ? 1 : 0;
    const size_t stage = 2 * (size_t)kT3Tile + (p.w_resident ? 0 : p.wunit);
    int NS = (int)((kT3SmemLimit - fixed - (p.w_resident ? wtotal : 0)) / stage);
    if (NS > kT3MaxStages) NS = kT3MaxStages;
    if (const char* e = getenv("TGCN_T3_NS")) { const int v = atoi(e); if (v >= 2 && v <= NS) NS = v; }
    if (NS < 2) return TGCN_OK;
    p.NS = NS;
    // issuer warps: each owns 2 GP accumulator columns in each of the two row-tile buffers (512 TMEM columns in all)
    int NI = 512 / (4 * GP);
    if (NI > kT3Issuers) NI = kT3Issuers;
    if (NI > K * p.DB) NI = K * p.DB;
    // at least two stages per issuer ring: with one, an issuer's next tile cannot land while it multiplies the current
    // one (mesh layer 1, 4 stages: 2 issuers x 2 stages 123.5 us, 4 issuers x 1 stage 136.0 us)
    if (NI > NS / 2) NI = NS / 2;
    if (const char* e = getenv("TGCN_T3_NI")) { const int v = atoi(e); if (v >= 1 && v <= kT3Issuers && v <= 512 / (4 * GP)) NI = v; }
    if (NI > NS) NI = NS;
    if (NI < 1) return TGCN_OK;
    p.NI = NI;
    p.NSi = NS / NI;
    p.NS = NS = p.NSi * NI;
    const size_t smem = fixed + (size_t)NS * stage + (p.w_resident ? wtotal : 0);
    CUtensorMap tmA;
    TGCN_PROPAGATE(make_tmap3(&tmA, stack, (uint64_t)D, (uint64_t)M, (uint64_t)K, (uint64_t)D * 4, (uint64_t)M * D * 4, 32, 128, 1,
                              CU_TENSOR_MAP_SWIZZLE_128B));
    TGCN_PROPAGATE(set_smem3(contract_fwd_tc3_kernel, smem, "contract_fwd_tc3"));
    const unsigned grid = (unsigned)min64(p.ntiles, kNumSMs);
    contract_fwd_tc3_kernel<<<grid, kT3ThreadsF, smem, st>>>(tmA, p);
    TGCN_LAUNCH_CHECK("contract_fwd_tc3");
    *launched = 1;
    return TGCN_OK;
}

// G_j for all orders; covered: D = 32 * DB <= 128 (one MMA N), G a multiple of 32, Q a power of two <= 128
int contract_bwd_x_tc3(const float* dout, const uint8_t* wimg, float* gstack, int Q, int N, int D, int DP, int G, int K,
                       cudaStream_t st, int* launched) {
    *launched = 0;
    const int64_t M = (int64_t)Q * N;
    if (!use_v3() || D % 32 != 0 || DP != D || D > 128 || G % 32 != 0 || G > 256 || Q < 1 || Q > 128 || (128 % Q) != 0 || M <= 0 ||
        M >= (int64_t)INT32_MAX - 256 || (reinterpret_cast<uintptr_t>(dout) & 15u) != 0 || (reinterpret_cast<uintptr_t>(gstack) & 15u) != 0)
        return TGCN_OK;
    BwdX3Params p{};
    p.wimg = wimg; p.gstack = gstack; p.M = (int)M; p.Q = Q; p.N = N; p.D = D; p.DP = DP; p.G = G; p.GB = G / 32; p.K = K;
    p.nt = 128 / Q;
    p.ntiles = (int)ceil_div(N, p.nt);
    p.wunit = (uint32_t)p.GB * 2u * (uint32_t)DP * kRowBytes;
    const size_t abytes = (size_t)p.GB * 2 * kT3Tile, fixed = 1024 + 256;
    const size_t wring = (size_t)kX3WRing * p.wunit;
    if (fixed + wring + abytes > kT3SmemLimit) return TGCN_OK;
    p.abufs = (fixed + wring + 2 * abytes <= kT3SmemLimit) ? 2 : 1;
    int NI = 512 / (4 * DP);                                   // two [128 x 2 DP] accumulators per issuer
    if (NI > kT3Issuers) NI = kT3Issuers;
    if (NI > K) NI = K;
    if (const char* e = getenv("TGCN_T3_NI")) { const int v = atoi(e); if (v >= 1 && v <= NI) NI = v; }
    if (NI < 1) return TGCN_OK;
    p.NI = NI;
    const size_t smem = fixed + wring + (size_t)p.abufs * abytes;
    CUtensorMap tmD;
    TGCN_PROPAGATE(make_tmap3(&tmD, dout, (uint64_t)G, (uint64_t)N, (uint64_t)Q, (uint64_t)G * 4, (uint64_t)N * G * 4, 32, (uint32_t)p.nt,
                              (uint32_t)Q, CU_TENSOR_MAP_SWIZZLE_128B));
    TGCN_PROPAGATE(set_smem3(contract_bwd_x_tc3_kernel, smem, "contract_bwd_x_tc3"));
    const unsigned grid = (unsigned)min64(p.ntiles, kNumSMs);
    contract_bwd_x_tc3_kernel<<<grid, kT3ThreadsF, smem, st>>>(tmD, p);
    TGCN_LAUNCH_CHECK("contract_bwd_x_tc3");
    *launched = 1;
    return TGCN_OK;
}

struct BwdW3Plan { bool ok; int GPw, DB, NBtot, NB, NY, MT, RU, NS, NI, P, units_per_cta, total_units, dpg; size_t smem; };

static BwdW3Plan make_bwd_w3_plan(int Q, int N, int D, int G, int K) {
    BwdW3Plan t{};
    const int64_t M = (int64_t)Q * N;
    if (!use_v3() || D % 32 != 0 || G < 1 || G > 128 || Q < 1 || M <= 0 || M >= (int64_t)INT32_MAX - 256) return t;
    t.GPw = (G + 31) / 32 * 32;
    t.DB = D / 32;
    t.NBtot = K * t.DB;
    const int max_tiles = 512 / (2 * t.GPw);                 // TMEM columns: one [128 x 2 GPw] accumulator per output tile
    if (max_tiles < 1) return t;
    // One tensor copy per (unit, column block) with a box of K orders when a y group can hold whole column blocks: the producer
    // thread issued one copy per block before, and the per-unit time hardly depended on the unit's size (RU = 32 / 16 / 8: 2.35 /
    // 1.82 / 1.67 us per unit, profiles/r02/contract_tc3_notes.txt) -- the copy ISSUE was the pace of the kernel.
    static const bool fuse_ok = [] { const char* e = getenv("TGCN_T3_FUSEA"); return !(e && e[0] == '0'); }();
    static const bool fuse_wide = [] { const char* e = getenv("TGCN_T3_FUSEA"); return !(e && e[0] == '1'); }();   // "1": D = 32 layers only
    if (fuse_ok && K <= 4 * max_tiles && K <= 256 && (fuse_wide || t.DB == 1)) {
        int dpg = (4 * max_tiles) / K;
        if (dpg > t.DB) dpg = t.DB;
        t.NY = (t.DB + dpg - 1) / dpg;
        t.dpg = (t.DB + t.NY - 1) / t.NY;                    // balance the y groups
        t.NB = K * t.dpg;
    } else {
        t.dpg = 0;
        t.NB = t.NBtot < 4 * max_tiles ? t.NBtot : 4 * max_tiles;
        t.NY = (t.NBtot + t.NB - 1) / t.NB;
        t.NB = (t.NBtot + t.NY - 1) / t.NY;                  // balance the y groups
    }
    t.MT = (t.NB + 3) / 4;
    const size_t fixed = 1024 + 256;
    int ru_max = 32;                                          // TGCN_T3_RU: experiment knob (rows per unit <= value)
    if (const char* e = getenv("TGCN_T3_RU")) { const int v = atoi(e); if (v >= 8) ru_max = v; }
    for (int ru : {32, 16, 8}) {
        // a unit's rows must be whole vertices (its dOut box is [ru / Q vertices] x [Q samples]) unless Q > ru
        if (ru % Q != 0 || ru > ru_max) continue;
        const size_t blk = (size_t)ru * kRowBytes;
        const size_t stage = 2 * (size_t)t.NB * blk + 2 * (size_t)(t.GPw / 32) * blk;
        int ns = (int)((kT3SmemLimit - fixed - 4 * blk) / stage);       // 4 blocks of slack behind the last stage
        if (ns > kT3MaxStages) ns = kT3MaxStages;
        if (ns >= 2) { t.RU = ru; t.NS = ns; t.smem = fixed + (size_t)ns * stage + 4 * blk; break; }
    }
    if (t.RU == 0) return t;
    if (const char* e = getenv("TGCN_T3_NS")) { const int v = atoi(e); if (v >= 2 && v <= t.NS) t.NS = v; }
    t.NI = t.MT < kT3Issuers ? t.MT : kT3Issuers;
    if (const char* e = getenv("TGCN_T3_NI")) { const int v = atoi(e); if (v >= 1 && v <= t.NI) t.NI = v; }
    t.total_units = (int)ceil_div(M, t.RU);
    int64_t want = kNumSMs / t.NY;
    if (want < 1) want = 1;
    if (want > t.total_units) want = t.total_units;
    t.units_per_cta = (int)ceil_div(t.total_units, want);
    t.P = (int)ceil_div(t.total_units, t.units_per_cta);
    t.ok = true;
    return t;
}

int bwd_w3_partials(int Q, int N, int D, int G, int K) {
    const BwdW3Plan t = make_bwd_w3_plan(Q, N, D, G, K);
    return t.ok ? t.P : 0;
}

int contract_bwd_w_tc3(const float* stack, const float* dout, float* partial, int* P_out, int Q, int N, int D, int G, int K,
                       cudaStream_t st, int* launched) {
    *launched = 0;
    const BwdW3Plan t = make_bwd_w3_plan(Q, N, D, G, K);
    if (!t.ok || (reinterpret_cast<uintptr_t>(stack) & 15u) != 0 || (reinterpret_cast<uintptr_t>(dout) & 15u) != 0 || (G % 4) != 0)
        return TGCN_OK;
    const int64_t M = (int64_t)Q * N;
    BwdW3Params p{};
    p.partial = partial; p.M = (int)M; p.Q = Q; p.N = N; p.D = D; p.G = G; p.GPw = t.GPw; p.K = K; p.DB = t.DB;
    p.NB = t.NB; p.MT = t.MT; p.RU = t.RU; p.NS = t.NS; p.NI = t.NI; p.units_per_cta = t.units_per_cta; p.total_units = t.total_units;
    p.dpg = t.dpg;
    CUtensorMap tmA, tmD;
    TGCN_PROPAGATE(make_tmap3(&tmA, stack, (uint64_t)D, (uint64_t)M, (uint64_t)K, (uint64_t)D * 4, (uint64_t)M * D * 4, 32, (uint32_t)t.RU,
                              p.dpg > 0 ? (uint32_t)K : 1u, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    // dOut[Q][N][G] addressed vertex-major: dimension 1 = sample (stride N*G), dimension 2 = vertex (stride G)
    TGCN_PROPAGATE(make_tmap3(&tmD, dout, (uint64_t)G, (uint64_t)Q, (uint64_t)N, (uint64_t)N * G * 4, (uint64_t)G * 4, 32, (uint32_t)Q,
                              (uint32_t)(t.RU / Q), CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    TGCN_PROPAGATE(set_smem3(contract_bwd_w_tc3_kernel, t.smem, "contract_bwd_w_tc3"));
    const dim3 grid((unsigned)t.P, (unsigned)t.NY);
    contract_bwd_w_tc3_kernel<<<grid, kT3ThreadsF, t.smem, st>>>(tmA, tmD, p);
    TGCN_LAUNCH_CHECK("contract_bwd_w_tc3");
    *P_out = t.P;
    *launched = 1;
    return TGCN_OK;
}

}  // namespace tgcn
