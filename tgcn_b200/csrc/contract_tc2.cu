// Second-generation tensor-core contraction kernels: persistent, warp-specialised, TMA-fed.
//
//   warp 8      : producer  -- 1-D bulk TMA copies (cp.async.bulk, completion on an mbarrier) of the raw fp32
//                 operand chunks into a ring of shared-memory slots, several slots ahead of the math
//   warps 0..7  : transform -- raw fp32 -> (tf32 hi, lo) swizzled UMMA operand tiles; later the TMEM epilogue
//   warp 9      : MMA       -- one thread issues tcgen05.mma (3xTF32) and commits to mbarriers
//
// The first-generation kernels (contract_tc.cu) staged global -> registers -> shared memory with one unit of
// prefetch and were latency bound (ncu: 12 % warps active, 8-22 % DRAM throughput, profiles/r01).  Here the
// bytes in flight per SM are set by the ring depth (3 slots x 20-30 KB), independent of registers.
//
// Covered shapes: D <= 32 (one 32-column block per order), stack rows 16-byte aligned; everything else falls
// back to the first-generation kernels.
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace tgcn {
using namespace tc;

constexpr int kT2Threads = 320;       // 8 transform warps + producer warp + MMA warp
constexpr int kT2Transform = 8;       // transform / epilogue warps
constexpr int kT2RawMax = 6;          // raw ring depth (run-time value R <= kT2RawMax, sized to fill shared memory)
constexpr int kT2Op = 2;              // operand stage depth (bwd_w)
constexpr int kT2OpF = 2;             // operand stage depth of the forward kernel (3 measured slower: fewer raw ring slots)
constexpr uint32_t kT2TileBytes = 128 * 128;

// Align the dynamic shared-memory window to 1024 B with pointer arithmetic on the __shared__ array itself,
// so that the compiler keeps the shared address space (LDS/STS instead of generic LD/ST).
__device__ __forceinline__ uint8_t* align1024_2(uint8_t* p) {
    const uint32_t a = tc::smem_u32(p);
    return p + (((a + 1023u) & ~1023u) - a);
}

template <int VEC>
__device__ __forceinline__ void lds_vec(const uint8_t* p, float* v) {
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *reinterpret_cast<const float*>(p);
    }
}

template <int VEC>
__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, uint32_t off, const float* x) {
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

// Shared-memory vector access by 32-bit shared-window address (no generic-address arithmetic in the hot loops).
template <int VEC>
__device__ __forceinline__ void lds_vec_u32(uint32_t addr, float* v) {
    if constexpr (VEC == 4) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
    } else if constexpr (VEC == 2) {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr));
    } else {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr));
    }
}

template <int VEC>
__device__ __forceinline__ void split_store_u32(uint32_t hi, uint32_t lo, const float* x) {
    float h[VEC], l[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { h[i] = tf32_hi(x[i]); l[i] = tf32_hi(x[i] - h[i]); }
    if constexpr (VEC == 4) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi), "f"(h[0]), "f"(h[1]), "f"(h[2]), "f"(h[3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo), "f"(l[0]), "f"(l[1]), "f"(l[2]), "f"(l[3]) : "memory");
    } else if constexpr (VEC == 2) {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(hi), "f"(h[0]), "f"(h[1]) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(lo), "f"(l[0]), "f"(l[1]) : "memory");
    } else {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(hi), "f"(h[0]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(lo), "f"(l[0]) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// forward:  out[m, g] = sum_j sum_d P_j[m, d] W'_j[d, g] + bias
// unit = (tile of 128 pairs, order j): raw slot = [128 x D fp32 (contiguous in the slab) | weight image of j]
// ------------------------------------------------------------------------------------------------
struct Fwd2Params {
    const float* stack; int64_t S;
    const uint8_t* wimg;           // per order: [hi | lo] x [GP rows x 128 B], K-major SWIZZLE_128B image
    const float* bias; int bias_mode;
    float* out;
    int M, Q, N, D, G, GP, K;
    int ntiles;
    uint32_t rawA_bytes;           // slot size reserved for the A chunk (128*D*4 rounded up to 1024)
    int R;                         // raw ring depth
    int dbg;
};

template <int VEC>
__global__ void __launch_bounds__(kT2Threads, 1)
contract_fwd_tc2_kernel(const Fwd2Params p) {
    constexpr int LPR = 32 / VEC;
    constexpr int NI = 16 / VEC;                       // instructions per transform warp per unit (16 rows)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024_2(smem_raw);
    const uint32_t wbytes = 2u * (uint32_t)p.GP * kRowBytes;
    const uint32_t op_bytes = 2 * kT2TileBytes + wbytes;          // [A hi | A lo | W hi | W lo]
    const uint32_t raw_bytes = p.rawA_bytes + wbytes;             // [A raw | W image]
    uint8_t* op_base = smem;
    uint8_t* raw_base = smem + kT2OpF * op_bytes;
    __shared__ __align__(8) uint64_t raw_full[kT2RawMax], raw_free[kT2RawMax], op_full[kT2OpF], op_free[kT2OpF], acc_full, acc_free;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t R = (uint32_t)p.R;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)p.GP);
    if (tid == 0) {
        for (int i = 0; i < kT2RawMax; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], kT2Transform); }
        for (int i = 0; i < kT2OpF; ++i) { mbar_init(&op_full[i], kT2Transform); mbar_init(&op_free[i], 1); }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_free, kT2Transform);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    // zero the A operand tiles once: the fast transform path never writes the padding columns (d >= D)
    for (int sidx = 0; sidx < kT2OpF; ++sidx)
        for (uint32_t i = tid; i < 2 * kT2TileBytes / 16; i += kT2Threads)
            reinterpret_cast<float4*>(op_base + sidx * op_bytes)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = tmem_base_s;

    if (warp == kT2Transform) {
        // ===================== producer =====================
        // lane 0 waits for the slot and posts the expected byte count; lanes 0..3 each copy a quarter of the
        // A chunk and lane 4 the weight image, so five bulk copies per unit are in flight at once
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int m0 = tile * 128;
            const int rows = min(128, p.M - m0);
            const uint32_t bytesA = (uint32_t)rows * (uint32_t)p.D * 4u;
            const bool direct = (bytesA & 15u) != 0;          // ragged tail: transform warps read global themselves
            const uint32_t piece = ((bytesA / 4u) + 15u) & ~15u;      // quarter, rounded up to 16 B
            for (int j = 0; j < p.K; ++j, ++g) {
                const uint32_t r = g % R;
                uint8_t* slot = raw_base + r * raw_bytes;
                if (lane == 0) {
                    if (g >= R) mbar_wait(&raw_free[r], ((g / R) - 1) & 1);
                    mbar_arrive_expect_tx(&raw_full[r], ((p.dbg & 16) ? 0u : wbytes) + (direct ? 0u : bytesA));
                }
                __syncwarp();
                if (lane < 4 && !direct) {
                    const uint32_t off = (uint32_t)lane * piece;
                    if (off < bytesA) {
                        const uint32_t len = min(piece, bytesA - off);
                        bulk_g2s(slot + off, reinterpret_cast<const uint8_t*>(p.stack + (int64_t)j * p.S + (int64_t)m0 * p.D) + off,
                                 len, &raw_full[r]);
                    }
                } else if (lane == 4 && !(p.dbg & 16)) {
                    bulk_g2s(slot + p.rawA_bytes, p.wimg + (int64_t)j * wbytes, wbytes, &raw_full[r]);
                }
                __syncwarp();
            }
        }
    } else if (warp == kT2Transform + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, (uint32_t)p.GP, 0, 0);
            const int nks = (p.D + 7) >> 3;
            uint32_t g = 0, ti = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
                if (ti >= 1) mbar_wait(&acc_free, (ti - 1) & 1);          // epilogue of the previous tile drained TMEM
                tcgen05_fence_after();
                for (int j = 0; j < p.K; ++j, ++g) {
                    const uint32_t s = g % kT2OpF;
                    mbar_wait(&op_full[s], (g / kT2OpF) & 1);
                    tcgen05_fence_after();
                    uint8_t* op = op_base + s * op_bytes;
                    const uint64_t dah = make_desc_kmajor(smem_u32(op)), dal = make_desc_kmajor(smem_u32(op + kT2TileBytes));
                    const uint64_t dbh = make_desc_kmajor(smem_u32(op + 2 * kT2TileBytes));
                    const uint64_t dbl = make_desc_kmajor(smem_u32(op + 2 * kT2TileBytes + wbytes / 2));
                    for (int ks = 0; ks < nks; ++ks) {
                        if (p.dbg & 1) break;
                        const uint64_t adv = (uint64_t)(ks * 2);
                        umma_tf32(tmem_acc, dal + adv, dbh + adv, idesc, (j | ks) ? 1u : 0u);
                        umma_tf32(tmem_acc, dah + adv, dbl + adv, idesc, 1u);
                        umma_tf32(tmem_acc, dah + adv, dbh + adv, idesc, 1u);
                    }
                    umma_commit(&op_free[s]);
                }
                umma_commit(&acc_full);
            }
        }
    } else {
        // ===================== transform + epilogue =====================
        const int sub = lane / LPR, c0 = (lane % LPR) * VEC;
        const bool cok = c0 < p.D;
        const int lg = warp & 3, ch = warp >> 2;              // TMEM lane group / column-chunk parity of this warp
        // unit-invariant addresses of this thread's NI vectors: raw slot offset and swizzled operand offset
        uint32_t src_off[NI], dst_off[NI];
#pragma unroll
        for (int e = 0; e < NI; ++e) {
            const uint32_t row = (uint32_t)(warp * 16 + e * VEC + sub);
            src_off[e] = (row * (uint32_t)p.D + (uint32_t)c0) * 4u;
            dst_off[e] = sw128_offset(row, (uint32_t)c0);
        }
        const uint32_t raw_u32 = smem_u32(raw_base), op_u32 = smem_u32(op_base);
        const int wpieces_f = (int)(wbytes / 16);
        uint32_t g = 0, ti = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++ti) {
            const int m0 = tile * 128;
            const int rows = min(128, p.M - m0);
            const bool direct = (((uint32_t)rows * (uint32_t)p.D * 4u) & 15u) != 0;
            if (!direct && p.dbg == 0) {
                // fast path: branch-free, all shared loads of a unit issued back to back
                for (int j = 0; j < p.K; ++j, ++g) {
                    const uint32_t r = g % R, s = g % kT2OpF;
                    const uint32_t slot_u32 = raw_u32 + r * raw_bytes;
                    float buf[16];
                    float4 wv[4];
                    mbar_wait(&raw_full[r], (g / R) & 1);
                    if (cok) {
#pragma unroll
                        for (int e = 0; e < NI; ++e) lds_vec_u32<VEC>(slot_u32 + src_off[e], &buf[e * VEC]);
                    }
                    const uint32_t wsrc_u32 = slot_u32 + p.rawA_bytes;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = tid + i * 256;
                        if (idx < wpieces_f) {
                            float t4[4];
                            lds_vec_u32<4>(wsrc_u32 + (uint32_t)idx * 16u, t4);
                            wv[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                        }
                    }
                    if (g >= kT2OpF) mbar_wait(&op_free[s], ((g / kT2OpF) - 1) & 1);
                    const uint32_t opa = op_u32 + s * op_bytes;
                    if (cok) {
#pragma unroll
                        for (int e = 0; e < NI; ++e)
                            split_store_u32<VEC>(opa + dst_off[e], opa + kT2TileBytes + dst_off[e], &buf[e * VEC]);
                    }
                    const uint32_t wdst_u32 = opa + 2 * kT2TileBytes;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = tid + i * 256;
                        if (idx < wpieces_f)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wdst_u32 + (uint32_t)idx * 16u),
                                         "f"(wv[i].x), "f"(wv[i].y), "f"(wv[i].z), "f"(wv[i].w) : "memory");
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&op_full[s]);
                        mbar_arrive(&raw_free[r]);
                    }
                }
            } else
            for (int j = 0; j < p.K; ++j, ++g) {
                const uint32_t r = g % R, s = g % kT2OpF;
                const uint8_t* slot = raw_base + r * raw_bytes;
                float buf[16];
                mbar_wait(&raw_full[r], (g / R) & 1);
#pragma unroll
                for (int e = 0; e < NI; ++e) {
                    const int row = warp * 16 + e * VEC + sub;
                    if (!cok) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) buf[e * VEC + i] = 0.f;
                    } else if (!direct) {
                        if (p.dbg & 4) { for (int i = 0; i < VEC; ++i) buf[e * VEC + i] = 1.f; } else
                        lds_vec<VEC>(slot + ((uint32_t)row * (uint32_t)p.D + (uint32_t)c0) * 4u, &buf[e * VEC]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i)
                            buf[e * VEC + i] = (row < rows) ? __ldg(p.stack + (int64_t)j * p.S + (int64_t)(m0 + row) * p.D + c0 + i) : 0.f;
                    }
                }
                // this warp's share of the weight image (16-byte pieces)
                float4 wv[4];
                const int wpieces = (int)(wbytes / 16);               // GP * 16
                const float4* wsrc = reinterpret_cast<const float4*>(slot + p.rawA_bytes);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = tid + i * 256;
                    if (idx < wpieces && !(p.dbg & 16)) wv[i] = wsrc[idx];
                }
                if (g >= kT2OpF) mbar_wait(&op_free[s], ((g / kT2OpF) - 1) & 1);
                uint8_t* op = op_base + s * op_bytes;
                if (!(p.dbg & 2)) {
#pragma unroll
                for (int e = 0; e < NI; ++e)
                    split_store<VEC>(op, op + kT2TileBytes, sw128_offset((uint32_t)(warp * 16 + e * VEC + sub), (uint32_t)c0),
                                     &buf[e * VEC]);
                }
                float4* wdst = reinterpret_cast<float4*>(op + 2 * kT2TileBytes);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = tid + i * 256;
                    if (idx < wpieces && !(p.dbg & 16)) wdst[idx] = wv[i];
                }
                // The raw slot is released only here: the stores above consumed every register loaded from it,
                // so no load from the slot can still be in flight when the producer's next bulk copy lands
                // (an arrive placed right after the loads let the TMA overwrite data not yet read: WAR race).
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&op_full[s]);
                    mbar_arrive(&raw_free[r]);
                }
            }
            // ---- epilogue of this tile ----
            mbar_wait(&acc_full, ti & 1);
            tcgen05_fence_after();
            const int m = m0 + lg * 32 + lane;
            const bool live = m < p.M;
            int n = 0, q = 0;
            if (live) { n = m / p.Q; q = m - n * p.Q; }
            float* dst = p.out + ((int64_t)q * p.N + n) * p.G;
            const bool vec = (p.G % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0);
            for (int cb = ch * 16; cb < p.GP; cb += 32) {
                float v[16];
                tmem_ld16(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)cb, v);
                if (!live || (p.dbg & 32)) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int gc = cb + i;
                    if (gc < p.G) {
                        if (p.bias_mode == TGCN_BIAS_PER_VERTEX) v[i] += __ldg(p.bias + (int64_t)n * p.G + gc);
                        else if (p.bias_mode == TGCN_BIAS_PER_FILTER) v[i] += __ldg(p.bias + gc);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_acc, ncols);
}

// ------------------------------------------------------------------------------------------------
// backward w.r.t. the mixed weights: dW'[(j,d), g] = sum_m P_j[m, d] dOut[m, g]      (MN-major operands)
// unit = 16 pairs: raw slot = [jc chunks of 16 x D fp32 (each contiguous in its slab) | 16 dOut rows of G fp32]
// ------------------------------------------------------------------------------------------------
constexpr int kBw2KT = 16;

struct BwdW2Params {
    const float* stack; int64_t S;
    const float* dout;
    float* partial;                // [P][K*D][G]
    int M, Q, N, D, G, GP, K;
    int DPAD, JC, MT;
    int units_per_cta;
    uint32_t rawA_bytes;           // JC * 16 * D * 4 rounded up to 128
    int R;                         // raw ring depth
};

template <int VEC>
__global__ void __launch_bounds__(kT2Threads, 1)
contract_bwd_w_tc2_kernel(const BwdW2Params p) {
    constexpr int LPR = 32 / VEC;
    constexpr int RG = kBw2KT / VEC;                    // row groups per order
    constexpr int NSLOT = 24 / VEC;                     // staged A vectors per lane (JC <= 12 -> JC*RG/8 items per warp)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024_2(smem_raw);
    const uint32_t blk = kBw2KT * kRowBytes;            // one 32-wide M/N block: 16 rows x 128 B
    const uint32_t a_part = (uint32_t)p.MT * 4 * blk;
    const uint32_t b_part = (uint32_t)(p.GP / 32) * blk;
    const uint32_t op_bytes = 2 * a_part + 2 * b_part;  // [A hi | A lo | B hi | B lo]
    const uint32_t rowG = (uint32_t)p.G * 4u;
    const uint32_t raw_bytes = p.rawA_bytes + kBw2KT * rowG;
    uint8_t* op_base = smem;
    uint8_t* raw_base = smem + kT2Op * op_bytes;
    __shared__ __align__(8) uint64_t raw_full[kT2RawMax], raw_free[kT2RawMax], op_full[kT2Op], op_free[kT2Op], acc_full;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t R = (uint32_t)p.R;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)(p.MT * p.GP));
    const int j0 = blockIdx.y * p.JC;
    const int jc = min(p.JC, p.K - j0);
    const int mn_cnt = jc * p.DPAD;
    const int total_units = (p.M + kBw2KT - 1) / kBw2KT;
    const int u_begin = blockIdx.x * p.units_per_cta;
    const int u_end = min(total_units, u_begin + p.units_per_cta);
    const uint32_t chunkA = (uint32_t)kBw2KT * (uint32_t)p.D * 4u;      // one order's 16 rows

    if (tid == 0) {
        for (int i = 0; i < kT2RawMax; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_free[i], kT2Transform); }
        for (int i = 0; i < kT2Op; ++i) { mbar_init(&op_full[i], kT2Transform); mbar_init(&op_free[i], 1); }
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
    // zero the operand stages once: padding rows / columns are never written afterwards
    for (uint32_t i = tid; i < kT2Op * op_bytes / 16; i += kT2Threads) reinterpret_cast<float4*>(op_base)[i] = make_float4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == kT2Transform) {
        // ===================== producer (all lanes issue copies) =====================
        uint32_t g = 0;
        for (int u = u_begin; u < u_end; ++u, ++g) {
            const uint32_t r = g % R;
            const int mbase = u * kBw2KT;
            const int rows = min(kBw2KT, p.M - mbase);
            const bool direct = rows < kBw2KT;                    // ragged last unit: transform warps read global
            uint8_t* slot = raw_base + r * raw_bytes;
            if (lane == 0) {
                if (g >= R) mbar_wait(&raw_free[r], ((g / R) - 1) & 1);
                mbar_arrive_expect_tx(&raw_full[r], direct ? 0u : (uint32_t)jc * chunkA + (uint32_t)rows * rowG);
            }
            __syncwarp();
            if (!direct) {
                for (int jl = lane; jl < jc; jl += 32)
                    bulk_g2s(slot + (uint32_t)jl * chunkA, p.stack + (int64_t)(j0 + jl) * p.S + (int64_t)mbase * p.D, chunkA,
                             &raw_full[r]);
                for (int k = lane; k < rows; k += 32) {
                    const int m = mbase + k;
                    const int n = m / p.Q, q = m - n * p.Q;
                    bulk_g2s(slot + p.rawA_bytes + (uint32_t)k * rowG, p.dout + ((int64_t)q * p.N + n) * p.G, rowG, &raw_full[r]);
                }
            }
            __syncwarp();
        }
    } else if (warp == kT2Transform + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, (uint32_t)p.GP, 1, 1);
            uint32_t g = 0;
            for (int u = u_begin; u < u_end; ++u, ++g) {
                const uint32_t s = g % kT2Op;
                mbar_wait(&op_full[s], (g / kT2Op) & 1);
                tcgen05_fence_after();
                uint8_t* ah = op_base + s * op_bytes;
                uint8_t* al = ah + a_part;
                uint8_t* bh = ah + 2 * a_part;
                uint8_t* bl = bh + b_part;
                for (int t = 0; t < p.MT; ++t) {
                    if (t * 128 >= mn_cnt) break;
                    const uint32_t acc = tmem_base + (uint32_t)(t * p.GP);
                    const uint32_t aoff = (uint32_t)t * 4 * blk;
                    for (int ks = 0; ks < kBw2KT / 8; ++ks) {
                        const uint32_t adv = (uint32_t)ks * kAtomBytes;
                        const uint64_t dah = make_desc_mnmajor(smem_u32(ah + aoff + adv), blk);
                        const uint64_t dal = make_desc_mnmajor(smem_u32(al + aoff + adv), blk);
                        const uint64_t dbh = make_desc_mnmajor(smem_u32(bh + adv), blk);
                        const uint64_t dbl = make_desc_mnmajor(smem_u32(bl + adv), blk);
                        umma_tf32(acc, dal, dbh, idesc, (g | ks) ? 1u : 0u);
                        umma_tf32(acc, dah, dbl, idesc, 1u);
                        umma_tf32(acc, dah, dbh, idesc, 1u);
                    }
                }
                umma_commit(&op_free[s]);
            }
            umma_commit(&acc_full);
        }
    } else {
        // ===================== transform + epilogue =====================
        const int sub = lane / LPR, c0 = (lane % LPR) * VEC;
        const bool cok = c0 < p.D;
        const int n_items = jc * RG;                               // (order, row group) pairs, dealt to the 8 warps
        const int GV = p.G >> 2;                                   // dOut float4 per row
        const int n_bitems = kBw2KT * GV;
        // unit-invariant addresses of this thread's staged vectors (raw slot offset, swizzled operand offset)
        uint32_t a_src[NSLOT], a_dst[NSLOT], b_src[2], b_dst[2];
#pragma unroll
        for (int e = 0; e < NSLOT; ++e) {
            const int item = warp + kT2Transform * e;
            const int jl = item / RG, rg = item - jl * RG;
            const int k = rg * VEC + sub;
            const int mn = jl * p.DPAD + c0;
            a_src[e] = (uint32_t)jl * chunkA + ((uint32_t)k * (uint32_t)p.D + (uint32_t)c0) * 4u;
            a_dst[e] = (uint32_t)(mn >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(mn & 31));
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int item = tid + 256 * e;
            const int k = item / GV, gv = item - k * GV, gc = gv * 4;
            b_src[e] = p.rawA_bytes + (uint32_t)k * rowG + (uint32_t)gv * 16u;
            b_dst[e] = (uint32_t)(gc >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(gc & 31));
        }
        const uint32_t raw_u32 = smem_u32(raw_base), op_u32 = smem_u32(op_base);
        uint32_t g = 0;
        for (int u = u_begin; u < u_end; ++u, ++g) {
            const uint32_t r = g % R, s = g % kT2Op;
            const uint8_t* slot = raw_base + r * raw_bytes;
            const int mbase = u * kBw2KT;
            const int rows = min(kBw2KT, p.M - mbase);
            const bool direct = rows < kBw2KT;
            float abuf[24];
            float4 bbuf[2];
            if (!direct) {
                // fast path: branch-free, every shared load of the unit issued back to back
                const uint32_t slot_u32 = raw_u32 + r * raw_bytes;
                float bb[2][4];
                mbar_wait(&raw_full[r], (g / R) & 1);
                if (cok) {
#pragma unroll
                    for (int e = 0; e < NSLOT; ++e)
                        if (warp + kT2Transform * e < n_items) lds_vec_u32<VEC>(slot_u32 + a_src[e], &abuf[e * VEC]);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (tid + 256 * e < n_bitems) lds_vec_u32<4>(slot_u32 + b_src[e], bb[e]);
                if (g >= kT2Op) mbar_wait(&op_free[s], ((g / kT2Op) - 1) & 1);
                const uint32_t ah32 = op_u32 + s * op_bytes, al32 = ah32 + a_part;
                const uint32_t bh32 = ah32 + 2 * a_part, bl32 = bh32 + b_part;
                if (cok) {
#pragma unroll
                    for (int e = 0; e < NSLOT; ++e)
                        if (warp + kT2Transform * e < n_items)
                            split_store_u32<VEC>(ah32 + a_dst[e], al32 + a_dst[e], &abuf[e * VEC]);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (tid + 256 * e < n_bitems) split_store_u32<4>(bh32 + b_dst[e], bl32 + b_dst[e], bb[e]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&op_full[s]);
                    mbar_arrive(&raw_free[r]);
                }
                continue;
            }
            mbar_wait(&raw_full[r], (g / R) & 1);
#pragma unroll
            for (int e = 0; e < NSLOT; ++e) {
                const int item = warp + kT2Transform * e;
                if (item < n_items) {
                    const int jl = item / RG, rg = item - jl * RG;
                    const int k = rg * VEC + sub;
                    if (!cok) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) abuf[e * VEC + i] = 0.f;
                    } else if (!direct) {
                        lds_vec<VEC>(slot + (uint32_t)jl * chunkA + ((uint32_t)k * (uint32_t)p.D + (uint32_t)c0) * 4u, &abuf[e * VEC]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i)
                            abuf[e * VEC + i] = (k < rows) ? __ldg(p.stack + (int64_t)(j0 + jl) * p.S + (int64_t)(mbase + k) * p.D + c0 + i) : 0.f;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int item = tid + 256 * e;
                if (item < n_bitems) {
                    const int k = item / GV, gv = item - k * GV;
                    if (!direct) {
                        bbuf[e] = *reinterpret_cast<const float4*>(slot + p.rawA_bytes + (uint32_t)k * rowG + (uint32_t)gv * 16u);
                    } else if (k < rows) {
                        const int m = mbase + k;
                        const int n = m / p.Q, q = m - n * p.Q;
                        bbuf[e] = __ldg(reinterpret_cast<const float4*>(p.dout + ((int64_t)q * p.N + n) * p.G) + gv);
                    } else {
                        bbuf[e] = make_float4(0, 0, 0, 0);
                    }
                }
            }
            if (g >= kT2Op) mbar_wait(&op_free[s], ((g / kT2Op) - 1) & 1);
            uint8_t* ah = op_base + s * op_bytes;
            uint8_t* al = ah + a_part;
            uint8_t* bh = ah + 2 * a_part;
            uint8_t* bl = bh + b_part;
#pragma unroll
            for (int e = 0; e < NSLOT; ++e) {
                const int item = warp + kT2Transform * e;
                if (item < n_items && cok) {
                    const int jl = item / RG, rg = item - jl * RG;
                    const int k = rg * VEC + sub;
                    const int mn = jl * p.DPAD + c0;
                    const uint32_t off = (uint32_t)(mn >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(mn & 31));
                    split_store<VEC>(ah, al, off, &abuf[e * VEC]);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int item = tid + 256 * e;
                if (item < n_bitems) {
                    const int k = item / GV, gv = item - k * GV;
                    const int gc = gv * 4;
                    const uint32_t off = (uint32_t)(gc >> 5) * blk + sw128b32_offset((uint32_t)k, (uint32_t)(gc & 31));
                    const float x[4] = {bbuf[e].x, bbuf[e].y, bbuf[e].z, bbuf[e].w};
                    split_store<4>(bh, bl, off, x);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&op_full[s]);
                mbar_arrive(&raw_free[r]);       // after the stores consumed everything loaded from the slot (see fwd)
            }
        }
        // ---- epilogue: partial[blockIdx.x][(j,d)][g] ----
        float* dst_base = p.partial + (int64_t)blockIdx.x * p.K * p.D * p.G;
        const int lg = warp & 3, ch = warp >> 2;
        if (u_begin < u_end) {
            mbar_wait(&acc_full, 0);
            tcgen05_fence_after();
            for (int t = 0; t < p.MT; ++t) {
                if (t * 128 >= mn_cnt) break;
                const int mn = t * 128 + lg * 32 + lane;
                const int jl = mn / p.DPAD, d = mn - jl * p.DPAD;
                const bool ok = mn < mn_cnt && d < p.D;
                float* dst = dst_base + ((int64_t)(j0 + jl) * p.D + d) * p.G;
                for (int cb = ch * 16; cb < p.GP; cb += 32) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * p.GP + cb), v);
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cb + i < p.G) dst[cb + i] = v[i];
                    }
                }
            }
        } else {
            for (int i = tid; i < jc * p.D * p.G; i += 256) dst_base[(int64_t)j0 * p.D * p.G + i] = 0.f;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int round_up2(int x, int m) { return (x + m - 1) / m * m; }
constexpr size_t kT2SmemLimit = 220 * 1024;

template <typename Kern>
static int set_smem2(Kern kern, size_t bytes, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu): %s", name, bytes, cudaGetErrorString(e));
    return TGCN_OK;
}

static bool aligned_rows16(const void* base, int64_t S) {
    return (reinterpret_cast<uintptr_t>(base) % 16 == 0) && (S % 4 == 0);
}

// returns 1 when the shape is covered (and the kernel was launched), 0 when the caller must use the v1 kernel
int contract_fwd_tc2(const float* stack, const uint8_t* wimg, const float* bias, int bias_mode, float* out,
                     int Q, int N, int D, int G, int GP, int K, cudaStream_t st, int* launched) {
    *launched = 0;
    const int64_t M = (int64_t)Q * N;
    const int64_t S = (int64_t)N * Q * D;
    if (D > 32 || GP > 64 || !aligned_rows16(stack, S) || M <= 0) return TGCN_OK;
    Fwd2Params p{};
    p.stack = stack; p.S = S; p.wimg = wimg; p.bias = bias; p.bias_mode = bias_mode; p.out = out;
    p.M = (int)M; p.Q = Q; p.N = N; p.D = D; p.G = G; p.GP = GP; p.K = K;
    p.ntiles = (int)ceil_div(M, 128);
    p.rawA_bytes = (uint32_t)round_up2(128 * D * 4, 1024);
    if (const char* e = getenv("TGCN_T2_PAD")) p.rawA_bytes += (uint32_t)atoi(e);
    const size_t wbytes = 2 * (size_t)GP * kRowBytes;
    const size_t op_total = 1024 + kT2OpF * (2 * (size_t)kT2TileBytes + wbytes);
    const size_t raw_slot = (size_t)p.rawA_bytes + wbytes;
    if (op_total + 2 * raw_slot > kT2SmemLimit) return TGCN_OK;
    int R = (int)((kT2SmemLimit - op_total) / raw_slot);
    if (R > kT2RawMax) R = kT2RawMax;
    if (const char* e = getenv("TGCN_T2_R")) { const int v = atoi(e); if (v >= 2 && v <= R) R = v; }
    p.R = R;
    if (const char* e = getenv("TGCN_T2_DBG")) p.dbg = atoi(e);
    const size_t smem = op_total + (size_t)R * raw_slot;
    const unsigned grid = (unsigned)min64(p.ntiles, kNumSMs);
    if (D % 4 == 0) {
        TGCN_PROPAGATE(set_smem2(contract_fwd_tc2_kernel<4>, smem, "contract_fwd_tc2"));
        contract_fwd_tc2_kernel<4><<<grid, kT2Threads, smem, st>>>(p);
    } else if (D % 2 == 0) {
        TGCN_PROPAGATE(set_smem2(contract_fwd_tc2_kernel<2>, smem, "contract_fwd_tc2"));
        contract_fwd_tc2_kernel<2><<<grid, kT2Threads, smem, st>>>(p);
    } else {
        TGCN_PROPAGATE(set_smem2(contract_fwd_tc2_kernel<1>, smem, "contract_fwd_tc2"));
        contract_fwd_tc2_kernel<1><<<grid, kT2Threads, smem, st>>>(p);
    }
    TGCN_LAUNCH_CHECK("contract_fwd_tc2");
    *launched = 1;
    return TGCN_OK;
}

// plan of the v2 bwd_w kernel; P (number of partial slabs) must match the workspace sizing in contract_tc.cu
struct BwdW2Plan { bool ok; int GP, DPAD, JC, MT, NY, P, units_per_cta, R; size_t smem; uint32_t rawA; };

static BwdW2Plan make_bwd_w2_plan(int Q, int N, int D, int G, int K) {
    BwdW2Plan t{};
    const int64_t M = (int64_t)Q * N;
    t.GP = round_up2(G, 32);
    t.DPAD = round_up2(D, 8);
    int mt = t.GP <= 256 ? 256 / t.GP : 0;
    if (mt > 3) mt = 3;
    int jc = mt > 0 ? (mt * 128) / t.DPAD : 0;
    if (jc > 12) jc = 12;
    if (jc > K) jc = K;
    t.JC = jc;
    t.MT = jc > 0 ? (jc * t.DPAD + 127) / 128 : 0;
    t.NY = jc > 0 ? (K + jc - 1) / jc : 0;
    const int64_t total_units = (M + kBw2KT - 1) / kBw2KT;
    int64_t want = t.NY > 0 ? (int64_t)kNumSMs / t.NY : 1;
    if (want < 1) want = 1;
    if (want > total_units) want = total_units > 0 ? total_units : 1;
    t.units_per_cta = (int)((total_units + want - 1) / want);
    if (t.units_per_cta < 1) t.units_per_cta = 1;
    t.P = (int)((total_units + t.units_per_cta - 1) / t.units_per_cta);
    if (t.P < 1) t.P = 1;
    const size_t blk = kBw2KT * kRowBytes;
    const size_t op = 2 * (size_t)t.MT * 4 * blk + 2 * (size_t)(t.GP / 32) * blk;
    t.rawA = (uint32_t)round_up2(jc * kBw2KT * D * 4, 128);
    const size_t raw = (size_t)t.rawA + (size_t)kBw2KT * G * 4;
    int R = (1024 + kT2Op * op + 2 * raw <= kT2SmemLimit) ? (int)((kT2SmemLimit - 1024 - kT2Op * op) / raw) : 0;
    if (R > kT2RawMax) R = kT2RawMax;
    if (const char* e = getenv("TGCN_T2_R")) { const int v = atoi(e); if (v >= 2 && v <= R) R = v; }
    t.R = R;
    t.smem = 1024 + kT2Op * op + (size_t)(R > 0 ? R : 0) * raw;
    t.ok = R >= 2 && jc >= 1 && D <= 32 && G % 4 == 0 && G <= 128 && (kBw2KT * (G / 4)) <= 512 && t.smem <= kT2SmemLimit && M > 0;
    return t;
}

int bwd_w2_partials(int Q, int N, int D, int G, int K) {
    const BwdW2Plan t = make_bwd_w2_plan(Q, N, D, G, K);
    return t.ok ? t.P : 0;
}

int contract_bwd_w_tc2(const float* stack, const float* dout, float* partial, int* P_out,
                       int Q, int N, int D, int G, int K, cudaStream_t st, int* launched) {
    *launched = 0;
    const BwdW2Plan t = make_bwd_w2_plan(Q, N, D, G, K);
    const int64_t S = (int64_t)N * Q * D;
    if (!t.ok || !aligned_rows16(stack, S) || (reinterpret_cast<uintptr_t>(dout) % 16) != 0 || ((16 * D * 4) % 16) != 0)
        return TGCN_OK;
    BwdW2Params p{};
    p.stack = stack; p.S = S; p.dout = dout; p.partial = partial;
    p.M = Q * N; p.Q = Q; p.N = N; p.D = D; p.G = G; p.GP = t.GP; p.K = K;
    p.DPAD = t.DPAD; p.JC = t.JC; p.MT = t.MT; p.units_per_cta = t.units_per_cta; p.rawA_bytes = t.rawA; p.R = t.R;
    dim3 grid((unsigned)t.P, (unsigned)t.NY);
    if (D % 4 == 0) {
        TGCN_PROPAGATE(set_smem2(contract_bwd_w_tc2_kernel<4>, t.smem, "contract_bwd_w_tc2"));
        contract_bwd_w_tc2_kernel<4><<<grid, kT2Threads, t.smem, st>>>(p);
    } else if (D % 2 == 0) {
        TGCN_PROPAGATE(set_smem2(contract_bwd_w_tc2_kernel<2>, t.smem, "contract_bwd_w_tc2"));
        contract_bwd_w_tc2_kernel<2><<<grid, kT2Threads, t.smem, st>>>(p);
    } else {
        TGCN_PROPAGATE(set_smem2(contract_bwd_w_tc2_kernel<1>, t.smem, "contract_bwd_w_tc2"));
        contract_bwd_w_tc2_kernel<1><<<grid, kT2Threads, t.smem, st>>>(p);
    }
    TGCN_LAUNCH_CHECK("contract_bwd_w_tc2");
    *P_out = t.P;
    *launched = 1;
    return TGCN_OK;
}

}  // namespace tgcn
