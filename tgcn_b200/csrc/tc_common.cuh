// tcgen05 / TMEM / mbarrier building blocks for the tensor-core contraction engine (sm_100a).
//
// Operand convention used throughout contract_tc.cu:
//   * fp32 data is split on the fly into tf32 "hi" (round-to-nearest) and "lo" = x - hi, and a
//     product is accumulated as hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator (3xTF32),
//     which keeps the contraction within ~2^-21 relative of exact fp32 -- plain TF32 (10-bit
//     mantissa) would not meet the 1e-4 parity bar.
//   * shared-memory operand tiles use the canonical SWIZZLE_128B layouts:
//       K-major : rows = M/N index, 32 fp32 (128 B) of the reduction index per row,
//                 8-row groups 1024 B apart (SBO = 1024);
//                 the 16-byte chunk index inside a row is XORed with (row & 7);
//       MN-major: rows = reduction index, 32 fp32 (128 B) of the M/N index per row.  For 32-bit
//                 (tf32) operands the only MN-major layout the tensor core accepts is
//                 SWIZZLE_128B_BASE32B: 4-row atoms (512 B, SBO apart), the 32-byte chunk index
//                 inside a row XORed with (row & 3); 32-wide M/N blocks LBO bytes apart.
//     Tiles are 1024-byte aligned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tgcn {
namespace tc {

constexpr uint32_t kRowBytes = 128;      // one swizzle row: 32 fp32
constexpr uint32_t kAtomBytes = 1024;    // 8 rows
constexpr unsigned long long kSpinLimit = 4000000000ull;  // ~2 s at 2 GHz: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (row, col) inside a [rows x 32 fp32] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
    return row * kRowBytes + ((((col >> 2) ^ (row & 7u)) << 4) | ((col & 3u) << 2));
}

// byte offset of element (row = reduction index, col = M/N index within a 32-wide block) inside a
// SWIZZLE_128B_BASE32B MN-major block (rows 128 B apart)
__device__ __forceinline__ uint32_t sw128b32_offset(uint32_t row, uint32_t col) {
    return row * kRowBytes + ((((col >> 3) ^ (row & 3u)) << 5) | ((col & 7u) << 2));
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const unsigned long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimit) __trap();   // never hang the device
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion (bytes) signalled on `bar`; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- proxies / fences ---------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; the base address lands in *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
    return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u;
}

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B
//   bits [0,14) start address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1,
//   [61,64) layout type (2 = SWIZZLE_128B)
//   layout type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// K-major tile: LBO is not used by the swizzled K-major layouts (canonical value 1), SBO = 1024
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr) {
    return make_smem_desc(saddr, 16, kAtomBytes, 2);
}
// MN-major tf32 tile (SWIZZLE_128B_BASE32B): 4-row atoms 512 B apart along the reduction index,
// 32-wide M/N blocks `block_stride_bytes` apart
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t saddr, uint32_t block_stride_bytes) {
    return make_smem_desc(saddr, block_stride_bytes, 512, 1);
}

// instruction descriptor, kind::tf32, fp32 accumulate
//   [4,6) c_format = 1 (F32), [7,10) a_format = 2 (TF32), [10,13) b_format = 2, [15] a_major, [16] b_major
//   (0 = K-major, 1 = MN-major), [17,23) N >> 3, [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
           ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

// hi/lo split of one fp32 into two swizzled tiles
__device__ __forceinline__ void store_split(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t off, float x) {
    const float h = tf32_hi(x);
    *reinterpret_cast<float*>(tile_hi + off) = h;
    *reinterpret_cast<float*>(tile_lo + off) = tf32_hi(x - h);   // rounded, not left to the tensor core to truncate
}

}  // namespace tc
}  // namespace tgcn
