// K1: CSR SpMM recursion over the vertex-major slab, layout transposes, basis / adjoint drivers.
//
// HBM-bound integer/float streaming work: no tensor cores here.  One thread produces one
// float4 of one output row; the dense operand rows are read as coalesced 16-byte vectors
// (a neighbour row is C*4 contiguous bytes), CSR (col,val) pairs are warp-broadcast loads that
// stay in L1, the axpy with T_{k-2} is fused so every slab is written exactly once.
#include <algorithm>
#include <utility>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"

namespace tgcn {

// ------------------------------------------------------------------------------------------------
// swap the two leading axes of in[A,B,D] -> out[B,A,D]
// ------------------------------------------------------------------------------------------------
constexpr int kTrTile = 8;  // TA = TB = 8 chunks of D floats per tile

__global__ void __launch_bounds__(256)
swap_axes_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int B, int D) {
    extern __shared__ float tile[];  // [TA][TB][D]
    const int a0 = blockIdx.y * kTrTile, b0 = blockIdx.x * kTrTile;
    const int na = min(kTrTile, A - a0), nb = min(kTrTile, B - b0);
    const int rowIn = nb * D;   // contiguous run in the source per a
    const int rowOut = na * D;  // contiguous run in the destination per b
    for (int i = threadIdx.x; i < na * rowIn; i += blockDim.x) {
        const int a = i / rowIn, rem = i - a * rowIn;
        tile[a * (kTrTile * D) + rem] = __ldg(in + ((int64_t)(a0 + a) * B + b0) * D + rem);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb * rowOut; i += blockDim.x) {
        const int b = i / rowOut, rem = i - b * rowOut;
        const int a = rem / D, d = rem - a * D;
        out[((int64_t)(b0 + b) * A + a0) * D + rem] = tile[a * (kTrTile * D) + b * D + d];
    }
}

__global__ void __launch_bounds__(256)
swap_axes_wide_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int B, int D) {
    const int64_t total = (int64_t)A * B * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int64_t ba = i / D;
        const int a = (int)(ba % A);
        const int64_t b = ba / A;
        out[i] = __ldg(in + ((int64_t)a * B + b) * D + d);
    }
}

static int swap_axes(const float* in, float* out, int A, int B, int D, cudaStream_t st) {
    if ((int64_t)A * B * D == 0) return TGCN_OK;
    if (D <= 128) {
        dim3 grid((unsigned)ceil_div(B, kTrTile), (unsigned)ceil_div(A, kTrTile));
        TGCN_SUPPORTED(grid.y <= 65535u, "swap_axes: leading axis %d too large for the tiled kernel", A);
        size_t smem = sizeof(float) * kTrTile * kTrTile * D;
        swap_axes_tiled_kernel<<<grid, 256, smem, st>>>(in, out, A, B, D);
    } else {
        int64_t total = (int64_t)A * B * D;
        int blocks = (int)min64(ceil_div(total, 256), (int64_t)kNumSMs * 16);
        swap_axes_wide_kernel<<<blocks, 256, 0, st>>>(in, out, A, B, D);
    }
    TGCN_LAUNCH_CHECK("swap_axes");
    return TGCN_OK;
}

// ------------------------------------------------------------------------------------------------
// x[Q,N,D] <-> slab[N, Q*Dp] with Dp >= D (zero padding columns): the layout change into the vertex-major slab
// fused with the padding the TMA-fed contraction wants (slab rows of whole 128-byte blocks), so neither a separate
// pad copy nor a slice ever runs.  A block owns TN consecutive vertices: per sample their TN*D inputs are one
// contiguous run, and the TN*Q*Dp slab floats are one contiguous run.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
slab_pack_kernel(const float* __restrict__ x, float* __restrict__ slab, int Q, int N, int D, int Dp, int TN) {
    extern __shared__ float ptile[];                 // [Q][TN][D]
    const int n0 = blockIdx.x * TN, tn = min(TN, N - n0);
    const int run = tn * D;
    if ((run & 3) == 0 && ((TN * D) & 3) == 0 && (((int64_t)N * D) & 3) == 0 && aligned16(x)) {
        const int run4 = run >> 2;                    // every sample's run starts on a 16-byte boundary
        for (int i = threadIdx.x; i < Q * run4; i += blockDim.x) {
            const int q = i / run4, r = i - q * run4;
            reinterpret_cast<float4*>(ptile + q * TN * D)[r] =
                __ldg(reinterpret_cast<const float4*>(x + ((int64_t)q * N + n0) * D) + r);
        }
    } else {
        for (int i = threadIdx.x; i < Q * run; i += blockDim.x) {
            const int q = i / run, r = i - q * run;
            ptile[q * TN * D + r] = __ldg(x + ((int64_t)q * N + n0) * D + r);
        }
    }
    __syncthreads();
    const int row = Q * Dp;
    float* dst = slab + (int64_t)n0 * row;
    if ((Dp & 3) == 0 && aligned16(slab)) {
        const int row4 = row >> 2, dp4 = Dp >> 2;
        for (int i = threadIdx.x; i < tn * row4; i += blockDim.x) {
            const int nl = i / row4, r4 = i - nl * row4;
            const int q = r4 / dp4, d = (r4 - q * dp4) * 4;
            const float* src = ptile + (q * TN + nl) * D + d;
            float4 v;
            v.x = d + 0 < D ? src[0] : 0.f; v.y = d + 1 < D ? src[1] : 0.f;
            v.z = d + 2 < D ? src[2] : 0.f; v.w = d + 3 < D ? src[3] : 0.f;
            reinterpret_cast<float4*>(dst)[i] = v;
        }
    } else {
        for (int i = threadIdx.x; i < tn * row; i += blockDim.x) {
            const int nl = i / row, r = i - nl * row;
            const int q = r / Dp, d = r - q * Dp;
            dst[i] = d < D ? ptile[(q * TN + nl) * D + d] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(256)
slab_unpack_kernel(const float* __restrict__ slab, float* __restrict__ x, int Q, int N, int D, int Dp, int TN) {
    extern __shared__ float ptile[];                 // [TN][Q][Dp]
    const int n0 = blockIdx.x * TN, tn = min(TN, N - n0);
    const int row = Q * Dp;
    const float* src = slab + (int64_t)n0 * row;
    if ((row & 3) == 0 && aligned16(slab)) {
        for (int i = threadIdx.x; i < tn * (row >> 2); i += blockDim.x)
            reinterpret_cast<float4*>(ptile)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int i = threadIdx.x; i < tn * row; i += blockDim.x) ptile[i] = __ldg(src + i);
    }
    __syncthreads();
    const int run = tn * D;
    for (int i = threadIdx.x; i < Q * run; i += blockDim.x) {
        const int q = i / run, r = i - q * run;
        const int nl = r / D, d = r - nl * D;
        x[((int64_t)q * N + n0) * D + r] = ptile[(nl * Q + q) * Dp + d];
    }
}

static int slab_tile_rows(int Q, int Dp) {
    // ~16 KB tiles: several blocks per SM overlap their load and store phases (a 48 KB tile measured 43.6 us on the
    // 40 MB mesh layer-1 input); never fewer than 8 rows while 48 KB allow it, so the per-sample runs stay long
    const int cap = (int)(12288 / ((int64_t)Q * Dp));
    int tn = (int)(4096 / ((int64_t)Q * Dp));
    if (tn < 8) tn = cap < 8 ? cap : 8;
    if (tn > 64) tn = 64;
    return tn < 1 ? 1 : tn;
}

int slab_pack(const float* x, float* slab, int Q, int N, int D, int Dp, cudaStream_t st) {
    if ((int64_t)Q * N * D == 0) return TGCN_OK;
    if (Dp == D && D > 128) return swap_axes(x, slab, Q, N, D, st);
    const int TN = slab_tile_rows(Q, Dp);
    const size_t smem = sizeof(float) * (size_t)Q * TN * Dp;
    TGCN_SUPPORTED(smem <= 48 * 1024, "to_slab: Q=%d D=%d too wide for the tile kernel", Q, Dp);
    slab_pack_kernel<<<(unsigned)ceil_div(N, TN), 256, smem, st>>>(x, slab, Q, N, D, Dp, TN);
    TGCN_LAUNCH_CHECK("slab_pack");
    return TGCN_OK;
}

int slab_unpack(const float* slab, float* x, int Q, int N, int D, int Dp, cudaStream_t st) {
    if ((int64_t)Q * N * D == 0) return TGCN_OK;
    if (Dp == D && D > 128) return swap_axes(slab, x, N, Q, D, st);
    const int TN = slab_tile_rows(Q, Dp);
    const size_t smem = sizeof(float) * (size_t)Q * TN * Dp;
    TGCN_SUPPORTED(smem <= 48 * 1024, "from_slab: Q=%d D=%d too wide for the tile kernel", Q, Dp);
    slab_unpack_kernel<<<(unsigned)ceil_div(N, TN), 256, smem, st>>>(slab, x, Q, N, D, Dp, TN);
    TGCN_LAUNCH_CHECK("slab_unpack");
    return TGCN_OK;
}

// ------------------------------------------------------------------------------------------------
// SpMM step
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& x) { fma4_packed(acc, w, x); }

// V = C/4 vectors per row; one thread = one float4 of one row.  `prev` may alias `out`.
template <bool kHasPrev>
__global__ void __launch_bounds__(256)
spmm_step_vec4_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                      const float* __restrict__ val, int N, const float4* __restrict__ in,
                      const float4* prev, float4* out, int V, float alpha, float beta) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * V) return;
    const int row = (int)(idx / V);
    const int v = (int)(idx - (int64_t)row * V);
    int e = __ldg(rowptr + row);
    const int e1 = __ldg(rowptr + row + 1);
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
    // 4 gathers in flight per thread (MLP); two accumulators keep the FMA chains short
    for (; e + 4 <= e1; e += 4) {
        const int c0 = __ldg(col + e), c1 = __ldg(col + e + 1), c2 = __ldg(col + e + 2), c3 = __ldg(col + e + 3);
        const float w0 = __ldg(val + e), w1 = __ldg(val + e + 1), w2 = __ldg(val + e + 2), w3 = __ldg(val + e + 3);
        const float4 x0 = __ldg(in + (int64_t)c0 * V + v);
        const float4 x1 = __ldg(in + (int64_t)c1 * V + v);
        const float4 x2 = __ldg(in + (int64_t)c2 * V + v);
        const float4 x3 = __ldg(in + (int64_t)c3 * V + v);
        fma4(acc0, w0, x0);
        fma4(acc1, w1, x1);
        fma4(acc0, w2, x2);
        fma4(acc1, w3, x3);
    }
    for (; e < e1; ++e) {
        const int c0 = __ldg(col + e);
        const float w0 = __ldg(val + e);
        fma4(acc0, w0, __ldg(in + (int64_t)c0 * V + v));
    }
    float4 r;
    r.x = alpha * (acc0.x + acc1.x);
    r.y = alpha * (acc0.y + acc1.y);
    r.z = alpha * (acc0.z + acc1.z);
    r.w = alpha * (acc0.w + acc1.w);
    if (kHasPrev) {
        const float4 p = prev[idx];
        r.x = fmaf(beta, p.x, r.x);
        r.y = fmaf(beta, p.y, r.y);
        r.z = fmaf(beta, p.z, r.z);
        r.w = fmaf(beta, p.w, r.w);
    }
    out[idx] = r;
}

// Persistent, software-pipelined variant.  With ~6 stored entries per row (meshes) a thread of the plain
// kernel spends its life in a chain of three dependent round trips (row bounds -> (col,val) -> gathers) and the
// gathers are in flight for only one of them.  Here a thread owns one float4 column of a strided set of rows
// and always has the NEXT batch of (col,val) pairs (same row, or the first batch of its next row) and the next
// row's bounds in flight while the current batch's gathers are outstanding.  Summation order per row is the
// same as spmm_step_vec4_kernel (full batches alternate two accumulators, the tail goes to the first), so the
// results are bit-identical.
template <bool kHasPrev>
__global__ void __launch_bounds__(256)
spmm_step_pipe_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                      const float* __restrict__ val, int N, const float4* __restrict__ in,
                      const float4* prev, float4* out, int V, float alpha, float beta, int rows_per_pass) {
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int slot = (int)(gtid / V);
    const int v = (int)(gtid - (int64_t)slot * V);
    if (slot >= rows_per_pass || slot >= N) return;
    int row = slot;
    int e = __ldg(rowptr + row), e1 = __ldg(rowptr + row + 1);
    int c[4] = {0, 0, 0, 0};
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    int nb = 0;
    auto fetch = [&](int ee, int ee1) {
        nb = min(4, ee1 - ee);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nb) { c[i] = __ldg(col + ee + i); w[i] = __ldg(val + ee + i); }
    };
    fetch(e, e1);
    while (true) {
        const int nrow = row + rows_per_pass;
        const bool has_next = nrow < N;
        int ne = 0, ne1 = 0;
        if (has_next) { ne = __ldg(rowptr + nrow); ne1 = __ldg(rowptr + nrow + 1); }
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        while (true) {
            float4 x[4];
            float cw[4];
            const int cnb = nb;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cw[i] = w[i];
                if (i < cnb) x[i] = __ldg(in + (int64_t)c[i] * V + v);
            }
            e += cnb;
            const bool more = e < e1;
            if (more) fetch(e, e1);
            else if (has_next) fetch(ne, ne1);
            if (cnb == 4) {
                fma4(a0, cw[0], x[0]);
                fma4(a1, cw[1], x[1]);
                fma4(a0, cw[2], x[2]);
                fma4(a1, cw[3], x[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (i < cnb) fma4(a0, cw[i], x[i]);
            }
            if (!more) break;
        }
        float4 r;
        r.x = alpha * (a0.x + a1.x);
        r.y = alpha * (a0.y + a1.y);
        r.z = alpha * (a0.z + a1.z);
        r.w = alpha * (a0.w + a1.w);
        const int64_t idx = (int64_t)row * V + v;
        if (kHasPrev) {
            const float4 p = prev[idx];
            r.x = fmaf(beta, p.x, r.x);
            r.y = fmaf(beta, p.y, r.y);
            r.z = fmaf(beta, p.z, r.z);
            r.w = fmaf(beta, p.w, r.w);
        }
        out[idx] = r;
        if (!has_next) break;
        row = nrow; e = ne; e1 = ne1;
    }
}

// Block-staged CSR variant.  A block owns RB consecutive rows: their (col, val) entries are ONE contiguous run of
// the CSR arrays, copied into shared memory by coalesced loads (col pre-multiplied by the row pitch and paired with
// val, so an entry is one 8-byte shared load instead of two global broadcast loads per lane).  Threads are laid out
// (float4 column, row lane): no per-element division, and per entry the instruction stream is LDS.64 + address +
// LDG.128 + 2 FFMA2 -- less than half of the plain kernel's -- at the same or higher occupancy (MINB blocks per SM).
// Blocks whose entry run exceeds the staging capacity walk the global CSR arrays instead (same order, rare).
// Accumulation order per output element is that of spmm_step_vec4_kernel (entries of complete groups of four
// alternate two accumulators, the tail goes to the first): bit-identical results.
constexpr int kCsmCap = 1536;      // staged entries per block (12 KB)
constexpr int kCsmMaxRows = 64;

template <bool kHasPrev, int MINB>
__global__ void __launch_bounds__(256, MINB)
spmm_step_csm_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ val,
                     int N, const float4* __restrict__ in, const float4* prev, float4* out, int V, int RB,
                     float alpha, float beta) {
    __shared__ int2 s_ent[kCsmCap];
    __shared__ int s_rp[kCsmMaxRows + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
    const int tid = ty * V + tx, nthr = V * TY;
    const int row0 = blockIdx.x * RB;
    const int rows = min(RB, N - row0);
    for (int i = tid; i <= rows; i += nthr) s_rp[i] = __ldg(rowptr + row0 + i);
    __syncthreads();
    const int e_lo = s_rp[0], n_ent = s_rp[rows] - e_lo;
    const bool staged = n_ent <= kCsmCap;
    if (staged) {
        for (int i = tid; i < n_ent; i += nthr)
            s_ent[i] = make_int2(__ldg(col + e_lo + i) * V, __float_as_int(__ldg(val + e_lo + i)));
    }
    __syncthreads();
    const float4* inv = in + tx;
    for (int r = ty; r < rows; r += TY) {
        const int eb = s_rp[r], cnt = s_rp[r + 1] - eb;
        const int64_t idx = (int64_t)(row0 + r) * V + tx;
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kHasPrev) pv = prev[idx];
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        if (staged) {
            const int2* ep = s_ent + (eb - e_lo);
            int j = 0;
            for (; j + 4 <= cnt; j += 4) {
                const int2 a = ep[j], b = ep[j + 1], c = ep[j + 2], d = ep[j + 3];
                const float4 x0 = __ldg(inv + a.x), x1 = __ldg(inv + b.x), x2 = __ldg(inv + c.x), x3 = __ldg(inv + d.x);
                fma4(acc0, __int_as_float(a.y), x0);
                fma4(acc1, __int_as_float(b.y), x1);
                fma4(acc0, __int_as_float(c.y), x2);
                fma4(acc1, __int_as_float(d.y), x3);
            }
            const int rem = cnt - j;                       // 0..3 tail entries, gathered together
            if (rem > 0) {
                const int2 a = ep[j];
                const int2 b = rem > 1 ? ep[j + 1] : a;
                const int2 c = rem > 2 ? ep[j + 2] : a;
                const float4 x0 = __ldg(inv + a.x), x1 = __ldg(inv + b.x), x2 = __ldg(inv + c.x);
                fma4(acc0, __int_as_float(a.y), x0);
                if (rem > 1) fma4(acc0, __int_as_float(b.y), x1);
                if (rem > 2) fma4(acc0, __int_as_float(c.y), x2);
            }
        } else {
            int e = eb;
            const int e1 = eb + cnt, ef = eb + (cnt & ~3);
            for (; e < ef; e += 2) {
                fma4(acc0, __ldg(val + e), __ldg(inv + __ldg(col + e) * V));
                fma4(acc1, __ldg(val + e + 1), __ldg(inv + __ldg(col + e + 1) * V));
            }
            for (; e < e1; ++e) fma4(acc0, __ldg(val + e), __ldg(inv + __ldg(col + e) * V));
        }
        float4 r4;
        r4.x = alpha * (acc0.x + acc1.x);
        r4.y = alpha * (acc0.y + acc1.y);
        r4.z = alpha * (acc0.z + acc1.z);
        r4.w = alpha * (acc0.w + acc1.w);
        if (kHasPrev) {
            r4.x = fmaf(beta, pv.x, r4.x);
            r4.y = fmaf(beta, pv.y, r4.y);
            r4.z = fmaf(beta, pv.z, r4.z);
            r4.w = fmaf(beta, pv.w, r4.w);
        }
        out[idx] = r4;
    }
}

// ---- register-tiled row-tile kernel (SPMM_RTILE) ---------------------------------------------------------------
// ncu on the staged-CSR kernel shows the L1 data pipe at 92 % of its wavefront rate: every (row, entry) pair costs one
// LDG.128 per float4 column no matter where the line comes from, so no LDG- or LDS-based gather can go faster.  This
// kernel cuts the NUMBER of gathers: a thread owns one float4 column of R consecutive rows (4R accumulators in
// registers).  The host plan (tgcn_rowtile_plan_host) lists, per tile of R rows, the DISTINCT source rows in ascending
// order, each with a dense R-vector of coefficients (0 where a row of the tile has no such entry).  A source row is
// loaded ONCE per tile and applied to all R rows: on the strip+Morton-ordered 1M-vertex geometric graph a tile of 8
// rows has 35 distinct sources for 96 entries (2.75x fewer gathers), on the cortical mesh 2.0x.  The accumulators are
// paired over ROWS -- {row 2p, row 2p+1} of one component in one 64-bit register -- so the coefficient pairs come
// straight from memory as the first operand of fma.rn.f32x2 and only the 4 components of x need a duplicating move.
// Summation order per output element: ascending source row, one chain (the other kernels alternate two chains), so
// results agree with them to fp32 rounding (not bit-identical); a zero coefficient multiplies the source value, so a
// non-finite activation reaches every row of a tile that shares the source (finite inputs: exact zeros, no effect).
template <int R>
__device__ __forceinline__ void rtile_apply(unsigned long long (&a)[4][R / 2], const float4& x,
                                            const unsigned long long (&wp)[R / 2]) {
    unsigned long long xx[4];
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx[0]) : "f"(x.x));
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx[1]) : "f"(x.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx[2]) : "f"(x.z));
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx[3]) : "f"(x.w));
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < R / 2; ++p)
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[c][p]) : "l"(wp[p]), "l"(xx[c]));
}

template <int R>
__device__ __forceinline__ void rtile_load_w(const float* w, int64_t s, unsigned long long (&wp)[R / 2]) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(w + s * R);   // shared or global (generic 16-byte loads)
#pragma unroll
    for (int i = 0; i < R / 4; ++i) {
        const ulonglong2 v = p[i];
        wp[2 * i] = v.x;
        wp[2 * i + 1] = v.y;
    }
}

// A block owns TY consecutive tiles, whose plan entries are ONE contiguous run of src[] / w[]: the run is copied into
// shared memory by coalesced loads first (source ids pre-multiplied by the row pitch), so the dependent chain of a
// source is LDS -> LDG.128 instead of LDG -> LDG -> LDG.128 (the first version, which walked the plan in global
// memory, was latency-bound at 565 us on the 1M-vertex graph).  Runs longer than the stage are walked in global memory.
constexpr int kRtCap = 768;        // staged (tile, source) pairs per block: 3 KB of ids + 768*R*4 bytes of coefficients
constexpr int kRtMaxTiles = 32;

template <bool kHasPrev, int R, int MINB>
__global__ void __launch_bounds__(256, MINB)
spmm_step_rtile_kernel(const int* __restrict__ tile_ptr, const int* __restrict__ src, const float* __restrict__ w,
                       int N, const float4* __restrict__ in, const float4* prev, float4* out, int V, int ntiles,
                       float alpha, float beta) {
    __shared__ __align__(16) float s_w[kRtCap * R];
    __shared__ int s_src[kRtCap];
    __shared__ int s_tp[kRtMaxTiles + 1];
    constexpr int U = R == 4 ? 4 : 2;          // sources in flight per thread
    const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
    const int tid = ty * V + tx, nthr = V * TY;
    const int t0 = blockIdx.x * TY;
    const int nt = min(TY, ntiles - t0);
    for (int i = tid; i <= nt; i += nthr) s_tp[i] = __ldg(tile_ptr + t0 + i);
    __syncthreads();
    const int s_lo = s_tp[0], n_src = s_tp[nt] - s_lo;
    const bool staged = n_src <= kRtCap;
    if (staged) {
        for (int i = tid; i < n_src; i += nthr) s_src[i] = __ldg(src + s_lo + i) * V;
        const float4* wg = reinterpret_cast<const float4*>(w + (int64_t)s_lo * R);
        float4* ws = reinterpret_cast<float4*>(s_w);
        for (int i = tid; i < n_src * (R / 4); i += nthr) ws[i] = __ldg(wg + i);
    }
    __syncthreads();
    if (ty >= nt) return;
    unsigned long long a[4][R / 2];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < R / 2; ++p) a[c][p] = 0ull;
    const float4* inv = in + tx;
    const int cnt = s_tp[ty + 1] - s_tp[ty];
    if (staged) {
        const int* sp = s_src + (s_tp[ty] - s_lo);
        const float* wp = s_w + (size_t)(s_tp[ty] - s_lo) * R;
        int j = 0;
        for (; j + U <= cnt; j += U) {
            int c[U];
            unsigned long long wv[U][R / 2];
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) c[u] = sp[j + u];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = __ldg(inv + c[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) rtile_load_w<R>(wp, j + u, wv[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) rtile_apply<R>(a, x[u], wv[u]);
        }
        for (; j < cnt; ++j) {
            unsigned long long wv[R / 2];
            const float4 x = __ldg(inv + sp[j]);
            rtile_load_w<R>(wp, j, wv);
            rtile_apply<R>(a, x, wv);
        }
    } else {
        for (int s = s_tp[ty]; s < s_tp[ty] + cnt; ++s) {
            unsigned long long wv[R / 2];
            const float4 x = __ldg(inv + (int64_t)__ldg(src + s) * V);
            rtile_load_w<R>(w, s, wv);
            rtile_apply<R>(a, x, wv);
        }
    }
    const int row0 = (t0 + ty) * R;
#pragma unroll
    for (int p = 0; p < R / 2; ++p) {
        float lo[4], hi[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) asm("mov.b64 {%0, %1}, %2;" : "=f"(lo[c]), "=f"(hi[c]) : "l"(a[c][p]));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = row0 + 2 * p + h;
            if (row >= N) break;
            const float* v = h ? hi : lo;
            const int64_t idx = (int64_t)row * V + tx;
            float4 r4 = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
            if (kHasPrev) {
                const float4 pv = prev[idx];
                r4.x = fmaf(beta, pv.x, r4.x);
                r4.y = fmaf(beta, pv.y, r4.y);
                r4.z = fmaf(beta, pv.z, r4.z);
                r4.w = fmaf(beta, pv.w, r4.w);
            }
            out[idx] = r4;
        }
    }
}

// Persistent, plan-prefetching build of the row-tile kernel ("SPMM_RTILE" = 2).  The kernel above pays three dependent
// global round trips per block (tile_ptr -> src/w -> gather) and hides them only through the 4 co-resident blocks: ncu
// shows 39 % warps active, SMs idle 20 % of the launch (ramp + tail of 4.4 waves) and neither L2 nor DRAM above 40 %.
// Here a block loops over its tile groups (g = blockIdx.x, += gridDim.x) and, while it gathers group i, cp.async brings
// the plan of group i+1 (src ids, coefficient vectors) and the tile pointers of group i+2 into the other shared-memory
// buffers, so an iteration costs ONE global round trip (the gather).  Same arithmetic, same order: bit-identical output.
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <bool kHasPrev, int R, int MINB, int U, int BT>
__global__ void __launch_bounds__(BT, MINB)
spmm_step_rtile_pipe_kernel(const int* __restrict__ tile_ptr, const int* __restrict__ src, const float* __restrict__ w,
                            int N, const float4* __restrict__ in, const float4* prev, float4* out, int V, int ntiles,
                            int ngroups, float alpha, float beta) {
    constexpr int CAP = BT == 128 ? 384 : (R == 8 ? 512 : kRtCap);      // two plan buffers within the 48 KB of static shared memory
    __shared__ __align__(16) float s_w[2][CAP * R];
    __shared__ int s_src[2][CAP];
    __shared__ int s_tp[3][kRtMaxTiles + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
    const int tid = ty * V + tx, nthr = V * TY;
    const int stride = gridDim.x;
    int g = blockIdx.x;
    if (g >= ngroups) return;
    // tile pointers of group gg -> s_tp[slot] (entries past the last tile repeat the final offset: empty tiles)
    auto fetch_tp = [&](int gg, int slot) {
        if (gg < ngroups && tid <= TY) cp_async4(&s_tp[slot][tid], tile_ptr + min(gg * TY + tid, ntiles));
    };
    auto fetch_plan = [&](int slot_tp, int buf) {
        const int lo = s_tp[slot_tp][0], n = s_tp[slot_tp][TY] - lo;
        if (n > CAP) return;
        for (int i = tid; i < n; i += nthr) cp_async4(&s_src[buf][i], src + lo + i);
        const float4* wg = reinterpret_cast<const float4*>(w + (int64_t)lo * R);
        float4* ws = reinterpret_cast<float4*>(s_w[buf]);
        for (int i = tid; i < n * (R / 4); i += nthr) cp_async16(ws + i, wg + i);
    };
    // Programmatic dependent launch (the host launches this kernel with the stream-serialization attribute): let the next
    // kernel of the stream -- the next recursion step -- be scheduled as this grid's blocks retire, and fetch this block's
    // first plan (static data) before waiting for the previous step's output: launch latency, ramp and the two dependent
    // plan loads of the prologue overlap the previous step's tail.  Every read of `in` / `prev` and every write of `out`
    // comes after griddepcontrol.wait.  (Both instructions are no-ops under a plain launch.)
    asm volatile("griddepcontrol.launch_dependents;");
    fetch_tp(g, 0);
    fetch_tp(g + stride, 1);
    cp_async_wait_all();
    __syncthreads();
    fetch_plan(0, 0);
    cp_async_wait_all();
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const float4* inv = in + tx;
    for (int it = 0; g < ngroups; g += stride, ++it) {
        const int cur = it % 3, nxt = (it + 1) % 3, buf = it & 1;
        fetch_tp(g + 2 * stride, (it + 2) % 3);
        if (g + stride < ngroups) fetch_plan(nxt, buf ^ 1);
        const int t0 = g * TY;
        if (t0 + ty < ntiles) {
            unsigned long long a[4][R / 2];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int p = 0; p < R / 2; ++p) a[c][p] = 0ull;
            const int s_lo = s_tp[cur][0], my0 = s_tp[cur][ty], cnt = s_tp[cur][ty + 1] - my0;
            if (s_tp[cur][TY] - s_lo <= CAP) {
                const int* sp = s_src[buf] + (my0 - s_lo);
                const float* wp = s_w[buf] + (size_t)(my0 - s_lo) * R;
                int j = 0;
                for (; j + U <= cnt; j += U) {
                    int c[U];
                    float4 x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) c[u] = sp[j + u] * V;
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u] = __ldg(inv + c[u]);
#pragma unroll
                    for (int u = 0; u < U; ++u) {          // coefficients from shared memory just in time
                        unsigned long long wv[R / 2];
                        rtile_load_w<R>(wp, j + u, wv);
                        rtile_apply<R>(a, x[u], wv);
                    }
                }
                for (; j < cnt; ++j) {
                    unsigned long long wv[R / 2];
                    const float4 x = __ldg(inv + sp[j] * V);
                    rtile_load_w<R>(wp, j, wv);
                    rtile_apply<R>(a, x, wv);
                }
            } else {
                for (int s = my0; s < my0 + cnt; ++s) {
                    unsigned long long wv[R / 2];
                    const float4 x = __ldg(inv + (int64_t)__ldg(src + s) * V);
                    rtile_load_w<R>(w, s, wv);
                    rtile_apply<R>(a, x, wv);
                }
            }
            const int row0 = (t0 + ty) * R;
#pragma unroll
            for (int p = 0; p < R / 2; ++p) {
                float lo[4], hi[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) asm("mov.b64 {%0, %1}, %2;" : "=f"(lo[c]), "=f"(hi[c]) : "l"(a[c][p]));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = row0 + 2 * p + h;
                    if (row >= N) break;
                    const float* v = h ? hi : lo;
                    const int64_t idx = (int64_t)row * V + tx;
                    float4 r4 = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
                    if (kHasPrev) {
                        const float4 pv = prev[idx];
                        r4.x = fmaf(beta, pv.x, r4.x);
                        r4.y = fmaf(beta, pv.y, r4.y);
                        r4.z = fmaf(beta, pv.z, r4.z);
                        r4.w = fmaf(beta, pv.w, r4.w);
                    }
                    out[idx] = r4;
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }
}

// Launch with programmatic stream serialization (see the kernel's prologue); TGCN_SPMM_PDL=0 falls back to a plain launch.
template <typename... KArgs, typename... Args>
static void launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    static const bool pdl = [] { const char* e = getenv("TGCN_SPMM_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface in TGCN_LAUNCH_CHECK
}

// ---- row-tile plan registry: plans are created by the host side once per CSR operand and looked up by (device,
// device address of the operand's `col` array, N).  The owner (csr.LaplacianCSR / parallel.RowPartitionedLayer) keeps the
// `col` tensor alive for as long as the plan is registered and destroys the plan before releasing it, so an address can
// never be matched after its allocation was recycled.  The lookup is lock-free while no plan exists.
struct RowTilePlan { const void* key; int device; const int* tile_ptr; const int* src; const float* w; int R, N, n_src_rows; bool live; };
static std::mutex g_rt_mu;
static std::vector<RowTilePlan> g_rt_plans;
static std::atomic<int> g_rt_live{0};

static bool find_rowtile_plan(const void* col, int N, RowTilePlan* out) {
    if (g_rt_live.load(std::memory_order_acquire) == 0) return false;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    std::lock_guard<std::mutex> lk(g_rt_mu);
    for (const RowTilePlan& rp : g_rt_plans)
        if (rp.live && rp.key == col && rp.N == N && rp.device == dev) { *out = rp; return true; }
    return false;
}

template <bool kHasPrev>
__global__ void __launch_bounds__(256)
spmm_step_scalar_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                        const float* __restrict__ val, int N, const float* __restrict__ in,
                        const float* prev, float* out, int64_t C, float alpha, float beta) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * C) return;
    const int row = (int)(idx / C);
    const int64_t c = idx - (int64_t)row * C;
    int e = __ldg(rowptr + row);
    const int e1 = __ldg(rowptr + row + 1);
    float acc = 0.f;
    for (; e < e1; ++e) acc = fmaf(__ldg(val + e), __ldg(in + (int64_t)__ldg(col + e) * C + c), acc);
    float r = alpha * acc;
    if (kHasPrev) r = fmaf(beta, prev[idx], r);
    out[idx] = r;
}

static int spmm_step(const int* rowptr, const int* col, const float* val, int N, const float* in,
                     const float* prev, float* out, int64_t C, float alpha, float beta,
                     cudaStream_t st) {
    if (N == 0 || C == 0) return TGCN_OK;
    TGCN_REQUIRE(in != out, "spmm_step: `in` must not alias `out`");
    const bool vec = (C % 4 == 0) && aligned16(in) && aligned16(out) && (prev == nullptr || aligned16(prev));
    // register-tiled row-tile kernels: whenever a row-tile plan is registered for this operand ("SPMM_RTILE": 0 = off,
    // 1 = one-shot kernel, 2/3 = persistent kernel, 4 = one-shot kernel with 4 blocks per SM also for 8-row tiles)
    RowTilePlan rt;
    const int rt_mode = tuning_value(kTuneSpmmRtile);
    if (vec && rt_mode != 0 && C / 4 <= 256 && find_rowtile_plan(col, N, &rt) &&
        (int64_t)rt.n_src_rows * (C / 4) < ((int64_t)1 << 31)) {             // src * V stays in int32
        const int V = (int)(C / 4);
        int TY = 256 / V;
        if (TY > kRtMaxTiles) TY = kRtMaxTiles;
        const int ntiles = (int)ceil_div(N, rt.R);
        const unsigned blocks = (unsigned)ceil_div(ntiles, TY);
        const dim3 bd((unsigned)V, (unsigned)TY);
#define TGCN_SPMM_RT(RR, MB)                                                                                        \
        do {                                                                                                        \
            if (prev) spmm_step_rtile_kernel<true, RR, MB><<<blocks, bd, 0, st>>>(rt.tile_ptr, rt.src, rt.w, N, (const float4*)in, \
                                                                                  (const float4*)prev, (float4*)out, V, ntiles, alpha, beta); \
            else spmm_step_rtile_kernel<false, RR, MB><<<blocks, bd, 0, st>>>(rt.tile_ptr, rt.src, rt.w, N, (const float4*)in, nullptr, \
                                                                              (float4*)out, V, ntiles, alpha, beta); \
        } while (0)
        // 2 (the default): persistent, plan-prefetching build for R = 4 (mesh layer 1: 23.1 -> 20.2 us per step inside
        // the training step, 1M-vertex geometric graph 529 -> 485 us; R = 8 measured slower that way and keeps the
        // one-shot kernel) with 8 gathers in flight per thread; 3: persistent for both R, 128-thread blocks at R = 4
        if ((rt_mode == 2 && rt.R == 4) || rt_mode == 3) {
            // mode 3 at R = 4: blocks of 128 threads, 8 per SM -- the per-iteration block barrier spans 4 warps, not 8
            const bool small = rt_mode == 3 && rt.R == 4 && V <= 128;
            const int TYp = small ? 128 / V : TY;
            const int per_sm = rt.R == 8 ? 3 : (small ? 8 : 4);
            const int ngroups = (int)ceil_div(ntiles, TYp);
            const unsigned pgrid = (unsigned)min64(ngroups, (int64_t)kNumSMs * per_sm);
            const dim3 bdp((unsigned)V, (unsigned)TYp);
#define TGCN_SPMM_RTP(RR, MB, UU, BB)                                                                                   \
            do {                                                                                                    \
                if (prev) launch_dependent(spmm_step_rtile_pipe_kernel<true, RR, MB, UU, BB>, dim3(pgrid), bdp, st, rt.tile_ptr, rt.src, rt.w, N, \
                                           (const float4*)in, (const float4*)prev, (float4*)out, V, ntiles, ngroups, alpha, beta); \
                else launch_dependent(spmm_step_rtile_pipe_kernel<false, RR, MB, UU, BB>, dim3(pgrid), bdp, st, rt.tile_ptr, rt.src, rt.w, N, \
                                      (const float4*)in, (const float4*)nullptr, (float4*)out, V, ntiles, ngroups, alpha, beta); \
            } while (0)
            if (rt.R == 8) TGCN_SPMM_RTP(8, 3, 2, 256); else if (small) TGCN_SPMM_RTP(4, 8, 8, 128); else TGCN_SPMM_RTP(4, 4, 8, 256);
#undef TGCN_SPMM_RTP
            TGCN_LAUNCH_CHECK("spmm_step");
            return TGCN_OK;
        }
        // one-shot kernel (the 5 / 6 / 8 blocks-per-SM builds of round 1 measured no faster than 4 and were dropped)
        if (rt.R == 8) {
            if (rt_mode == 4) TGCN_SPMM_RT(8, 4); else TGCN_SPMM_RT(8, 3);
        } else {
            TGCN_SPMM_RT(4, 4);
        }
#undef TGCN_SPMM_RT
        TGCN_LAUNCH_CHECK("spmm_step");
        return TGCN_OK;
    }
    // staged-CSR kernel: units digit = blocks per SM it is compiled for (4, 5, 6, 8; 0 = off), tens digit = rows per
    // thread (0 = 4); 100 + that (the default, 106)
    // = only for slabs that do not stay in L2 (>= 96 MB), where it measured 10 % faster than the other variants
    int csm_b = tuning_value(kTuneSpmmCsm);
    if (csm_b >= 100) csm_b = ((int64_t)N * C * 4 >= (int64_t)96 << 20) ? csm_b - 100 : 0;
    const int csm_p = csm_b / 10 > 0 ? csm_b / 10 : 4;      // tens digit: rows per thread (default 4)
    csm_b %= 10;
    if (vec && csm_b > 0 && C / 4 <= 256 && (int64_t)N * (C / 4) < ((int64_t)1 << 30)) {   // col * V stays in int32 (halo rows included)
        const int V = (int)(C / 4);
        int TY = 256 / V;
        if (TY > kCsmMaxRows) TY = kCsmMaxRows;
        int RB = csm_p * TY;                               // rows per thread: the staging cost is paid once per csm_p rows
        if (RB > kCsmMaxRows) RB = kCsmMaxRows;
        const unsigned blocks = (unsigned)ceil_div(N, RB);
        const dim3 bd((unsigned)V, (unsigned)TY);
#define TGCN_SPMM_CSM(MB)                                                                                          \
        do {                                                                                                       \
            if (prev) spmm_step_csm_kernel<true, MB><<<blocks, bd, 0, st>>>(rowptr, col, val, N, (const float4*)in, \
                                                                             (const float4*)prev, (float4*)out, V, RB, alpha, beta); \
            else spmm_step_csm_kernel<false, MB><<<blocks, bd, 0, st>>>(rowptr, col, val, N, (const float4*)in, nullptr, \
                                                                         (float4*)out, V, RB, alpha, beta);          \
        } while (0)
        if (csm_b >= 8) TGCN_SPMM_CSM(8); else if (csm_b >= 6) TGCN_SPMM_CSM(6); else if (csm_b >= 5) TGCN_SPMM_CSM(5); else TGCN_SPMM_CSM(4);
#undef TGCN_SPMM_CSM
        TGCN_LAUNCH_CHECK("spmm_step");
        return TGCN_OK;
    }
    const int pipe_bps = tuning_value(kTuneSpmmPipe);      // blocks per SM of the persistent kernel (0 = off)
    if (vec && pipe_bps > 0 && C / 4 <= 4096) {
        const int V = (int)(C / 4);
        int64_t blocks = (int64_t)kNumSMs * pipe_bps;
        const int64_t need = ceil_div((int64_t)N * V, 256);
        if (blocks > need) blocks = need;
        const int rows_per_pass = (int)((blocks * 256) / V);
        if (rows_per_pass >= 1) {
            if (prev) spmm_step_pipe_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, val, N, (const float4*)in,
                                                                                   (const float4*)prev, (float4*)out, V, alpha, beta, rows_per_pass);
            else spmm_step_pipe_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, val, N, (const float4*)in, nullptr,
                                                                                (float4*)out, V, alpha, beta, rows_per_pass);
            TGCN_LAUNCH_CHECK("spmm_step");
            return TGCN_OK;
        }
    }
    if (vec) {
        const int V = (int)(C / 4);
        const int64_t total = (int64_t)N * V;
        const unsigned blocks = (unsigned)ceil_div(total, 256);
        if (prev)
            spmm_step_vec4_kernel<true><<<blocks, 256, 0, st>>>(rowptr, col, val, N, (const float4*)in,
                                                                (const float4*)prev, (float4*)out, V, alpha, beta);
        else
            spmm_step_vec4_kernel<false><<<blocks, 256, 0, st>>>(rowptr, col, val, N, (const float4*)in, nullptr,
                                                                 (float4*)out, V, alpha, beta);
    } else {
        const int64_t total = (int64_t)N * C;
        TGCN_SUPPORTED(ceil_div(total, 256) < (int64_t)INT32_MAX, "spmm_step: slab too large for the scalar path");
        const unsigned blocks = (unsigned)ceil_div(total, 256);
        if (prev)
            spmm_step_scalar_kernel<true><<<blocks, 256, 0, st>>>(rowptr, col, val, N, in, prev, out, C, alpha, beta);
        else
            spmm_step_scalar_kernel<false><<<blocks, 256, 0, st>>>(rowptr, col, val, N, in, nullptr, out, C, alpha, beta);
    }
    TGCN_LAUNCH_CHECK("spmm_step");
    return TGCN_OK;
}

// y -= x (textbook-recursion adjoint only)
__global__ void __launch_bounds__(256) sub_inplace_kernel(float* y, const float* __restrict__ x, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] -= __ldg(x + i);
}

// Xt[k,q,n,d] in the reference layout from the internal stack; reproduces the reference's own
// fp32 operation `2 * X - Xt[k-2]` (gcn.py:153) element by element.
__global__ void __launch_bounds__(256)
basis_to_reference_kernel(const float* __restrict__ stack, float* __restrict__ Xt, int Q, int N, int D, int K,
                          int recursion) {
    const int64_t per = (int64_t)Q * N * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int64_t qn = i / D;
        const int n = (int)(qn % N);
        const int q = (int)(qn / N);
        const int64_t src = ((int64_t)n * Q + q) * D + d;
        float even = 0.f, odd = 0.f;
        for (int k = 0; k < K; ++k) {
            const float p = __ldg(stack + (int64_t)k * per + src);
            float r;
            if (recursion == TGCN_RECURSION_CHEBYSHEV || k < 2) {
                r = p;
            } else {
                r = 2.f * p - ((k & 1) ? odd : even);
            }
            if (k & 1) odd = r; else even = r;
            Xt[(int64_t)k * per + i] = r;
        }
    }
}

}  // namespace tgcn

using namespace tgcn;

// Host-side row-tile plan (include/tgcn_b200.h): per tile of R consecutive rows the DISTINCT source rows in ascending order
// (src_host) and, per source, the R coefficients it has in the rows of the tile (w_host, 0 where a row has no such entry;
// duplicates within a row are summed).  tile_ptr_host[ceil(N / R) + 1].  src_host == w_host == NULL: size query.  pad > 1
// rounds every non-empty tile up to a multiple of `pad` entries with zero-coefficient repeats of its last source.  Returns
// the number of (tile, source) pairs, or -1 on bad arguments.
extern "C" int64_t tgcn_rowtile_plan_host(const int32_t* rowptr_host, const int32_t* col_host, const float* val_host, int N,
                                          int R, int pad, int32_t* tile_ptr_host, int32_t* src_host, float* w_host) {
    if (N < 0 || (R != 4 && R != 8) || pad < 1 || pad > 16 || !rowptr_host || (rowptr_host[N] > 0 && (!col_host || !val_host))) return -1;
    if ((src_host == nullptr) != (w_host == nullptr)) return -1;
    const int nt = (int)ceil_div(N, R);
    int64_t total = 0;
    std::vector<std::pair<int32_t, int32_t>> ent;     // (col, CSR position) of the tile's entries
    for (int t = 0; t < nt; ++t) {
        const int r0 = t * R, r1 = (int)min64((int64_t)N, (int64_t)r0 + R);
        const int e0 = rowptr_host[r0], e1 = rowptr_host[r1];
        ent.clear();
        for (int e = e0; e < e1; ++e) ent.push_back({col_host[e], e});
        std::sort(ent.begin(), ent.end());
        if (tile_ptr_host) tile_ptr_host[t] = (int32_t)total;
        const int64_t tile_start = total;
        int r = r0;                                    // row lookup for CSR positions: positions ascend within a source
        for (size_t i = 0; i < ent.size();) {
            size_t j = i;
            while (j < ent.size() && ent[j].first == ent[i].first) ++j;
            if (src_host) {
                src_host[total] = ent[i].first;
                float* wv = w_host + total * R;
                for (int q = 0; q < R; ++q) wv[q] = 0.f;
                r = r0;
                for (size_t k = i; k < j; ++k) {
                    const int e = ent[k].second;
                    while (rowptr_host[r + 1] <= e) ++r;
                    wv[r - r0] += val_host[e];
                }
            }
            ++total;
            i = j;
        }
        // pad the tile's run to a multiple of `pad` with zero-coefficient repeats of its last source (an L1 hit for the
        // kernel), so that a kernel that keeps `pad` gathers in flight never runs its one-at-a-time tail loop
        if (!ent.empty())
            for (; (total - tile_start) % pad != 0; ++total) {
                if (src_host) {
                    src_host[total] = ent.back().first;
                    for (int q = 0; q < R; ++q) w_host[total * R + q] = 0.f;
                }
            }
    }
    if (tile_ptr_host) tile_ptr_host[nt] = (int32_t)total;
    return total;
}

// Register / drop a row-tile plan for the CSR operand whose column array lives at device address `col_dev`
// (device arrays owned by the caller; they must outlive the plan); n_src_rows = rows of the gathered operand
// (every source id is < n_src_rows; N for a square operand, more with halo rows).  Returns a handle >= 0.
extern "C" int64_t tgcn_rowtile_plan_create(const int32_t* col_dev, int N, int n_src_rows, int R, const int32_t* tile_ptr_dev,
                                            const int32_t* src_dev, const float* w_dev) {
    if (!col_dev || !tile_ptr_dev || !src_dev || !w_dev || (R != 4 && R != 8) || N < 1 || n_src_rows < 1 || !aligned16(w_dev)) {
        set_error(TGCN_ERR_INVALID, "tgcn_rowtile_plan_create: bad arguments");
        return -1;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }     // the plan belongs to the current device
    std::lock_guard<std::mutex> lk(g_rt_mu);
    const RowTilePlan np{col_dev, dev, tile_ptr_dev, src_dev, w_dev, R, N, n_src_rows, true};
    g_rt_live.fetch_add(1, std::memory_order_release);
    for (size_t i = 0; i < g_rt_plans.size(); ++i)
        if (!g_rt_plans[i].live) { g_rt_plans[i] = np; return (int64_t)i; }
    g_rt_plans.push_back(np);
    return (int64_t)g_rt_plans.size() - 1;
}

extern "C" int tgcn_rowtile_plan_destroy(int64_t handle) {
    std::lock_guard<std::mutex> lk(g_rt_mu);
    if (handle < 0 || handle >= (int64_t)g_rt_plans.size() || !g_rt_plans[handle].live)
        return set_error(TGCN_ERR_INVALID, "tgcn_rowtile_plan_destroy: unknown handle %lld", (long long)handle);
    g_rt_plans[handle].live = false;
    g_rt_live.fetch_sub(1, std::memory_order_release);
    return TGCN_OK;
}

extern "C" int tgcn_to_slab(const float* x, float* slab, int Q, int N, int D, void* stream) {
    TGCN_REQUIRE(x && slab, "tgcn_to_slab: null pointer");
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0, "tgcn_to_slab: negative size");
    return slab_pack(x, slab, Q, N, D, D, as_stream(stream));
}

extern "C" int tgcn_from_slab(const float* slab, float* x, int Q, int N, int D, void* stream) {
    TGCN_REQUIRE(x && slab, "tgcn_from_slab: null pointer");
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0, "tgcn_from_slab: negative size");
    return slab_unpack(slab, x, Q, N, D, D, as_stream(stream));
}

extern "C" int tgcn_spmm_step(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                              const float* in, const float* prev, float* out, int64_t C,
                              float alpha, float beta, void* stream) {
    TGCN_REQUIRE(N >= 0 && C >= 0, "tgcn_spmm_step: negative size");
    if (N == 0 || C == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptr && in && out, "tgcn_spmm_step: null pointer");
    return spmm_step(rowptr, col, val, N, in, prev, out, C, alpha, beta, as_stream(stream));
}

namespace tgcn {
// basis of x[Q,N,D] in slabs of Q*Dp columns (Dp >= D: zero padding columns, which every recursion step keeps zero)
int cheb_basis_padded(const int32_t* rowptr, const int32_t* col, const float* val, int N, const float* x, float* stack,
                      int Q, int D, int Dp, int K, int recursion, cudaStream_t st) {
    const int64_t C = (int64_t)Q * Dp;
    const int64_t S = (int64_t)N * C;
    if (S == 0 || D == 0) return TGCN_OK;
    TGCN_PROPAGATE(slab_pack(x, stack, Q, N, D, Dp, st));
    for (int k = 1; k < K; ++k) {
        float* cur = stack + (int64_t)k * S;
        const float* in = stack + (int64_t)(k - 1) * S;
        if (recursion == TGCN_RECURSION_REFERENCE || k == 1) {
            TGCN_PROPAGATE(spmm_step(rowptr, col, val, N, in, nullptr, cur, C, 1.f, 0.f, st));
        } else {
            TGCN_PROPAGATE(spmm_step(rowptr, col, val, N, in, stack + (int64_t)(k - 2) * S, cur, C, 2.f, -1.f, st));
        }
    }
    return TGCN_OK;
}
}  // namespace tgcn

extern "C" int tgcn_cheb_basis(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                               const float* x, float* stack, int Q, int D, int K, int recursion,
                               void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0 && K >= 1, "tgcn_cheb_basis: bad sizes Q=%d N=%d D=%d K=%d", Q, N, D, K);
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE || recursion == TGCN_RECURSION_CHEBYSHEV,
                 "tgcn_cheb_basis: unknown recursion %d", recursion);
    if ((int64_t)Q * N * D == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptr && x && stack, "tgcn_cheb_basis: null pointer");
    return cheb_basis_padded(rowptr, col, val, N, x, stack, Q, D, D, K, recursion, as_stream(stream));
}

extern "C" int tgcn_basis_to_reference(const float* stack, float* Xt, int Q, int N, int D, int K,
                                       int recursion, void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0 && K >= 1, "tgcn_basis_to_reference: bad sizes");
    const int64_t per = (int64_t)Q * N * D;
    if (per == 0) return TGCN_OK;
    TGCN_REQUIRE(stack && Xt, "tgcn_basis_to_reference: null pointer");
    const int blocks = (int)min64(ceil_div(per, 256), (int64_t)kNumSMs * 16);
    basis_to_reference_kernel<<<blocks, 256, 0, as_stream(stream)>>>(stack, Xt, Q, N, D, K, recursion);
    TGCN_LAUNCH_CHECK("basis_to_reference");
    return TGCN_OK;
}

namespace tgcn {
int cheb_adjoint_padded(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N, float* gstack, float* dx,
                        int Q, int D, int Dp, int K, int recursion, cudaStream_t st) {
    const int64_t C = (int64_t)Q * Dp;
    const int64_t S = (int64_t)N * C;
    if (S == 0 || D == 0) return TGCN_OK;
    if (recursion == TGCN_RECURSION_REFERENCE) {
        // P_j = L P_{j-1}:  A_{K-1} = G_{K-1};  A_j = G_j + L^T A_{j+1}   (Horner), dx = A_0
        for (int j = K - 2; j >= 0; --j) {
            float* a = gstack + (int64_t)j * S;
            TGCN_PROPAGATE(spmm_step(rowptrT, colT, valT, N, gstack + (int64_t)(j + 1) * S, a, a, C, 1.f, 1.f, st));
        }
    } else {
        // T_k = 2 L T_{k-1} - T_{k-2}:  A_{k-1} += 2 L^T A_k;  A_{k-2} -= A_k;  finally A_0 += L^T A_1
        const int blocks = (int)min64(ceil_div(S, 256), (int64_t)kNumSMs * 16);
        for (int k = K - 1; k >= 2; --k) {
            float* ak = gstack + (int64_t)k * S;
            float* a1 = gstack + (int64_t)(k - 1) * S;
            float* a2 = gstack + (int64_t)(k - 2) * S;
            TGCN_PROPAGATE(spmm_step(rowptrT, colT, valT, N, ak, a1, a1, C, 2.f, 1.f, st));
            sub_inplace_kernel<<<blocks, 256, 0, st>>>(a2, ak, S);
            TGCN_LAUNCH_CHECK("sub_inplace");
        }
        if (K >= 2) TGCN_PROPAGATE(spmm_step(rowptrT, colT, valT, N, gstack + S, gstack, gstack, C, 1.f, 1.f, st));
    }
    return slab_unpack(gstack, dx, Q, N, D, Dp, st);
}
}  // namespace tgcn

extern "C" int tgcn_cheb_adjoint(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N,
                                 float* gstack, float* dx, int Q, int D, int K, int recursion,
                                 void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0 && K >= 1, "tgcn_cheb_adjoint: bad sizes");
    if ((int64_t)Q * N * D == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptrT && gstack && dx, "tgcn_cheb_adjoint: null pointer");
    return cheb_adjoint_padded(rowptrT, colT, valT, N, gstack, dx, Q, D, D, K, recursion, as_stream(stream));
}
