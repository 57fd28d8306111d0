// K1: CSR SpMM recursion over the vertex-major slab, layout transposes, basis / adjoint drivers.
//
// HBM-bound integer/float streaming work: no tensor cores here.  One thread produces one
// float4 of one output row; the dense operand rows are read as coalesced 16-byte vectors
// (a neighbour row is C*4 contiguous bytes), CSR (col,val) pairs are warp-broadcast loads that
// stay in L1, the axpy with T_{k-2} is fused so every slab is written exactly once.
#include "common.cuh"

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
// SpMM step
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& x) {
    acc.x = fmaf(w, x.x, acc.x);
    acc.y = fmaf(w, x.y, acc.y);
    acc.z = fmaf(w, x.z, acc.z);
    acc.w = fmaf(w, x.w, acc.w);
}

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

extern "C" int tgcn_to_slab(const float* x, float* slab, int Q, int N, int D, void* stream) {
    TGCN_REQUIRE(x && slab, "tgcn_to_slab: null pointer");
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0, "tgcn_to_slab: negative size");
    return swap_axes(x, slab, Q, N, D, as_stream(stream));
}

extern "C" int tgcn_from_slab(const float* slab, float* x, int Q, int N, int D, void* stream) {
    TGCN_REQUIRE(x && slab, "tgcn_from_slab: null pointer");
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0, "tgcn_from_slab: negative size");
    return swap_axes(slab, x, N, Q, D, as_stream(stream));
}

extern "C" int tgcn_spmm_step(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                              const float* in, const float* prev, float* out, int64_t C,
                              float alpha, float beta, void* stream) {
    TGCN_REQUIRE(N >= 0 && C >= 0, "tgcn_spmm_step: negative size");
    if (N == 0 || C == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptr && in && out, "tgcn_spmm_step: null pointer");
    return spmm_step(rowptr, col, val, N, in, prev, out, C, alpha, beta, as_stream(stream));
}

extern "C" int tgcn_cheb_basis(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                               const float* x, float* stack, int Q, int D, int K, int recursion,
                               void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0 && K >= 1, "tgcn_cheb_basis: bad sizes Q=%d N=%d D=%d K=%d", Q, N, D, K);
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE || recursion == TGCN_RECURSION_CHEBYSHEV,
                 "tgcn_cheb_basis: unknown recursion %d", recursion);
    const int64_t C = (int64_t)Q * D;
    const int64_t S = (int64_t)N * C;
    if (S == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptr && x && stack, "tgcn_cheb_basis: null pointer");
    cudaStream_t st = as_stream(stream);
    TGCN_PROPAGATE(swap_axes(x, stack, Q, N, D, st));
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

extern "C" int tgcn_cheb_adjoint(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N,
                                 float* gstack, float* dx, int Q, int D, int K, int recursion,
                                 void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 0 && K >= 1, "tgcn_cheb_adjoint: bad sizes");
    const int64_t C = (int64_t)Q * D;
    const int64_t S = (int64_t)N * C;
    if (S == 0) return TGCN_OK;
    TGCN_REQUIRE(rowptrT && gstack && dx, "tgcn_cheb_adjoint: null pointer");
    cudaStream_t st = as_stream(stream);
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
    return swap_axes(gstack, dx, N, Q, D, st);
}
