// K2 / K4 (fp32 CUDA-core engine): weight contraction forward, its two backward GEMMs, the
// weight mixing that maps the reference recursion onto the power basis, and the bias gradient.
//
// The tcgen05 engine lives in contract_tc.cu; this file is the exact-fp32 engine
// (TGCN_ENGINE_FFMA) that also serves shapes the tensor-core tiles do not cover.
#include "common.cuh"

namespace tgcn {

// ------------------------------------------------------------------------------------------------
// Generic strided small GEMM:  C[z][m][c] = sum_kk A[z][m][kk] * B[z][kk][c]  (+ bias)
//   rows m are (vertex, sample) pairs, m = n*Q + q.
//   A/C rows are addressed either as slab rows (m * ld) or as API rows ((q*N + n) * ld).
//   A columns kk may span several slabs: offset = (kk / a_split) * a_split_stride + kk % a_split.
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    const float* A;
    int a_api_rows;          // 0: row offset m*lda, 1: ((m%Q)*N + m/Q)*lda
    int64_t lda;
    int a_split;             // columns per slab (>= Kred when A is a single matrix)
    int64_t a_split_stride;  // elements between slabs
    int64_t a_batch_stride;
    const float* B;
    int64_t ldb_k, ldb_c, b_batch_stride;
    float* C;
    int c_api_rows;
    int64_t ldc, c_batch_stride;
    const float* bias;  // added to C; bias_mode PER_VERTEX: bias[n*Ncol + c], PER_FILTER: bias[c]
    int bias_mode;
    int M, Kred, Ncol, Q, N;
};

constexpr int kGemmTM = 64;       // rows per block
constexpr int kGemmThreads = 128;
constexpr int kGemmKC = 16;       // reduction chunk staged in shared memory

template <int TX>  // TX column groups of 4 -> 4*TX columns per block; RM = TX/2 rows per thread
__global__ void __launch_bounds__(kGemmThreads)
small_gemm_kernel(const GemmParams p) {
    constexpr int RM = TX / 2;
    constexpr int NC = 4 * TX;
    constexpr int LDS_A = kGemmTM + 4;
    __shared__ __align__(16) float As[kGemmKC][LDS_A];  // transposed: [kk][m]
    __shared__ __align__(16) float Bs[kGemmKC][NC];
    __shared__ int64_t rowA[kGemmTM];
    __shared__ int64_t rowC[kGemmTM];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.x * kGemmTM;
    const int c0 = blockIdx.y * NC;
    const int z = blockIdx.z;
    const float* __restrict__ A = p.A + (int64_t)z * p.a_batch_stride;
    const float* __restrict__ B = p.B + (int64_t)z * p.b_batch_stride;
    float* C = p.C + (int64_t)z * p.c_batch_stride;

    if (tid < kGemmTM) {
        const int m = m0 + tid;
        int64_t ra = -1, rc = -1;
        if (m < p.M) {
            const int n = m / p.Q, q = m - n * p.Q;
            const int64_t api = (int64_t)q * p.N + n;
            ra = (p.a_api_rows ? api : (int64_t)m) * p.lda;
            rc = (p.c_api_rows ? api : (int64_t)m) * p.ldc;
        }
        rowA[tid] = ra;
        rowC[tid] = rc;
    }
    __syncthreads();

    float acc[RM][4];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    for (int k0 = 0; k0 < p.Kred; k0 += kGemmKC) {
        const int kc = min(kGemmKC, p.Kred - k0);
        // A chunk -> As[kk][m]; consecutive threads walk kk first (contiguous in global for slab rows)
        for (int i = tid; i < kGemmTM * kGemmKC; i += kGemmThreads) {
            const int m = i / kGemmKC, kk = i - m * kGemmKC;
            float v = 0.f;
            if (kk < kc && rowA[m] >= 0) {
                const int kg = k0 + kk;
                const int s = kg / p.a_split;
                v = __ldg(A + rowA[m] + (int64_t)s * p.a_split_stride + (kg - s * p.a_split));
            }
            As[kk][m] = v;
        }
        for (int i = tid; i < kGemmKC * NC; i += kGemmThreads) {
            const int kk = i / NC, c = i - kk * NC;
            float v = 0.f;
            if (kk < kc && c0 + c < p.Ncol) v = __ldg(B + (int64_t)(k0 + kk) * p.ldb_k + (int64_t)(c0 + c) * p.ldb_c);
            Bs[kk][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kGemmKC; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a[RM];
#pragma unroll
            for (int r = 0; r < RM; ++r) a[r] = As[kk][ty * RM + r];
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                acc[r][0] = fmaf(a[r], w.x, acc[r][0]);
                acc[r][1] = fmaf(a[r], w.y, acc[r][1]);
                acc[r][2] = fmaf(a[r], w.z, acc[r][2]);
                acc[r][3] = fmaf(a[r], w.w, acc[r][3]);
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < RM; ++r) {
        const int ml = ty * RM + r;
        const int64_t rc = rowC[ml];
        if (rc < 0) continue;
        const int n = (m0 + ml) / p.Q;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int col = c0 + tx * 4 + c;
            if (col >= p.Ncol) continue;
            float v = acc[r][c];
            if (p.bias_mode == TGCN_BIAS_PER_VERTEX) v += __ldg(p.bias + (int64_t)n * p.Ncol + col);
            else if (p.bias_mode == TGCN_BIAS_PER_FILTER) v += __ldg(p.bias + col);
            C[rc + col] = v;
        }
    }
}

static int launch_small_gemm(const GemmParams& p, int batches, cudaStream_t st) {
    if (p.M == 0 || p.Ncol == 0 || batches == 0) return TGCN_OK;
    const unsigned gx = (unsigned)ceil_div(p.M, kGemmTM);
    if (p.Ncol <= 16) {
        dim3 grid(gx, (unsigned)ceil_div(p.Ncol, 16), (unsigned)batches);
        small_gemm_kernel<4><<<grid, kGemmThreads, 0, st>>>(p);
    } else if (p.Ncol <= 32) {
        dim3 grid(gx, (unsigned)ceil_div(p.Ncol, 32), (unsigned)batches);
        small_gemm_kernel<8><<<grid, kGemmThreads, 0, st>>>(p);
    } else {
        dim3 grid(gx, (unsigned)ceil_div(p.Ncol, 64), (unsigned)batches);
        small_gemm_kernel<16><<<grid, kGemmThreads, 0, st>>>(p);
    }
    TGCN_LAUNCH_CHECK("small_gemm");
    return TGCN_OK;
}

// ------------------------------------------------------------------------------------------------
// dWmix[jd][g] = sum_m stack[j][m][d] * dout[api(m)][g]   -- long reduction, tiny output.
// Deterministic: persistent blocks accumulate in registers, write partials, second pass sums.
// ------------------------------------------------------------------------------------------------
constexpr int kBwTM = 64;
constexpr int kBwThreads = 256;
constexpr int kBwJD = 32;  // (j,d) outputs per block
constexpr int kBwG = 32;   // g outputs per block

static int bwd_w_partitions(int64_t M, int JD, int G) {
    const int64_t ny = ceil_div(JD, kBwJD), nz = ceil_div(G, kBwG);
    int64_t want = (4 * (int64_t)kNumSMs) / (ny * nz);
    if (want < 1) want = 1;
    const int64_t tiles = ceil_div(M, kBwTM);
    if (want > tiles) want = tiles;
    if (want < 1) want = 1;
    return (int)want;
}

__global__ void __launch_bounds__(kBwThreads)
contract_bwd_w_kernel(const float* __restrict__ stack, const float* __restrict__ dout, float* __restrict__ partial,
                      int M, int Q, int N, int D, int G, int JD, int64_t S) {
    __shared__ __align__(16) float As[kBwTM][kBwJD];
    __shared__ float Ds[kBwTM][kBwG + 1];
    __shared__ int64_t colOff[kBwJD];
    const int tid = threadIdx.x;
    const int g = tid & 31, jq = tid >> 5;  // 8 warps: warp jq owns jd = jq*4 .. jq*4+3
    const int jd0 = blockIdx.y * kBwJD, g0 = blockIdx.z * kBwG;
    if (tid < kBwJD) {
        const int jd = jd0 + tid;
        int64_t off = -1;
        if (jd < JD) {
            const int j = jd / D;
            off = (int64_t)j * S + (jd - j * D);
        }
        colOff[tid] = off;
    }
    __syncthreads();
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int tiles = (M + kBwTM - 1) / kBwTM;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = t * kBwTM;
        for (int i = tid; i < kBwTM * kBwJD; i += kBwThreads) {
            const int m = i >> 5, c = i & 31;
            float v = 0.f;
            if (m0 + m < M && colOff[c] >= 0) v = __ldg(stack + colOff[c] + (int64_t)(m0 + m) * D);
            As[m][c] = v;
        }
        for (int i = tid; i < kBwTM * kBwG; i += kBwThreads) {
            const int m = i >> 5, c = i & 31;
            float v = 0.f;
            const int mm = m0 + m;
            if (mm < M && g0 + c < G) {
                const int n = mm / Q, q = mm - n * Q;
                v = __ldg(dout + ((int64_t)q * N + n) * G + g0 + c);
            }
            Ds[m][c] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int m = 0; m < kBwTM; ++m) {
            const float dv = Ds[m][g];
            const float4 a = *reinterpret_cast<const float4*>(&As[m][jq * 4]);
            acc0 = fmaf(a.x, dv, acc0);
            acc1 = fmaf(a.y, dv, acc1);
            acc2 = fmaf(a.z, dv, acc2);
            acc3 = fmaf(a.w, dv, acc3);
        }
        __syncthreads();
    }
    if (g0 + g < G) {
        float* dst = partial + (int64_t)blockIdx.x * JD * G;
        const int jd = jd0 + jq * 4;
        if (jd + 0 < JD) dst[(int64_t)(jd + 0) * G + g0 + g] = acc0;
        if (jd + 1 < JD) dst[(int64_t)(jd + 1) * G + g0 + g] = acc1;
        if (jd + 2 < JD) dst[(int64_t)(jd + 2) * G + g0 + g] = acc2;
        if (jd + 3 < JD) dst[(int64_t)(jd + 3) * G + g0 + g] = acc3;
    }
}

// out[i] = sum_p partial[p][i] in a FIXED order: 8 lanes per element take the partials p = l, l + 8, ... (each a
// sequential sum), then a 3-step butterfly -- the same tree for every launch, so the result is deterministic, and
// the P dependent loads of the old one-thread-per-element version (19 us at P = 148) become P / 8.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int P, int64_t n) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = t >> 3;
    const int l = (int)(t & 7);
    float s = 0.f;
    if (i < n)
        for (int p = l; p < P; p += 8) s += __ldg(partial + (int64_t)p * n + i);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (i < n && l == 0) out[i] = s;
}

// ------------------------------------------------------------------------------------------------
// weight mixing (reference recursion):  M[k,j] = c_j * (-1)^((k-j)/2) for k >= j, k-j even,
// c_j = 1 (j < 2) or 2 (j >= 2); from Xt_0 = P_0, Xt_1 = P_1, Xt_k = 2 P_k - Xt_{k-2}.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mix_weights_kernel(const float* __restrict__ src, float* __restrict__ dst, int K, int64_t inner, int transpose) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * inner) return;
    const int a = (int)(i / inner);
    const int64_t e = i - (int64_t)a * inner;
    double s = 0.0;
    if (!transpose) {  // dst[j=a] = sum_{k = a, a+2, ...} M[k,a] src[k]
        const double c = a < 2 ? 1.0 : 2.0;
        double sign = 1.0;
        for (int k = a; k < K; k += 2, sign = -sign) s += sign * c * (double)__ldg(src + (int64_t)k * inner + e);
    } else {           // dst[k=a] = sum_{j = a, a-2, ...} M[a,j] src[j]
        double sign = 1.0;
        for (int j = a; j >= 0; j -= 2, sign = -sign) {
            const double c = j < 2 ? 1.0 : 2.0;
            s += sign * c * (double)__ldg(src + (int64_t)j * inner + e);
        }
    }
    dst[i] = (float)s;
}

// The same mix between weight tensors whose per-order row counts differ: src [K][Ds][G] -> dst [K][Dd][G].  Forward
// (Dd >= Ds): the mixed weights of a slab padded to Dd columns per sample, padding rows zero; backward (Dd <= Ds): the
// gradient of the real rows, padding rows dropped.
__global__ void __launch_bounds__(256)
mix_weights_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int K, int Ds, int Dd, int G, int transpose) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t inner_d = (int64_t)Dd * G, inner_s = (int64_t)Ds * G;
    if (i >= (int64_t)K * inner_d) return;
    const int a = (int)(i / inner_d);
    const int64_t e = i - (int64_t)a * inner_d;
    if (e >= inner_s) { dst[i] = 0.f; return; }        // padding row of the destination (rows are the slow index)
    double s = 0.0;
    if (!transpose) {
        const double c = a < 2 ? 1.0 : 2.0;
        double sign = 1.0;
        for (int k = a; k < K; k += 2, sign = -sign) s += sign * c * (double)__ldg(src + (int64_t)k * inner_s + e);
    } else {
        double sign = 1.0;
        for (int j = a; j >= 0; j -= 2, sign = -sign) {
            const double c = j < 2 ? 1.0 : 2.0;
            s += sign * c * (double)__ldg(src + (int64_t)j * inner_s + e);
        }
    }
    dst[i] = (float)s;
}

// ------------------------------------------------------------------------------------------------
// bias gradient
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bias_grad_vertex_kernel(const float* __restrict__ dout, float* __restrict__ db, int Q, int64_t NG) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= NG) return;
    float s = 0.f;
    for (int q = 0; q < Q; ++q) s += __ldg(dout + (int64_t)q * NG + i);
    db[i] = s;
}

// partial[blk][g] = sum over the block's rows; rows = Q*N, G columns (any G)
__global__ void __launch_bounds__(256)
bias_grad_filter_kernel(const float* __restrict__ dout, float* __restrict__ partial, int64_t rows, int G) {
    __shared__ float red[256];
    const int tid = threadIdx.x;
    for (int gbase = 0; gbase < G; gbase += 32) {
        const int g = gbase + (tid & 31);
        float s = 0.f;
        if (g < G)
            for (int64_t r = blockIdx.x * 8 + (tid >> 5); r < rows; r += (int64_t)gridDim.x * 8)
                s += __ldg(dout + r * G + g);
        red[tid] = s;
        __syncthreads();
        if (tid < 32) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w * 32 + tid];
            if (g < G) partial[(int64_t)blockIdx.x * G + g] = t;
        }
        __syncthreads();
    }
}

static int bias_filter_blocks(int64_t rows) {
    int64_t b = ceil_div(rows, 8 * 16);
    if (b > 2 * kNumSMs) b = 2 * kNumSMs;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_mix_weights(const float* src, float* dst, int K, int64_t inner, int recursion,
                                int transpose, void* stream) {
    TGCN_REQUIRE(K >= 1 && inner >= 0, "tgcn_mix_weights: bad sizes");
    if (inner == 0) return TGCN_OK;
    TGCN_REQUIRE(src && dst && src != dst, "tgcn_mix_weights: null or aliased pointer");
    cudaStream_t st = as_stream(stream);
    if (recursion == TGCN_RECURSION_CHEBYSHEV) {
        cudaError_t e = cudaMemcpyAsync(dst, src, sizeof(float) * K * inner, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_mix_weights: %s", cudaGetErrorString(e));
        return TGCN_OK;
    }
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE, "tgcn_mix_weights: unknown recursion %d", recursion);
    mix_weights_kernel<<<(unsigned)ceil_div((int64_t)K * inner, 256), 256, 0, st>>>(src, dst, K, inner, transpose);
    TGCN_LAUNCH_CHECK("mix_weights");
    return TGCN_OK;
}

namespace tgcn {
// src [K][Ds][G] -> dst [K][Dd][G] with the recursion's mix (identity for the textbook recursion); see the kernel
int mix_weights_rows(const float* src, float* dst, int K, int Ds, int Dd, int G, int recursion, int transpose, cudaStream_t st) {
    if ((int64_t)K * Dd * G == 0) return TGCN_OK;
    if (Ds == Dd) return tgcn_mix_weights(src, dst, K, (int64_t)Ds * G, recursion, transpose, st);
    if (recursion == TGCN_RECURSION_CHEBYSHEV) {
        const int rows = Ds < Dd ? Ds : Dd;
        if (Dd > Ds) {
            cudaError_t e0 = cudaMemsetAsync(dst, 0, sizeof(float) * K * Dd * G, st);
            if (e0 != cudaSuccess) return set_error(TGCN_ERR_CUDA, "mix_weights: %s", cudaGetErrorString(e0));
        }
        cudaError_t e = cudaMemcpy2DAsync(dst, sizeof(float) * Dd * G, src, sizeof(float) * Ds * G, sizeof(float) * rows * G, K,
                                          cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "mix_weights: %s", cudaGetErrorString(e));
        return TGCN_OK;
    }
    mix_weights_rows_kernel<<<(unsigned)ceil_div((int64_t)K * Dd * G, 256), 256, 0, st>>>(src, dst, K, Ds, Dd, G, transpose);
    TGCN_LAUNCH_CHECK("mix_weights");
    return TGCN_OK;
}
int tc_supported(int Q, int N, int D, int G, int K);
int64_t tc_fwd_scratch_bytes(int Q, int N, int D, int G, int K);
int64_t tc_bwd_scratch_bytes(int Q, int N, int D, int G, int K);
int contract_fwd_tc(const float* stack, const float* Wmix, const float* bias, int bias_mode, float* out, void* scratch,
                    int Q, int N, int D, int G, int K, cudaStream_t st);
int contract_bwd_x_tc(const float* dout, const float* Wmix, float* gstack, void* scratch,
                      int Q, int N, int D, int G, int K, cudaStream_t st);
int contract_bwd_w_tc(const float* stack, const float* dout, float* partial, int* P_out,
                      int Q, int N, int D, int G, int K, cudaStream_t st);
uint8_t* tc_bwd_partial_ptr(void* scratch, int Q, int N, int D, int G, int K);
}  // namespace tgcn

// engine resolution: AUTO picks the tensor-core engine whenever its tiles cover the shape
static int resolve_engine(const char* who, int engine, int Q, int N, int D, int G, int K, int* use_tc) {
    TGCN_REQUIRE(engine == TGCN_ENGINE_AUTO || engine == TGCN_ENGINE_FFMA || engine == TGCN_ENGINE_TCGEN05,
                 "%s: unknown engine %d", who, engine);
    const int ok = tc_supported(Q, N, D, G, K);
    if (engine == TGCN_ENGINE_TCGEN05)
        TGCN_SUPPORTED(ok, "%s: tcgen05 engine does not cover D=%d G=%d", who, D, G);
    *use_tc = (engine != TGCN_ENGINE_FFMA) && ok;
    return TGCN_OK;
}

static int check_dims(const char* who, int Q, int N, int D, int G, int K) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && G >= 1 && K >= 1, "%s: bad sizes Q=%d N=%d D=%d G=%d K=%d", who, Q, N, D, G, K);
    TGCN_SUPPORTED((int64_t)Q * N < (int64_t)INT32_MAX, "%s: Q*N = %lld exceeds 2^31-1", who, (long long)Q * N);
    return TGCN_OK;
}

extern "C" int64_t tgcn_contract_fwd_scratch(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return tc_supported(Q, N, D, G, K) ? tc_fwd_scratch_bytes(Q, N, D, G, K) : 0;
}

extern "C" int tgcn_contract_fwd(const float* stack, const float* Wmix, const float* bias, int bias_mode,
                                 float* out, void* scratch, int Q, int N, int D, int G, int K, int engine,
                                 void* stream) {
    TGCN_PROPAGATE(check_dims("tgcn_contract_fwd", Q, N, D, G, K));
    if ((int64_t)Q * N == 0) return TGCN_OK;
    TGCN_REQUIRE(stack && Wmix && out, "tgcn_contract_fwd: null pointer");
    TGCN_REQUIRE(bias_mode == TGCN_BIAS_NONE || bias, "tgcn_contract_fwd: bias_mode %d without bias", bias_mode);
    int use_tc = 0;
    TGCN_PROPAGATE(resolve_engine("tgcn_contract_fwd", engine, Q, N, D, G, K, &use_tc));
    if (use_tc) {
        TGCN_REQUIRE(scratch, "tgcn_contract_fwd: the tcgen05 engine needs tgcn_contract_fwd_scratch() bytes of scratch");
        return contract_fwd_tc(stack, Wmix, bias, bias_mode, out, scratch, Q, N, D, G, K, as_stream(stream));
    }
    GemmParams p{};
    p.A = stack; p.a_api_rows = 0; p.lda = D; p.a_split = D; p.a_split_stride = (int64_t)N * Q * D; p.a_batch_stride = 0;
    p.B = Wmix; p.ldb_k = G; p.ldb_c = 1; p.b_batch_stride = 0;
    p.C = out; p.c_api_rows = 1; p.ldc = G; p.c_batch_stride = 0;
    p.bias = bias; p.bias_mode = bias_mode;
    p.M = Q * N; p.Kred = K * D; p.Ncol = G; p.Q = Q; p.N = N;
    return launch_small_gemm(p, 1, as_stream(stream));
}

extern "C" int tgcn_contract_bwd_x(const float* dout, const float* Wmix, float* gstack, void* scratch,
                                   int Q, int N, int D, int G, int K, int engine, void* stream) {
    TGCN_PROPAGATE(check_dims("tgcn_contract_bwd_x", Q, N, D, G, K));
    if ((int64_t)Q * N == 0) return TGCN_OK;
    TGCN_REQUIRE(dout && Wmix && gstack, "tgcn_contract_bwd_x: null pointer");
    int use_tc = 0;
    TGCN_PROPAGATE(resolve_engine("tgcn_contract_bwd_x", engine, Q, N, D, G, K, &use_tc));
    if (use_tc) {
        TGCN_REQUIRE(scratch, "tgcn_contract_bwd_x: the tcgen05 engine needs tgcn_contract_bwd_w_workspace() bytes of scratch");
        return contract_bwd_x_tc(dout, Wmix, gstack, scratch, Q, N, D, G, K, as_stream(stream));
    }
    GemmParams p{};
    p.A = dout; p.a_api_rows = 1; p.lda = G; p.a_split = G; p.a_split_stride = 0; p.a_batch_stride = 0;
    p.B = Wmix; p.ldb_k = 1; p.ldb_c = G; p.b_batch_stride = (int64_t)D * G;
    p.C = gstack; p.c_api_rows = 0; p.ldc = D; p.c_batch_stride = (int64_t)N * Q * D;
    p.bias = nullptr; p.bias_mode = TGCN_BIAS_NONE;
    p.M = Q * N; p.Kred = G; p.Ncol = D; p.Q = Q; p.N = N;
    return launch_small_gemm(p, K, as_stream(stream));
}

extern "C" int64_t tgcn_contract_bwd_w_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    const int64_t a = (int64_t)bwd_w_partitions((int64_t)Q * N, K * D, G) * K * D * G;
    const int64_t b = (int64_t)bias_filter_blocks((int64_t)Q * N) * G;
    int64_t bytes = (int64_t)sizeof(float) * (a > b ? a : b);
    if (tc_supported(Q, N, D, G, K)) {
        const int64_t t = tc_bwd_scratch_bytes(Q, N, D, G, K);
        if (t > bytes) bytes = t;
    }
    return (bytes + 255) & ~(int64_t)255;
}

extern "C" int tgcn_contract_bwd_w(const float* stack, const float* dout, float* dWmix, void* workspace,
                                   int Q, int N, int D, int G, int K, int engine, void* stream) {
    TGCN_PROPAGATE(check_dims("tgcn_contract_bwd_w", Q, N, D, G, K));
    TGCN_REQUIRE(dWmix, "tgcn_contract_bwd_w: null pointer");
    cudaStream_t st = as_stream(stream);
    const int JD = K * D;
    const int64_t M = (int64_t)Q * N;
    if (M == 0) {
        cudaError_t e = cudaMemsetAsync(dWmix, 0, sizeof(float) * JD * G, st);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_contract_bwd_w: %s", cudaGetErrorString(e));
        return TGCN_OK;
    }
    TGCN_REQUIRE(stack && dout && workspace, "tgcn_contract_bwd_w: null pointer");
    int use_tc = 0;
    TGCN_PROPAGATE(resolve_engine("tgcn_contract_bwd_w", engine, Q, N, D, G, K, &use_tc));
    if (use_tc) {
        // partials live behind the bwd_x weight image inside the same workspace
        float* partial = reinterpret_cast<float*>(tc_bwd_partial_ptr(workspace, Q, N, D, G, K));
        int P = 0;
        TGCN_PROPAGATE(contract_bwd_w_tc(stack, dout, partial, &P, Q, N, D, G, K, st));
        const int64_t n = (int64_t)JD * G;
        reduce_partials_kernel<<<(unsigned)ceil_div(n * 8, 256), 256, 0, st>>>(partial, dWmix, P, n);
        TGCN_LAUNCH_CHECK("reduce_partials");
        return TGCN_OK;
    }
    const int P = bwd_w_partitions(M, JD, G);
    dim3 grid((unsigned)P, (unsigned)ceil_div(JD, kBwJD), (unsigned)ceil_div(G, kBwG));
    contract_bwd_w_kernel<<<grid, kBwThreads, 0, st>>>(stack, dout, (float*)workspace, (int)M, Q, N, D, G, JD,
                                                      (int64_t)N * Q * D);
    TGCN_LAUNCH_CHECK("contract_bwd_w");
    const int64_t n = (int64_t)JD * G;
    reduce_partials_kernel<<<(unsigned)ceil_div(n * 8, 256), 256, 0, st>>>((const float*)workspace, dWmix, P, n);
    TGCN_LAUNCH_CHECK("reduce_partials");
    return TGCN_OK;
}

extern "C" int tgcn_bias_grad(const float* dout, float* db, void* workspace, int Q, int N, int G, int bias_mode,
                              void* stream) {
    TGCN_REQUIRE(Q >= 0 && N >= 0 && G >= 1, "tgcn_bias_grad: bad sizes");
    TGCN_REQUIRE(db, "tgcn_bias_grad: null pointer");
    cudaStream_t st = as_stream(stream);
    const int64_t NG = (int64_t)N * G;
    if (bias_mode == TGCN_BIAS_PER_VERTEX) {
        if (NG == 0) return TGCN_OK;
        TGCN_REQUIRE(dout || Q == 0, "tgcn_bias_grad: null pointer");
        bias_grad_vertex_kernel<<<(unsigned)ceil_div(NG, 256), 256, 0, st>>>(dout, db, Q, NG);
        TGCN_LAUNCH_CHECK("bias_grad_vertex");
        return TGCN_OK;
    }
    TGCN_REQUIRE(bias_mode == TGCN_BIAS_PER_FILTER, "tgcn_bias_grad: bias_mode %d has no gradient", bias_mode);
    const int64_t rows = (int64_t)Q * N;
    if (rows == 0) {
        cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * G, st);
        if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_bias_grad: %s", cudaGetErrorString(e));
        return TGCN_OK;
    }
    TGCN_REQUIRE(dout && workspace, "tgcn_bias_grad: null pointer");
    const int B = bias_filter_blocks(rows);
    bias_grad_filter_kernel<<<B, 256, 0, st>>>(dout, (float*)workspace, rows, G);
    TGCN_LAUNCH_CHECK("bias_grad_filter");
    reduce_partials_kernel<<<(unsigned)ceil_div((int64_t)G * 8, 256), 256, 0, st>>>((const float*)workspace, db, B, G);
    TGCN_LAUNCH_CHECK("reduce_partials");
    return TGCN_OK;
}
