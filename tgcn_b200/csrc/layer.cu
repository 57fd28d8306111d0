// Whole-layer entry points: one host call per layer direction so that the Python side issues two
// C calls per layer per training step (and the whole step can be captured in a CUDA graph).
//
// Slab width: the streaming kernels work on slabs of Q*Dp columns, Dp >= D = H*F.  tgcn_layer_slab_width picks
// Dp = D rounded up to a multiple of 32 when that costs <= 12.5 % more slab bytes and the tensor-core engine runs
// (its TMA-fed kernels, csrc/contract_tc3.cu, want slab rows of whole 128-byte blocks: the cortical-mesh layer 1
// has D = 30); the padding columns are written as zeros by the layout kernel, stay zero through the recursion,
// meet zero rows of the mixed weights, and are dropped again on the way out of the backward.
#include "common.cuh"

using namespace tgcn;

namespace tgcn {
int cheb_basis_padded(const int32_t* rowptr, const int32_t* col, const float* val, int N, const float* x, float* stack,
                      int Q, int D, int Dp, int K, int recursion, cudaStream_t st);
int cheb_adjoint_padded(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N, float* gstack, float* dx,
                        int Q, int D, int Dp, int K, int recursion, cudaStream_t st);
int mix_weights_rows(const float* src, float* dst, int K, int Ds, int Dd, int G, int recursion, int transpose, cudaStream_t st);
int tc_supported(int Q, int N, int D, int G, int K);
}  // namespace tgcn

static int64_t wmix_bytes(int D, int G, int K) {
    return (((int64_t)sizeof(float) * K * D * G) + 255) & ~(int64_t)255;
}

extern "C" int tgcn_layer_slab_width(int Q, int N, int D, int G, int K, int engine) {
    if (D < 1) return D;
    const int Dp = (D + 31) / 32 * 32;
    if (Dp == D || engine == TGCN_ENGINE_FFMA || (Dp - D) * 8 > D) return D;
    return tc_supported(Q, N, Dp, G, K) ? Dp : D;
}

extern "C" int tgcn_layer_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                              const float* x, const float* W, const float* bias, int bias_mode,
                              float* out, float* stack, void* workspace,
                              int Q, int D, int Dp, int G, int K, int recursion, int engine, void* stream) {
    TGCN_REQUIRE(W && workspace, "tgcn_layer_fwd: null weight / workspace pointer");
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && Dp >= D && G >= 1 && K >= 1, "tgcn_layer_fwd: bad sizes Q=%d N=%d D=%d Dp=%d G=%d K=%d", Q, N, D, Dp, G, K);
    TGCN_REQUIRE(recursion == TGCN_RECURSION_REFERENCE || recursion == TGCN_RECURSION_CHEBYSHEV, "tgcn_layer_fwd: unknown recursion %d", recursion);
    float* Wmix = reinterpret_cast<float*>(workspace);            // [K,Dp,G], kept for the backward
    void* scratch = reinterpret_cast<char*>(workspace) + wmix_bytes(Dp, G, K);
    cudaStream_t st = as_stream(stream);
    if ((int64_t)Q * N > 0) {
        TGCN_REQUIRE(rowptr && x && stack, "tgcn_layer_fwd: null pointer");
        TGCN_PROPAGATE(cheb_basis_padded(rowptr, col, val, N, x, stack, Q, D, Dp, K, recursion, st));
    }
    TGCN_PROPAGATE(mix_weights_rows(W, Wmix, K, D, Dp, G, recursion, 0, st));
    return tgcn_contract_fwd(stack, Wmix, bias, bias_mode, out, scratch, Q, N, Dp, G, K, engine, stream);
}

// forward workspace: [ Wmix (K*Dp*G fp32, 256-byte padded) | contraction scratch ]; pass the slab width Dp as D
extern "C" int64_t tgcn_layer_fwd_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return wmix_bytes(D, G, K) + tgcn_contract_fwd_scratch(Q, N, D, G, K);
}

extern "C" int tgcn_layer_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N,
                              const float* dout, const float* stack, const float* Wmix,
                              float* dW, float* db, int bias_mode, float* dx, float* gstack, void* workspace,
                              int Q, int D, int Dp, int G, int K, int recursion, int engine, void* stream) {
    TGCN_REQUIRE(dW && workspace, "tgcn_layer_bwd: null pointer");
    TGCN_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && Dp >= D && G >= 1 && K >= 1, "tgcn_layer_bwd: bad sizes");
    cudaStream_t st = as_stream(stream);
    // dW: P^T dOut on the power basis, then the transposed mix back onto the reference's weights.
    // The mixed gradient is staged at the tail of `workspace` (the partials occupy its head).
    const int64_t ws_bytes = tgcn_contract_bwd_w_workspace(Q, N, Dp, G, K);
    float* dWmix = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ws_bytes);
    TGCN_PROPAGATE(tgcn_contract_bwd_w(stack, dout, dWmix, workspace, Q, N, Dp, G, K, engine, stream));
    TGCN_PROPAGATE(mix_weights_rows(dWmix, dW, K, Dp, D, G, recursion, 1, st));
    if (bias_mode != TGCN_BIAS_NONE) {
        TGCN_REQUIRE(db, "tgcn_layer_bwd: bias_mode %d without db", bias_mode);
        TGCN_PROPAGATE(tgcn_bias_grad(dout, db, workspace, Q, N, G, bias_mode, stream));
    }
    if (dx) {
        TGCN_REQUIRE(gstack && Wmix, "tgcn_layer_bwd: dx requested without gstack/Wmix");
        TGCN_PROPAGATE(tgcn_contract_bwd_x(dout, Wmix, gstack, workspace, Q, N, Dp, G, K, engine, stream));
        if ((int64_t)Q * N > 0) {
            TGCN_REQUIRE(rowptrT, "tgcn_layer_bwd: dx requested without the CSR of L^T");
            TGCN_PROPAGATE(cheb_adjoint_padded(rowptrT, colT, valT, N, gstack, dx, Q, D, Dp, K, recursion, st));
        }
    }
    return TGCN_OK;
}

// total scratch for tgcn_layer_bwd: reduction partials + the staged mixed weight gradient; pass the slab width Dp as D
extern "C" int64_t tgcn_layer_bwd_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return tgcn_contract_bwd_w_workspace(Q, N, D, G, K) + (int64_t)sizeof(float) * K * D * G;
}
