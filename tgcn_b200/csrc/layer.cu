// Whole-layer entry points: one host call per layer direction so that the Python side issues two
// C calls per layer per training step (and the whole step can be captured in a CUDA graph).
#include "common.cuh"

using namespace tgcn;

static int64_t wmix_bytes(int D, int G, int K) {
    return (((int64_t)sizeof(float) * K * D * G) + 255) & ~(int64_t)255;
}

extern "C" int tgcn_layer_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                              const float* x, const float* W, const float* bias, int bias_mode,
                              float* out, float* stack, void* workspace,
                              int Q, int D, int G, int K, int recursion, int engine, void* stream) {
    TGCN_REQUIRE(W && workspace, "tgcn_layer_fwd: null weight / workspace pointer");
    float* Wmix = reinterpret_cast<float*>(workspace);            // [K,D,G], kept for the backward
    void* scratch = reinterpret_cast<char*>(workspace) + wmix_bytes(D, G, K);
    TGCN_PROPAGATE(tgcn_cheb_basis(rowptr, col, val, N, x, stack, Q, D, K, recursion, stream));
    TGCN_PROPAGATE(tgcn_mix_weights(W, Wmix, K, (int64_t)D * G, recursion, 0, stream));
    return tgcn_contract_fwd(stack, Wmix, bias, bias_mode, out, scratch, Q, N, D, G, K, engine, stream);
}

// forward workspace: [ Wmix (K*D*G fp32, 256-byte padded) | contraction scratch ]
extern "C" int64_t tgcn_layer_fwd_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return wmix_bytes(D, G, K) + tgcn_contract_fwd_scratch(Q, N, D, G, K);
}

extern "C" int tgcn_layer_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N,
                              const float* dout, const float* stack, const float* Wmix,
                              float* dW, float* db, int bias_mode, float* dx, float* gstack, void* workspace,
                              int Q, int D, int G, int K, int recursion, int engine, void* stream) {
    TGCN_REQUIRE(dW && workspace, "tgcn_layer_bwd: null pointer");
    // dW: P^T dOut on the power basis, then the transposed mix back onto the reference's weights.
    // The mixed gradient is staged at the tail of `workspace` (the partials occupy its head).
    const int64_t ws_bytes = tgcn_contract_bwd_w_workspace(Q, N, D, G, K);
    float* dWmix = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + ws_bytes);
    TGCN_PROPAGATE(tgcn_contract_bwd_w(stack, dout, dWmix, workspace, Q, N, D, G, K, engine, stream));
    TGCN_PROPAGATE(tgcn_mix_weights(dWmix, dW, K, (int64_t)D * G, recursion, 1, stream));
    if (bias_mode != TGCN_BIAS_NONE) {
        TGCN_REQUIRE(db, "tgcn_layer_bwd: bias_mode %d without db", bias_mode);
        TGCN_PROPAGATE(tgcn_bias_grad(dout, db, workspace, Q, N, G, bias_mode, stream));
    }
    if (dx) {
        TGCN_REQUIRE(gstack && Wmix, "tgcn_layer_bwd: dx requested without gstack/Wmix");
        TGCN_PROPAGATE(tgcn_contract_bwd_x(dout, Wmix, gstack, workspace, Q, N, D, G, K, engine, stream));
        TGCN_PROPAGATE(tgcn_cheb_adjoint(rowptrT, colT, valT, N, gstack, dx, Q, D, K, recursion, stream));
    }
    return TGCN_OK;
}

// total scratch for tgcn_layer_bwd: reduction partials + the staged mixed weight gradient
extern "C" int64_t tgcn_layer_bwd_workspace(int Q, int N, int D, int G, int K) {
    if (Q < 0 || N < 0 || D < 1 || G < 1 || K < 1) return 0;
    return tgcn_contract_bwd_w_workspace(Q, N, D, G, K) + (int64_t)sizeof(float) * K * D * G;
}
