/*
 * tgcn_b200.h -- C-ABI of the B200-native time-vertex Chebyshev graph-convolution path.
 *
 * Drop-in boundary for the hot path of cassianobecker/tgcn (`tgcn/nn/gcn.py`).  The
 * reference has no FFI (it is pure Python over ATen); each entry point below names the
 * reference code whose device work it replaces (file:line into the reference tree).
 * Callers are the `torch.autograd.Function`s in `tgcn_b200/nn/functional.py` (ctypes), or any
 * C/C++ host.  See INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - all tensors are fp32, dense, contiguous; index arrays are int32;
 *   - `stream` is a `cudaStream_t` passed as `void*` (PyTorch: torch.cuda.current_stream().cuda_stream);
 *   - every call is asynchronous on `stream`, never synchronises the device, never allocates or
 *     frees device memory and is legal inside CUDA-graph stream capture;
 *   - return value: 0 on success, <0 on error (TGCN_ERR_*); `tgcn_last_error()` returns the
 *     message of the calling thread's last failure.  There is no CPU fallback.
 *
 * Shapes: Q batch, N vertices (padded), D = H*F features per vertex of
 * the input signal (H = horizon, F = in_channels; H = 1 for the spatial-only layers), G =
 * out_channels, K = filter order.  API tensors keep the reference layout x[Q,N,D], out[Q,N,G].
 * Internally the signal is a vertex-major "slab" [N, C] with C = Q*D columns (column = q*D + d),
 * so a neighbour row is one contiguous, coalesced C*4-byte read; the stacked basis is
 * `stack[K][N][C]`.
 */
#ifndef TGCN_B200_H
#define TGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGCN_OK 0
#define TGCN_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, misaligned buffer) */
#define TGCN_ERR_UNSUPPORTED (-2) /* shape outside what the kernels cover */
#define TGCN_ERR_CUDA (-3)        /* a CUDA runtime call / kernel launch failed */

/* bias_mode */
#define TGCN_BIAS_NONE 0
#define TGCN_BIAS_PER_VERTEX 1 /* bias[1,N,G]: TGCNCheb_H / TGCNCheb  (gcn.py:96, :22)  */
#define TGCN_BIAS_PER_FILTER 2 /* bias[1,1,G]: GCNCheb                (gcn.py:172)      */

/* recursion */
#define TGCN_RECURSION_REFERENCE 0 /* Xt_k = 2 L^k X - Xt_{k-2}: what gcn.py:146-153 computes   */
#define TGCN_RECURSION_CHEBYSHEV 1 /* T_k  = 2 L T_{k-1} - T_{k-2}: textbook recursion           */

/* contraction engine */
#define TGCN_ENGINE_AUTO 0
#define TGCN_ENGINE_FFMA 1    /* fp32 CUDA-core path (exact fp32 accumulation)                  */
#define TGCN_ENGINE_TCGEN05 2 /* tcgen05.mma kind::tf32, 3xTF32 split, fp32 accumulate in TMEM   */
#define TGCN_ENGINE_RESIDENT 3 /* sample-resident fused layer kernels (small graphs), fp32 FFMA    */

/* Fused dropout (nn.Dropout of the reference models, pytorch_hcp_tgcn.py:106,125,136,150): out = x * mask / (1-p)
 * with mask ~ Bernoulli(1-p) drawn per activation element from a counter-based hash of (seed, *step, element index).
 * `step` points to a device uint32 the caller advances once per training step (NULL: 0), so a CUDA-graph replay
 * draws a fresh mask.  NULL descriptor or p <= 0: no dropout.  p must be < 1. */
typedef struct tgcn_dropout {
    float p;
    uint32_t seed;
    const uint32_t* step;
} tgcn_dropout_t;

int tgcn_version(void);
const char* tgcn_last_error(void);
/* 1 when the library was compiled for sm_100a and the current device is compute capability 10.x */
int tgcn_device_supported(void);
/* kernel launches issued by this library so far (host counter; a launch recorded during CUDA-graph
 * capture counts once) */
long long tgcn_launch_count(void);
/* Kernel-variant selection for tests and sweeps (also readable from the environment as TGCN_<KEY>):
 * "SPMM_PIPE" = blocks per SM of the persistent, software-pipelined per-entry SpMM kernel (0 = plain kernel),
 * "SPMM_CSM" = stage the CSR entries of a row block in shared memory (per-entry kernel), "SPMM_RTILE" = use
 * registered row-tile plans (2, the default: persistent plan-prefetching kernel for 4-row tiles, launched with
 * programmatic stream serialization -- TGCN_SPMM_PDL=0 in the environment turns that off; 1 = one-shot kernel;
 * 3 = persistent for 8-row tiles too, 128-thread blocks for 4-row tiles; 4 = one-shot kernel, 8-row tiles built for 4
 * blocks per SM; 0 = off), "RES_TC" = contraction of
 * the resident forward kernel on tcgen05 with 3xTF32 operands and TMEM accumulators (1) or on the fp32 FFMA pipe (0,
 * the default: measured faster at the resident shapes), "RES_ENT" = keep each thread's CSR entries in registers
 * across the K steps of the resident forward (bit-identical; 0 default).  The per-entry SpMM variants are
 * bit-identical with each other; RES_TC changes the contraction's rounding (<= 5e-6). */
int tgcn_set_tuning(const char* key, int value);

/* ---- row-tile plans: register-tiled SpMM ("SPMM_RTILE" tuning key, on whenever a plan is registered) ---- */
/* A thread of the row-tile kernel owns one float4 column of R consecutive rows (R = 4 or 8) and loads every DISTINCT
 * source row of the tile once, applying it to all R rows with a dense coefficient vector -- 2-2.75x fewer gathers than
 * one per CSR entry on graphs whose row order has locality (coarsening order, Morton order).  Same mathematical
 * operation as the reference's sparse product (gcn_matmul.py:154, gcn.py:147); the summation order per output element
 * is ascending source row, so results agree with the other SpMM kernels to fp32 rounding, not bit for bit.
 * tgcn_rowtile_plan_host builds the plan on the host (src_host == w_host == NULL: size query; returns the number of
 * (tile, source) pairs; pad > 1 pads every non-empty tile to a multiple of `pad` entries with zero-coefficient repeats
 * of its last source, so a kernel with `pad` gathers in flight has no one-at-a-time tail; 1 = no padding); upload the arrays (w_dev 16-byte aligned) and register them keyed by the device address of
 * the CSR `col` array; n_src_rows = number of rows of the gathered operand (every source id < n_src_rows: N for a
 * square operand, N + halo rows for a row partition).  The arrays stay owned by the caller and must outlive the plan.
 * A plan holds the operand's VALUES (the coefficient vectors): destroy and rebuild it when val[] changes.  A zero
 * coefficient still multiplies its source row, so a non-finite activation reaches every row of a tile that shares the
 * source (exact for finite inputs). */
int64_t tgcn_rowtile_plan_host(const int32_t* rowptr_host, const int32_t* col_host, const float* val_host, int N, int R,
                               int pad, int32_t* tile_ptr_host, int32_t* src_host, float* w_host);
int64_t tgcn_rowtile_plan_create(const int32_t* col_dev, int N, int n_src_rows, int R, const int32_t* tile_ptr_dev,
                                 const int32_t* src_dev, const float* w_dev);
int tgcn_rowtile_plan_destroy(int64_t handle);

/* ---- layout ------------------------------------------------------------------------------ */
/* x[Q,N,D] -> slab[N,Q,D]   (replaces X.permute(1,3,2,0).reshape(N,-1), gcn_matmul.py:152-153) */
int tgcn_to_slab(const float* x, float* slab, int Q, int N, int D, void* stream);
/* slab[N,Q,D] -> x[Q,N,D]   (replaces res.reshape(..).permute(3,0,2,1), gcn_matmul.py:155-156) */
int tgcn_from_slab(const float* slab, float* x, int Q, int N, int D, void* stream);

/* ---- K1: CSR SpMM recursion step ---------------------------------------------------------- */
/* out[r,:] = alpha * sum_e val[e] * in[col[e],:] + beta * prev[r,:]   for r in [0,N), C columns.
 * Replaces einsum("nm,qmhf->qnhf", L, X) + `2*X - Xt[k-2]` + the slab copy (gcn.py:147-153;
 * torch.mm at gcn_matmul.py:154).  `prev` may be NULL (beta ignored) and may alias `out`;
 * `in` must not alias `out`.  `in` may have more rows than N (halo rows of a row partition). */
int tgcn_spmm_step(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                   const float* in, const float* prev, float* out, int64_t C,
                   float alpha, float beta, void* stream);

/* Whole basis: stack[0] = slab(x); stack[j] per `recursion` (REFERENCE: powers L^j x;
 * CHEBYSHEV: T_j x).  Replaces `_time_chebyshev` / `_chebyshev` (gcn.py:126-154, :208-237, :52-79). */
int tgcn_cheb_basis(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                    const float* x, float* stack, int Q, int D, int K, int recursion, void* stream);

/* Stacked basis in the reference's own layout Xt[K,Q,N,D] from the internal stack
 * (for callers of `layer._time_chebyshev(x)`): Xt_k = sum_j M[k,j] stack[j]. */
int tgcn_basis_to_reference(const float* stack, float* Xt, int Q, int N, int D, int K,
                            int recursion, void* stream);

/* ---- weights ------------------------------------------------------------------------------ */
/* REFERENCE recursion only: Wmix[j] = sum_k M[k,j] W[k]   (transpose=0, forward weights) or
 * dW[k] = sum_j M[k,j] dWmix[j] (transpose=1, gradient), M from Xt_k = sum_j M[k,j] L^j X.
 * `inner` = D*G elements per order.  CHEBYSHEV recursion: plain copy. */
int tgcn_mix_weights(const float* src, float* dst, int K, int64_t inner, int recursion,
                     int transpose, void* stream);

/* ---- K2: weight contraction ---------------------------------------------------------------- */
/* out[q,n,g] = sum_{j,d} stack[j][n][q*D+d] * Wmix[j][d][g] (+ bias)
 * Replaces einsum("kqnhf,khfg->qng") + bias add (gcn.py:113-116, :39-42, :194-198). */
/* `scratch`: tgcn_contract_fwd_scratch(...) bytes (weight image of the tcgen05 engine; may be NULL
 * when that returns 0 or engine == TGCN_ENGINE_FFMA). */
int64_t tgcn_contract_fwd_scratch(int Q, int N, int D, int G, int K);
int tgcn_contract_fwd(const float* stack, const float* Wmix, const float* bias, int bias_mode,
                      float* out, void* scratch, int Q, int N, int D, int G, int K, int engine,
                      void* stream);

/* ---- K4: backward ------------------------------------------------------------------------ */
/* Bytes of scratch the backward entry points need (deterministic two-pass reductions of dW / db). */
int64_t tgcn_contract_bwd_w_workspace(int Q, int N, int D, int G, int K);
/* dWmix[j][d][g] = sum_{n,q} stack[j][n][q*D+d] * dout[q,n,g]   (autograd of gcn.py:113). */
int tgcn_contract_bwd_w(const float* stack, const float* dout, float* dWmix, void* workspace,
                        int Q, int N, int D, int G, int K, int engine, void* stream);
/* gstack[j][n][q*D+d] = sum_g dout[q,n,g] * Wmix[j][d][g]     (autograd of gcn.py:113 w.r.t. Xt) */
int tgcn_contract_bwd_x(const float* dout, const float* Wmix, float* gstack, void* workspace,
                        int Q, int N, int D, int G, int K, int engine, void* stream);
/* dx[Q,N,D] from gstack (destroyed) through the adjoint recursion with L^T given as CSR
 * (autograd of gcn.py:146-153). */
int tgcn_cheb_adjoint(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N,
                      float* gstack, float* dx, int Q, int D, int K, int recursion, void* stream);
/* db: per-vertex db[n,g] = sum_q dout[q,n,g]; per-filter db[g] = sum_{q,n} dout[q,n,g]. */
int tgcn_bias_grad(const float* dout, float* db, void* workspace, int Q, int N, int G, int bias_mode,
                   void* stream);

/* ---- K3: permuted max-pool ---------------------------------------------------------------- */
/* y[q,m,g] = max_{s<p} x[q,m*p+s,g]; idx = first maximal s (NaN wins), torch.max(dim) rule.
 * Replaces gcn_pool / gcn_pool_4 (gcn.py:246-255).  `relu` != 0 applies max(x,0) first
 * (F.relu before the pool, pytorch_hcp_tgcn.py:135-137); `drop` (needs relu) then applies the dropout that sits
 * between them (drop1, pytorch_hcp_tgcn.py:136).  idx is uint8 [Q,N/p,G]. */
int tgcn_pool_max_fwd(const float* x, float* y, uint8_t* idx, int Q, int N, int G, int p, int relu,
                      const tgcn_dropout_t* drop, void* stream);
/* dx[q,m*p+s,g] = (s == idx) ? dy[q,m,g] : 0; with relu: additionally 0 where the source did not pass the ReLU --
 * decided from the pool input `x` (x <= 0), or from the forward's pooled output `y` when it is given (y <= 0; the
 * only form that is valid with dropout, whose 1/(1-p) scale is then applied: a positive y means its source was kept). */
int tgcn_pool_max_bwd(const float* dy, const uint8_t* idx, const float* x, const float* y, float* dx, int Q, int N,
                      int G, int p, int relu, const tgcn_dropout_t* drop, void* stream);

/* ---- fused whole-layer entry points (one host call per layer direction) -------------------- */
/* Slab width Dp >= D = H*F of the streaming kernels for this layer: D rounded up to a multiple of 32 when that
 * costs <= 12.5 % more slab bytes and the tensor-core engine runs (its TMA-fed kernels read slab rows of whole
 * 128-byte blocks; the cortical-mesh layer has D = 30), else D.  The padding columns hold zeros and never reach
 * the caller: x, W, dW, dx keep the reference shapes. */
int tgcn_layer_slab_width(int Q, int N, int D, int G, int K, int engine);
/* forward: layout change (+ padding) + basis + mix + contraction.  `stack` [K,N,Q*Dp] and `workspace`
 * (tgcn_layer_fwd_workspace(Q, N, Dp, G, K) bytes; its head is Wmix[K,Dp,G]) are caller-owned and re-used by the
 * backward. */
int64_t tgcn_layer_fwd_workspace(int Q, int N, int D, int G, int K);
int tgcn_layer_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int N,
                   const float* x, const float* W, const float* bias, int bias_mode,
                   float* out, float* stack, void* workspace,
                   int Q, int D, int Dp, int G, int K, int recursion, int engine, void* stream);
/* backward: dW [K,D,G] (always), db (if bias_mode != NONE), dx [Q,N,D] (if dx != NULL; needs gstack [K,N,Q*Dp]).
 * `workspace` holds tgcn_layer_bwd_workspace(Q, N, Dp, G, K) bytes. */
int64_t tgcn_layer_bwd_workspace(int Q, int N, int D, int G, int K);
int tgcn_layer_bwd(const int32_t* rowptrT, const int32_t* colT, const float* valT, int N,
                   const float* dout, const float* stack, const float* Wmix,
                   float* dW, float* db, int bias_mode, float* dx, float* gstack, void* workspace,
                   int Q, int D, int Dp, int G, int K, int recursion, int engine, void* stream);

/* ---- sample-resident fused layer (graphs whose per-sample slab [N,D] fits in shared memory) ---- */
/* One CTA per sample runs the whole layer: recursion in shared memory, contraction accumulated in
 * registers across all K orders, bias (+ optional ReLU + permuted max-pool) in the epilogue; the
 * basis is written to HBM once for the backward.  Replaces, per layer, the K-1 `bmm` launches, the
 * einsum and the following F.relu + gcn_pool(_4) of the reference models (gcn.py:108-154, :189-237,
 * :246-255; pytorch_hcp_tgcn.py:133-141).  `W` is the RAW layer weight [K,D,G] (the recursion's
 * basis change is applied on the fly).  Small batches are split over 2 or 4 CTAs per sample (a
 * thread-block cluster exchanging rows through distributed shared memory in the forward).
 *   out  [Q,N,G] or NULL;  y [Q,N/pool_p,G] + idx (uint8) or NULL (at least one of out / y);
 *   relu != 0 applies max(.,0) before the pool;  stack: tgcn_resident_stack_bytes(...) bytes or NULL
 *   (NULL = inference, no backward).  1 = supported for these sizes, 0 = use the streaming path. */
int tgcn_resident_supported(int N, int D, int G, int K, int64_t nnz);
int64_t tgcn_resident_stack_bytes(int Q, int N, int D, int K);
int64_t tgcn_resident_bwd_workspace(int Q, int N, int D, int G, int K);
/* bytes of the weight images ([K][DP][GP] mixed weights and their transpose) the forward writes into
 * `wimages` and the backward reads back */
int64_t tgcn_resident_weights_bytes(int D, int G, int K);
/* The resident kernels read L~ as a PACKED CSR: rowinfo[n] = (start, len) (2*N int32 rounded up to a multiple of 4, start even) and
 * entries[e] = (col, float bits of val) (2*E int32, 16-byte aligned), with each row's entries ordered
 * so that rows sharing a quarter-warp gather from different shared-memory bank groups.  Build it once
 * per graph on the host with tgcn_pack_csr_host (entries_host == NULL: size query; returns E) using
 * classes = tgcn_resident_pack_classes(Q, N, D, backward) and upload both arrays.  Any `classes`
 * value gives correct results; the matching one gives the fewest bank conflicts. */
int64_t tgcn_pack_csr_host(const int32_t* rowptr_host, const int32_t* col_host, const float* val_host, int N,
                           int classes, int32_t* rowinfo_host, int32_t* entries_host);
int tgcn_resident_pack_classes(int Q, int N, int D, int backward);
int tgcn_resident_layer_fwd(const int32_t* rowinfo, const int32_t* entries, int N, int64_t E,
                            const float* x, const float* W, const float* bias, int bias_mode,
                            float* out, float* y, uint8_t* idx, int pool_p, int relu,
                            const tgcn_dropout_t* drop, float* stack,
                            float* wimages, int Q, int D, int G, int K, int recursion, void* stream);
/* Backward of the above.  Pass exactly one of `dout` [Q,N,G] (un-pooled output was returned) or
 * `dy` [Q,N/pool_p,G] with the forward's `idx` and pooled output `y` (the max-pool / ReLU gradient
 * routing is applied on the fly).  Produces dW [K,D,G], db (per bias_mode), dx [Q,N,D] (if non-NULL;
 * needs the packed CSR of L^T).  `workspace`: tgcn_resident_bwd_workspace(...) bytes.  Deterministic.
 * `drop` (fused pool + relu only): the descriptor given to the forward -- only its p is used (the gradient of a
 * positive pooled output is scaled by 1/(1-p); no mask is regenerated). */
int tgcn_resident_layer_bwd(const int32_t* rowinfoT, const int32_t* entriesT, int N, int64_t E,
                            const float* dout, const float* dy, const uint8_t* idx, const float* y,
                            int pool_p, int relu, const tgcn_dropout_t* drop, const float* stack,
                            const float* wimages,
                            float* dW, float* db, int bias_mode, float* dx, void* workspace,
                            int Q, int D, int G, int K, int recursion, void* stream);

/* ---- fused classifier head -------------------------------------------------- */
/* logp[Q,C] = log_softmax(fc2(relu(batchnorm(fc1(x))))) with x[Q,I], W1[Hd,I], b1[Hd], gamma/beta[Hd] (NULL: 1/0),
 * W2[C,Hd], b2[C]; replaces fc1 / dense1_bn / relu / fc2 / log_softmax of pytorch_hcp_tgcn.py:143-155 (about ten
 * ATen launches) by two launches.  training != 0: batch statistics (biased variance) normalise and the running
 * estimates are updated like torch.nn.BatchNorm1d (momentum, unbiased variance); else the running estimates are
 * used.  act[Q,Hd] (post-ReLU), xhat[Q,Hd] and invstd[Hd] are saved for the backward.  C <= 32. */
int tgcn_head_fwd(const float* x, const float* W1, const float* b1, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, int training,
                  const float* W2, const float* b2, const tgcn_dropout_t* drop, float* act, float* xhat, float* invstd,
                  float* logp, void* workspace, int Q, int I, int Hd, int C, void* stream);
/* Large fc1 (I * Hd > 2^21 weights, Q <= 8, I % 4 == 0; the cortical-mesh model: 167 424 x 200 = 134 MB): the weight is
 * streamed once by a dedicated kernel that needs `workspace` of tgcn_head_workspace(Q, I, Hd) bytes (0: not needed). */
int64_t tgcn_head_workspace(int Q, int I, int Hd);

/* Optimizer step of fc1.weight fused into the head backward (large fc1 only, tgcn_head_fused_update_supported):
 *     buf = momentum * buf + dW1 / world;  W1 -= lr * buf          (torch.optim.SGD semantics, pytorch_hcp_tgcn.py:169,259)
 * applied while W1 is streamed for dx, with dW1 = sum over ranks of dh_r^T x_r formed on the fly -- the 134 MB gradient
 * is never written, and for world > 1 never exchanged: every rank publishes its activations x [Q, I] and dh [Q, Hd] in
 * an IPC-mapped region (tgcn_peer_alloc / _export / _import; tgcn_peer_region_bytes(Q*I + Q*Hd, 2) bytes) and reads
 * the peers' copies over NVLink inside the update kernel (replaces the DataParallel gradient gather for this tensor,
 * pytorch_hcp_tgcn.py:271-272).  `regions`: host array [world] of region bases as mapped in this process (NULL for
 * world 1); `state`: 4 device uint32 owned by this rank, zero-initialised (NULL for world 1). */
typedef struct tgcn_fc1_update {
    float lr, momentum;
    float* mom;             /* [Hd, I] momentum buffer (device) */
    int world, rank;
    void* const* regions;
    unsigned int* state;
    float* gather;          /* world > 1: device scratch of world * (pad4(Q*I) + pad4(Q*Hd)) floats (the peers' x / dh land here) */
} tgcn_fc1_update_t;
int tgcn_head_fused_update_supported(int Q, int I, int Hd);
/* Backward of tgcn_head_fwd (training statistics): gradients of every operand from dlogp[Q,C]; dx may be NULL;
 * dh_scratch holds Q*Hd floats.  `drop`: the forward's descriptor (only p is used).  `upd` != NULL: fc1.weight is
 * updated in place as described above and dW1 may be NULL; `upd` == NULL: dW1 is written. */
int tgcn_head_bwd(const float* dlogp, const float* logp, const float* act, const float* xhat, const float* invstd,
                  const float* x, float* W1, const float* gamma, const float* W2,
                  float* dx, float* dW1, float* db1, float* dgamma, float* dbeta, float* dW2, float* db2,
                  float* dh_scratch, const tgcn_dropout_t* drop, const tgcn_fc1_update_t* upd, int Q, int I, int Hd, int C,
                  void* stream);

/* ---- data-parallel step tail: gradient allreduce fused with SGD over NVLink peer memory -------- */
/* Replaces DataParallel's gradient gather + optim.SGD(momentum).step() (pytorch_hcp_tgcn.py:164-169, 271-272).
 * Every rank allocates one region (tgcn_peer_region_bytes(n, nseg) bytes for n gradient elements in nseg tensors) with
 * tgcn_peer_alloc, exports its 64-byte IPC handle, and imports the peers' handles (same node; the handles travel
 * over any host channel, e.g. torch.distributed.all_gather_object).  tgcn_peer_allreduce_sgd then issues two
 * launches: pack this rank's gradients into its region and publish a step flag; wait for every rank's flag, read
 * all ranks' gradients over NVLink in rank order (bit-identical replicas), average, and apply
 * buf = momentum * buf + g; param -= lr * buf.  Capturable in a CUDA graph; bounded waits.  world <= 8. */
int tgcn_peer_alloc(int64_t bytes, void** ptr_out);
int tgcn_peer_free(void* ptr);
int tgcn_peer_export(void* ptr, unsigned char* handle64_host);
int tgcn_peer_import(const unsigned char* handle64_host, void** ptr_out);
int tgcn_peer_close(void* ptr);
int64_t tgcn_peer_region_bytes(int64_t n, int nseg);
int tgcn_peer_allreduce_sgd(void* const* regions_host, int world, int rank, const float* const* grads_host,
                            float* const* params_host, float* const* moms_host, const int64_t* numels_host,
                            int nseg, float lr, float momentum, unsigned int* state, void* stream);

/* ---- row-partitioned graph: halo rows read from the owners over NVLink (BASELINE configs[3]) -------- */
/* Each rank keeps its basis slabs in a region allocated with tgcn_peer_alloc and mapped by its peers (tgcn_peer_export /
 * _import); a 256-byte zero-initialised flag line sits at byte offset flag_off[r] of rank r's region.  After producing a
 * slab a rank calls tgcn_halo_signal (pushes its running step count to every rank's flag line); before the next
 * recursion step it calls tgcn_halo_pull, which waits for the owners' counts and gathers the halo rows -- owner[h],
 * row[h] = owning rank and its local row of halo row h -- from slab `slab_off[owner]` of the owner's region into `dst`
 * [n_halo, C] with 16-byte P2P loads.  Replaces the reference's single-device dense matmul for graphs that are
 * partitioned by rows (gcn_matmul.py:152-156).  `state`: 4 zero-initialised device uint32 per rank.  Capturable. */
int tgcn_halo_signal(void* const* regions_host, const int64_t* flag_off_host, int world, int rank,
                     unsigned int* state, void* stream);
int tgcn_halo_pull(void* const* regions_host, const int64_t* slab_off_host, const int64_t* flag_off_host, int world,
                   int rank, const int32_t* owner, const int32_t* row, int n_halo, int64_t C, float* dst,
                   unsigned int* state, void* stream);

/* ---- host-side graph preprocessing (CPU, no device work) ----------------------------------- */
/* One level of greedy heavy-edge (Graclus-normalised) matching: replaces the pure-Python loop
 * `metis_one_level` (gcn/coarsening.py:119-165) bit-exactly.  rr/cc/vv: COO triplets sorted by
 * row; visit: visiting order; weights: vertex degrees; cluster[n] receives the parent ids.
 * _f32/_f64 select the dtype the reference's numpy arithmetic would run in. */
int tgcn_pair_one_level_f32(const int64_t* rr_host, const int64_t* cc_host, const float* vv_host,
                            int64_t nnz, const int64_t* visit_host, const float* weights_host,
                            int64_t n, int32_t* cluster_host);
int tgcn_pair_one_level_f64(const int64_t* rr_host, const int64_t* cc_host, const double* vv_host,
                            int64_t nnz, const int64_t* visit_host, const double* weights_host,
                            int64_t n, int32_t* cluster_host);

#ifdef __cplusplus
}
#endif
#endif /* TGCN_B200_H */
