// Gradient allreduce fused with the SGD-momentum update over NVLink peer memory.
//
// The data-parallel step ends with "average the gradients over the ranks, then update the parameters"
// (reference: torch.nn.DataParallel's gather + optim.SGD.step, pytorch_hcp_tgcn.py:164-169,271-272).  With NCCL that
// is a multi-tensor copy, an allreduce kernel and an optimizer kernel, ~35 us of mostly launch/latency at the
// 1.4 MB gradient size of the parcellation model.  Here every rank owns one cudaMalloc'ed region that its peers
// map through CUDA IPC:  [ flat gradients, buffer 0 | buffer 1 | ready flags, one per source rank ].
//   peer_pack_kernel   : copies this rank's gradients into flat[step & 1], then the last block to finish PUSHES
//                        `ready = step + 1` into slot [rank] of EVERY rank's flag line (system-scope fence, then one
//                        remote release store per peer, issued by different threads);
//   peer_reduce_sgd_kernel : polls its OWN flag line (local memory, one thread per source rank -- no NVLink round
//                        trip per poll and no serial walk over the peers) until every slot shows the step, then each
//                        thread reads ITS elements from
//                        all ranks' buffers straight over NVLink (P2P loads), sums them in rank order -- the same
//                        order on every rank, so the replicas stay bit-identical --, scales by 1/world and applies
//                        buf = momentum * buf + g;  param -= lr * buf.  No intermediate averaged-gradient tensor.
// Two-shot variant (large gradients / many ranks, chosen identically on every rank from n and world): the one-shot
// kernel makes every rank pull ALL ranks' copies of the WHOLE gradient ((world-1) * n elements over NVLink per rank).
//   peer_rs_kernel     : reduce-scatter -- rank r sums only slice r of the flat index space from all ranks' buffers
//                        (rank order), scales by 1/world and PUSHES the averaged slice into every rank's `avg` buffer
//                        (P2P stores), then publishes a second flag;
//   peer_apply_kernel  : waits for every rank's second flag and applies the SGD update from its local `avg` buffer.
// NVLink traffic per rank drops to 2 (world-1)/world * n at the price of a second flag round trip.
// Buffers alternate with the step parity, so a rank may start writing step s+2 only after all peers have posted
// ready(s+1), i.e. after they finished reading step s: one flag exchange per step is the only synchronisation.
// Waits are bounded (TGCN_PEER_TIMEOUT_S, default 120 s, then trap) so a lost peer cannot hang the device for ever.  Step numbers live in device memory,
// so the two launches are CUDA-graph capturable.
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace tgcn {

constexpr int kPeerMaxSeg = 24;

struct PeerSegs {
    const float* grad[kPeerMaxSeg];
    float* param[kPeerMaxSeg];
    float* mom[kPeerMaxSeg];
    int64_t off[kPeerMaxSeg + 1];     // offsets into the flat buffer (multiples of 4 elements)
    int64_t len[kPeerMaxSeg];         // elements of each tensor
    int nseg;
};

struct PeerRanks {
    const float* flat[kPeerMaxWorld];           // each rank's region base
    unsigned int* flag[kPeerMaxWorld];          // each rank's flag line (kPeerMaxWorld slots, slot s written by rank s)
    float* avg[kPeerMaxWorld];                  // two-shot: each rank's averaged-gradient buffers (2 x n)
    unsigned int* flag2[kPeerMaxWorld];         // two-shot: second flag line (averaged slices pushed)
    int world, rank;
    unsigned long long timeout_ns;
};

// Segment offsets are multiples of 4 elements, so every tensor is walked as float4 (plus a scalar tail) and ONE
// grid-wide index space covers all tensors: one float4 per thread, all loads of a thread in flight at once -- the
// remote (NVLink) round trip is paid once per thread, not once per element.
__device__ __forceinline__ int find_seg(const PeerSegs& segs, int64_t e) {
    int s = 0;
#pragma unroll 1
    while (s + 1 < segs.nseg && segs.off[s + 1] <= e) ++s;
    return s;
}

__global__ void __launch_bounds__(256)
peer_pack_kernel(const PeerSegs segs, const PeerRanks ranks, float* flat_base, int64_t n, const unsigned int* step_ctr,
                 unsigned int* done_blocks) {
    __shared__ int s_last;
    const unsigned int step = *step_ctr;
    float* flat = flat_base + (int64_t)(step & 1u) * n;
    for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < n / 4; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = 4 * i4;
        const int s = find_seg(segs, e);
        const int64_t loc = e - segs.off[s], len = segs.len[s];
        const float* src = segs.grad[s];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src) {
            if (loc + 4 <= len) v = *reinterpret_cast<const float4*>(src + loc);
            else {
                if (loc + 0 < len) v.x = src[loc + 0];
                if (loc + 1 < len) v.y = src[loc + 1];
                if (loc + 2 < len) v.z = src[loc + 2];
            }
        }
        *reinterpret_cast<float4*>(flat + e) = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) *done_blocks = 0u;
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < ranks.world) {   // last block: every block's stores are fenced; publish to all ranks
        __threadfence_system();
        st_release_sys(ranks.flag[threadIdx.x] + ranks.rank, step + 1u);
    }
}

__device__ __forceinline__ float4 ldcv4(const float* p) { return ld_volatile_f4(p); }

__global__ void __launch_bounds__(256)
peer_reduce_sgd_kernel(const PeerSegs segs, const PeerRanks ranks, int64_t n, float lr, float momentum,
                       unsigned int* step_ctr, unsigned int* done_blocks) {
    const unsigned int step = *step_ctr;
    if ((int)threadIdx.x < ranks.world)                // own flag line: slot r is pushed by rank r's pack kernel
        peer_wait_flag(ranks.flag[ranks.rank] + threadIdx.x, step + 1u, ranks.timeout_ns);
    __syncthreads();
    const int64_t par = (int64_t)(step & 1u) * n;
    const float inv = 1.0f / (float)ranks.world;
    for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < n / 4; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = 4 * i4;
        float4 gr[kPeerMaxWorld];
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < ranks.world) gr[r] = ldcv4(ranks.flat[r] + par + e);          // all ranks' loads in flight together
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < ranks.world) { g.x += gr[r].x; g.y += gr[r].y; g.z += gr[r].z; g.w += gr[r].w; }   // rank order: same on every rank
        const int s = find_seg(segs, e);
        const int64_t loc = e - segs.off[s], len = segs.len[s];
        float* prm = segs.param[s] + loc;
        float* mom = segs.mom[s] + loc;
        const float gg[4] = {g.x * inv, g.y * inv, g.z * inv, g.w * inv};
        if (loc + 4 <= len) {
            float4 m = *reinterpret_cast<float4*>(mom), w = *reinterpret_cast<float4*>(prm);
            m.x = fmaf(momentum, m.x, gg[0]); m.y = fmaf(momentum, m.y, gg[1]);
            m.z = fmaf(momentum, m.z, gg[2]); m.w = fmaf(momentum, m.w, gg[3]);
            w.x = fmaf(-lr, m.x, w.x); w.y = fmaf(-lr, m.y, w.y); w.z = fmaf(-lr, m.z, w.z); w.w = fmaf(-lr, m.w, w.w);
            *reinterpret_cast<float4*>(mom) = m;
            *reinterpret_cast<float4*>(prm) = w;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (loc + c < len) {
                    const float bcur = fmaf(momentum, mom[c], gg[c]);
                    mom[c] = bcur;
                    prm[c] = fmaf(-lr, bcur, prm[c]);
                }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        if (prev == gridDim.x - 1) {
            *done_blocks = 0u;
            *step_ctr = step + 1u;
        }
    }
}

__device__ __forceinline__ void sgd_apply4(const PeerSegs& segs, int64_t e, const float (&gg)[4], float lr, float momentum) {
    const int s = find_seg(segs, e);
    const int64_t loc = e - segs.off[s], len = segs.len[s];
    float* prm = segs.param[s] + loc;
    float* mom = segs.mom[s] + loc;
    if (loc + 4 <= len) {
        float4 m = *reinterpret_cast<float4*>(mom), w = *reinterpret_cast<float4*>(prm);
        m.x = fmaf(momentum, m.x, gg[0]); m.y = fmaf(momentum, m.y, gg[1]);
        m.z = fmaf(momentum, m.z, gg[2]); m.w = fmaf(momentum, m.w, gg[3]);
        w.x = fmaf(-lr, m.x, w.x); w.y = fmaf(-lr, m.y, w.y); w.z = fmaf(-lr, m.z, w.z); w.w = fmaf(-lr, m.w, w.w);
        *reinterpret_cast<float4*>(mom) = m;
        *reinterpret_cast<float4*>(prm) = w;
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (loc + c < len) {
                const float bcur = fmaf(momentum, mom[c], gg[c]);
                mom[c] = bcur;
                prm[c] = fmaf(-lr, bcur, prm[c]);
            }
    }
}

// two-shot, first half: rank r reduces float4 indices [r * per, (r + 1) * per) and pushes the average to every rank
__global__ void __launch_bounds__(256)
peer_rs_kernel(const PeerRanks ranks, int64_t n, const unsigned int* step_ctr, unsigned int* done_blocks) {
    __shared__ int s_last;
    const unsigned int step = *step_ctr;
    if ((int)threadIdx.x < ranks.world)
        peer_wait_flag(ranks.flag[ranks.rank] + threadIdx.x, step + 1u, ranks.timeout_ns);
    __syncthreads();
    const int64_t par = (int64_t)(step & 1u) * n;
    const float inv = 1.0f / (float)ranks.world;
    const int64_t n4 = n / 4, per = (n4 + ranks.world - 1) / ranks.world;
    const int64_t lo = (int64_t)ranks.rank * per, hi = lo + per < n4 ? lo + per : n4;
    for (int64_t i4 = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < hi; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = 4 * i4;
        float4 gr[kPeerMaxWorld];
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < ranks.world) gr[r] = ldcv4(ranks.flat[r] + par + e);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < ranks.world) { g.x += gr[r].x; g.y += gr[r].y; g.z += gr[r].z; g.w += gr[r].w; }
        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
#pragma unroll
        for (int r = 0; r < kPeerMaxWorld; ++r)
            if (r < ranks.world) *reinterpret_cast<float4*>(ranks.avg[r] + par + e) = g;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) *done_blocks = 0u;
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < ranks.world) {
        __threadfence_system();
        st_release_sys(ranks.flag2[threadIdx.x] + ranks.rank, step + 1u);
    }
}

// two-shot, second half: every slice has arrived in the local `avg` buffer -> SGD with momentum on all parameters
__global__ void __launch_bounds__(256)
peer_apply_kernel(const PeerSegs segs, const PeerRanks ranks, int64_t n, float lr, float momentum, unsigned int* step_ctr,
                  unsigned int* done_blocks) {
    const unsigned int step = *step_ctr;
    if ((int)threadIdx.x < ranks.world)
        peer_wait_flag(ranks.flag2[ranks.rank] + threadIdx.x, step + 1u, ranks.timeout_ns);
    __syncthreads();
    const float* avg = ranks.avg[ranks.rank] + (int64_t)(step & 1u) * n;
    for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < n / 4; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = 4 * i4;
        const float4 g = ldcv4(avg + e);
        const float gg[4] = {g.x, g.y, g.z, g.w};
        sgd_apply4(segs, e, gg, lr, momentum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        if (prev == gridDim.x - 1) {
            *done_blocks = 0u;
            *step_ctr = step + 1u;
        }
    }
}

// world == 1: no exchange; the same update read straight from the gradient tensors (one launch)
__global__ void __launch_bounds__(256)
sgd_direct_kernel(const PeerSegs segs, int64_t n, float lr, float momentum, unsigned int* step_ctr, unsigned int* done_blocks) {
    for (int64_t i4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i4 < n / 4; i4 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = 4 * i4;
        const int s = find_seg(segs, e);
        const int64_t loc = e - segs.off[s], len = segs.len[s];
        const float* src = segs.grad[s];
        float* prm = segs.param[s] + loc;
        float* mom = segs.mom[s] + loc;
        if (loc + 4 <= len) {
            const float4 g = src ? *reinterpret_cast<const float4*>(src + loc) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 m = *reinterpret_cast<float4*>(mom), w = *reinterpret_cast<float4*>(prm);
            m.x = fmaf(momentum, m.x, g.x); m.y = fmaf(momentum, m.y, g.y);
            m.z = fmaf(momentum, m.z, g.z); m.w = fmaf(momentum, m.w, g.w);
            w.x = fmaf(-lr, m.x, w.x); w.y = fmaf(-lr, m.y, w.y); w.z = fmaf(-lr, m.z, w.z); w.w = fmaf(-lr, m.w, w.w);
            *reinterpret_cast<float4*>(mom) = m;
            *reinterpret_cast<float4*>(prm) = w;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (loc + c < len) {
                    const float bcur = fmaf(momentum, mom[c], src ? src[loc + c] : 0.f);
                    mom[c] = bcur;
                    prm[c] = fmaf(-lr, bcur, prm[c]);
                }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        if (prev == gridDim.x - 1) { *done_blocks = 0u; *step_ctr = *step_ctr + 1u; }
    }
}

// ---- halo rows of a row-partitioned slab, read straight from the owners' memory ----
// Every rank keeps its basis slabs in an IPC-mapped region.  Before recursion step j a rank needs slab j-1's rows of the
// vertices its rows reference but other ranks own.  halo_signal_kernel (one thread per peer) publishes "my slab is
// complete" by pushing a step count into slot [rank] of every rank's flag line; halo_pull_kernel waits for the owners'
// counts and copies the rows with P2P loads (16-byte, coalesced along the row) into the local extended slab.
// Replaces index_select + NCCL send/recv per step (host-driven, several launches) by two launches without host sync.
__global__ void halo_signal_kernel(PeerRanks ranks, unsigned int* state) {
    const unsigned int v = state[0] + 1u;
    __threadfence_system();
    if ((int)threadIdx.x < ranks.world) st_release_sys(ranks.flag[threadIdx.x] + ranks.rank, v);
    __syncthreads();
    if (threadIdx.x == 0) state[0] = v;
}

struct HaloSrc { const float* slab[kPeerMaxWorld]; };      // slab j-1 of every rank, as mapped in this process

__global__ void __launch_bounds__(256)
halo_pull_kernel(const HaloSrc src, const PeerRanks ranks, const int* __restrict__ owner, const int* __restrict__ row,
                 int n_halo, int V, float4* __restrict__ dst, unsigned int* state, unsigned int* done_blocks) {
    const unsigned int expect = state[1] + 1u;
    if ((int)threadIdx.x < ranks.world && (int)threadIdx.x != ranks.rank)
        peer_wait_flag(ranks.flag[ranks.rank] + threadIdx.x, expect, ranks.timeout_ns);
    __syncthreads();
    const int64_t total = (int64_t)n_halo * V;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(i / V), v = (int)(i - (int64_t)h * V);
        const float* base = src.slab[__ldg(owner + h)] + ((int64_t)__ldg(row + h) * V + v) * 4;
        dst[i] = ld_volatile_f4(base);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done_blocks, 1u);
        if (prev == gridDim.x - 1) { *done_blocks = 0u; state[1] = expect; }
    }
}

}  // namespace tgcn

using namespace tgcn;

// regions_host[r]: rank r's slab region as mapped here; flag_off_host[r]: byte offset of its flag line (256 B, zeroed);
// state: 4 device uint32 of this rank (zeroed): [0] signals sent, [1] pulls done, [2] block counter
extern "C" int tgcn_halo_signal(void* const* regions_host, const int64_t* flag_off_host, int world, int rank,
                                unsigned int* state, void* stream) {
    TGCN_REQUIRE(regions_host && flag_off_host && state, "tgcn_halo_signal: null pointer");
    TGCN_SUPPORTED(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "tgcn_halo_signal: world %d rank %d", world, rank);
    PeerRanks ranks{};
    ranks.world = world; ranks.rank = rank; ranks.timeout_ns = peer_timeout_ns();
    for (int r = 0; r < world; ++r)
        ranks.flag[r] = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(regions_host[r]) + flag_off_host[r]);
    halo_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(ranks, state);
    TGCN_LAUNCH_CHECK("halo_signal");
    return TGCN_OK;
}

// dst[h, :] = slab_{owner[h]}[row[h], :] for the n_halo halo rows (C floats each, C % 4 == 0); slab_off_host[r]: element
// offset of the wanted slab inside rank r's region.  Waits until every peer has signalled once more than this rank has pulled.
extern "C" int tgcn_halo_pull(void* const* regions_host, const int64_t* slab_off_host, const int64_t* flag_off_host, int world,
                              int rank, const int32_t* owner, const int32_t* row, int n_halo, int64_t C, float* dst,
                              unsigned int* state, void* stream) {
    TGCN_REQUIRE(regions_host && slab_off_host && flag_off_host && state, "tgcn_halo_pull: null pointer");
    TGCN_SUPPORTED(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "tgcn_halo_pull: world %d rank %d", world, rank);
    TGCN_REQUIRE(C % 4 == 0 && aligned16(dst), "tgcn_halo_pull: rows must be whole 16-byte vectors");
    TGCN_REQUIRE(n_halo == 0 || (owner && row && dst), "tgcn_halo_pull: null halo tables");
    HaloSrc src{};
    PeerRanks ranks{};
    ranks.world = world; ranks.rank = rank; ranks.timeout_ns = peer_timeout_ns();
    for (int r = 0; r < world; ++r) {
        TGCN_REQUIRE(regions_host[r] && (slab_off_host[r] % 4) == 0, "tgcn_halo_pull: bad region / slab offset for rank %d", r);
        src.slab[r] = reinterpret_cast<const float*>(regions_host[r]) + slab_off_host[r];
        ranks.flag[r] = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(regions_host[r]) + flag_off_host[r]);
    }
    const int V = (int)(C / 4);
    int64_t blocks = ceil_div((int64_t)n_halo * V, 256);
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)kNumSMs * 4) blocks = (int64_t)kNumSMs * 4;
    halo_pull_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, ranks, owner, row, n_halo, V, reinterpret_cast<float4*>(dst),
                                                                     state, state + 2);
    TGCN_LAUNCH_CHECK("halo_pull");
    return TGCN_OK;
}

// ---- peer regions -------------------------------------------------------------------------------------------
extern "C" int tgcn_peer_alloc(int64_t bytes, void** ptr_out) {
    TGCN_REQUIRE(bytes > 0 && ptr_out, "tgcn_peer_alloc: bad arguments");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_peer_alloc: %s", cudaGetErrorString(e));
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_peer_alloc: %s", cudaGetErrorString(e));
    *ptr_out = p;
    return TGCN_OK;
}

extern "C" int tgcn_peer_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_peer_free: %s", cudaGetErrorString(e));
    return TGCN_OK;
}

// 64-byte opaque handle another process on the same node can open (cudaIpcGetMemHandle)
extern "C" int tgcn_peer_export(void* ptr, unsigned char* handle64_host) {
    TGCN_REQUIRE(ptr && handle64_host, "tgcn_peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_peer_export: %s", cudaGetErrorString(e));
    memcpy(handle64_host, &h, 64);
    return TGCN_OK;
}

extern "C" int tgcn_peer_import(const unsigned char* handle64_host, void** ptr_out) {
    TGCN_REQUIRE(handle64_host && ptr_out, "tgcn_peer_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_peer_import: %s", cudaGetErrorString(e));
    *ptr_out = p;
    return TGCN_OK;
}

extern "C" int tgcn_peer_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) return set_error(TGCN_ERR_CUDA, "tgcn_peer_close: %s", cudaGetErrorString(e));
    return TGCN_OK;
}

// bytes of one rank's region for n gradient elements in nseg tensors (each padded to 4 elements): two flat
// gradient buffers, two averaged-gradient buffers (two-shot variant) and the two flag lines
static int64_t peer_flat_elems(int64_t n, int nseg) { return n + 3 * (int64_t)nseg; }
extern "C" int64_t tgcn_peer_region_bytes(int64_t n, int nseg) {
    return (n < 0 || nseg < 0) ? 0 : 4 * peer_flat_elems(n, nseg) * (int64_t)sizeof(float) + 512;
}

namespace tgcn {
// Region layout for a flat buffer of n (padded) elements: [flat0 | flat1 | flags (256 B) | flags2 (256 B) | avg0 | avg1].
// The flag lines sit right behind the two flat buffers, so a region that is only ever packed and read (bighead.cu)
// needs 2 n floats + 512 bytes.
static void peer_fill_ranks(PeerRanks& ranks, void* const* regions_host, int world, int rank, int64_t n) {
    ranks.world = world; ranks.rank = rank; ranks.timeout_ns = peer_timeout_ns();
    for (int r = 0; r < world; ++r) {
        char* base = reinterpret_cast<char*>(regions_host[r]);
        ranks.flat[r] = reinterpret_cast<const float*>(base);
        ranks.flag[r] = reinterpret_cast<unsigned int*>(base + 2 * n * sizeof(float));
        ranks.flag2[r] = reinterpret_cast<unsigned int*>(base + 2 * n * sizeof(float) + 256);
        ranks.avg[r] = reinterpret_cast<float*>(base + 2 * n * sizeof(float) + 512);
    }
}

// Copy `nseg` tensors of this rank into flat[step & 1] of its region and publish the step flag to every rank
// (state[0] = step counter, state[1] = block counter).  Used by the gradient exchange below and by bighead.cu.
int peer_pack_launch(void* const* regions_host, int world, int rank, const float* const* srcs_host, const int64_t* numels_host,
                     int nseg, unsigned int* state, cudaStream_t st) {
    TGCN_SUPPORTED(nseg >= 1 && nseg <= kPeerMaxSeg, "peer_pack: %d tensors (max %d)", nseg, kPeerMaxSeg);
    PeerSegs segs{};
    segs.nseg = nseg;
    int64_t n = 0;
    for (int s = 0; s < nseg; ++s) {
        TGCN_REQUIRE(!srcs_host[s] || aligned16(srcs_host[s]), "peer_pack: tensor %d is not 16-byte aligned", s);
        segs.grad[s] = srcs_host[s];
        segs.off[s] = n;
        segs.len[s] = numels_host[s];
        n += (numels_host[s] + 3) & ~(int64_t)3;
    }
    segs.off[nseg] = n;
    PeerRanks ranks{};
    for (int r = 0; r < world; ++r) TGCN_REQUIRE(regions_host[r], "peer_pack: null region for rank %d", r);
    peer_fill_ranks(ranks, regions_host, world, rank, n);
    const int blocks = (int)min64(ceil_div(n > 0 ? n / 4 : 1, 256), (int64_t)kNumSMs * 8);
    peer_pack_kernel<<<blocks, 256, 0, st>>>(segs, ranks, reinterpret_cast<float*>(regions_host[rank]), n, state, state + 1);
    TGCN_LAUNCH_CHECK("peer_pack");
    return TGCN_OK;
}
}  // namespace tgcn

// One data-parallel optimizer step.  regions_host[r]: rank r's region base as seen from THIS process (own region
// for r == rank); grads/params/moms/numels: nseg host arrays describing the parameter tensors in a fixed order
// (identical on every rank); state: device buffer of 4 uint32 owned by this rank (step counter, block counters).
extern "C" int tgcn_peer_allreduce_sgd(void* const* regions_host, int world, int rank, const float* const* grads_host,
                                       float* const* params_host, float* const* moms_host, const int64_t* numels_host,
                                       int nseg, float lr, float momentum, unsigned int* state, void* stream) {
    TGCN_REQUIRE(regions_host && grads_host && params_host && moms_host && numels_host && state, "tgcn_peer_allreduce_sgd: null pointer");
    TGCN_SUPPORTED(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "tgcn_peer_allreduce_sgd: world %d rank %d", world, rank);
    TGCN_SUPPORTED(nseg >= 1 && nseg <= kPeerMaxSeg, "tgcn_peer_allreduce_sgd: %d parameter tensors (max %d)", nseg, kPeerMaxSeg);
    PeerSegs segs{};
    segs.nseg = nseg;
    int64_t n = 0;
    for (int s = 0; s < nseg; ++s) {
        TGCN_REQUIRE(params_host[s] && moms_host[s] && numels_host[s] >= 0, "tgcn_peer_allreduce_sgd: bad segment %d", s);
        segs.grad[s] = grads_host[s]; segs.param[s] = params_host[s]; segs.mom[s] = moms_host[s];
        segs.off[s] = n;
        segs.len[s] = numels_host[s];
        n += (numels_host[s] + 3) & ~(int64_t)3;
        TGCN_REQUIRE(aligned16(params_host[s]) && aligned16(moms_host[s]) && (!grads_host[s] || aligned16(grads_host[s])),
                     "tgcn_peer_allreduce_sgd: segment %d is not 16-byte aligned", s);
    }
    segs.off[nseg] = n;
    cudaStream_t st = as_stream(stream);
    const int blocks = (int)min64(ceil_div(n > 0 ? n / 4 : 1, 256), (int64_t)kNumSMs * 8);
    if (world == 1) {
        sgd_direct_kernel<<<blocks, 256, 0, st>>>(segs, n, lr, momentum, state, state + 2);
        TGCN_LAUNCH_CHECK("sgd_direct");
        return TGCN_OK;
    }
    PeerRanks ranks{};
    for (int r = 0; r < world; ++r) TGCN_REQUIRE(regions_host[r], "tgcn_peer_allreduce_sgd: null region for rank %d", r);
    peer_fill_ranks(ranks, regions_host, world, rank, n);
    float* my_flat = reinterpret_cast<float*>(regions_host[rank]);
    peer_pack_kernel<<<blocks, 256, 0, st>>>(segs, ranks, my_flat, n, state, state + 1);
    TGCN_LAUNCH_CHECK("peer_pack");
    // one-shot (every rank reads everything: (world-1) n elements in, one flag round trip) or two-shot (reduce-scatter +
    // pushed all-gather: 2 (world-1)/world n elements, two flag round trips).  Decided from (n, world) only, so every
    // rank takes the same branch; TGCN_PEER_TWOSHOT=0/1 forces it.
    static const int forced = [] { const char* e = getenv("TGCN_PEER_TWOSHOT"); return e ? atoi(e) : -1; }();
    const bool twoshot = forced >= 0 ? forced != 0 : (world >= 4 && n * (int64_t)sizeof(float) >= ((int64_t)2 << 20));
    if (!twoshot) {
        peer_reduce_sgd_kernel<<<blocks, 256, 0, st>>>(segs, ranks, n, lr, momentum, state, state + 2);
        TGCN_LAUNCH_CHECK("peer_reduce_sgd");
        return TGCN_OK;
    }
    const int rs_blocks = (int)min64(ceil_div(ceil_div(n / 4, world), 256), (int64_t)kNumSMs * 4);
    peer_rs_kernel<<<rs_blocks > 0 ? rs_blocks : 1, 256, 0, st>>>(ranks, n, state, state + 3);
    TGCN_LAUNCH_CHECK("peer_rs");
    peer_apply_kernel<<<blocks, 256, 0, st>>>(segs, ranks, n, lr, momentum, state, state + 2);
    TGCN_LAUNCH_CHECK("peer_apply");
    return TGCN_OK;
}
