// Host-side (CPU) native helpers for the graph-preprocessing inputs of the hot path.
//
// `tgcn_pair_one_level_f32/_f64` is the sequential greedy heavy-edge matching of the reference's
// `metis_one_level` (gcn/coarsening.py:119-165), which there is a pure-Python loop over vertices
// and their CSR rows.  Integer outputs must be bit-exact, so the floating-point expression keeps
// the reference's evaluation order and precision: score = w_ij * (1/d_i + 1/d_j), evaluated in the
// dtype of the edge weights (numpy >= 2 keeps float32 for float32 inputs), strict '>' against the
// running maximum, first maximum wins.  Compiled without FMA contraction.
//
// This file contains no device code; it lives in the same shared library so that the Python side
// has one C-ABI to load.
#include <cstdint>
#include <vector>
#include "common.cuh"

namespace {

template <typename T>
int pair_one_level(const int64_t* rr, const int64_t* cc, const T* vv, int64_t nnz, const int64_t* visit,
                   const T* weights, int64_t n, int32_t* cluster) {
    if (n <= 0 || nnz <= 0) return TGCN_OK;
    std::vector<uint8_t> taken((size_t)n, 0);
    std::vector<int32_t> first((size_t)n, 0), length((size_t)n, 0);
    // Row extents exactly as the reference derives them (coarsening.py:134-139): the entry that
    // opens a new row is still counted for the previous slot, and slots only advance when the row
    // id grows (so an empty row shifts every later slot).  Reproduced on purpose.
    int64_t last = rr[0];
    int64_t slot = 0;
    for (int64_t e = 0; e < nnz; ++e) {
        length[(size_t)slot] += 1;
        if (rr[e] > last) {
            last = rr[e];
            if (slot + 1 >= n) return tgcn::set_error(TGCN_ERR_INVALID, "pair_one_level: row ids exceed n");
            first[(size_t)slot + 1] = (int32_t)e;
            slot += 1;
        }
    }
    for (int64_t i = 0; i < n; ++i) cluster[i] = 0;
    int32_t count = 0;
    for (int64_t pos = 0; pos < n; ++pos) {
        const int64_t v = visit[pos];
        if (v < 0 || v >= n) return tgcn::set_error(TGCN_ERR_INVALID, "pair_one_level: visit order out of range");
        if (taken[(size_t)v]) continue;
        taken[(size_t)v] = 1;
        T best_val = (T)0;
        int64_t best = -1;
        const int64_t base = first[(size_t)v];
        const int32_t len = length[(size_t)v];
        for (int32_t off = 0; off < len; ++off) {
            if (base + off >= nnz) return tgcn::set_error(TGCN_ERR_INVALID, "pair_one_level: row extent past nnz");
            const int64_t u = cc[base + off];
            T score;
            if (taken[(size_t)u]) {
                score = (T)0;
            } else {
                const volatile T inv_v = (T)1 / weights[v];
                const volatile T inv_u = (T)1 / weights[u];
                const volatile T s = inv_v + inv_u;
                score = vv[base + off] * s;
            }
            if (score > best_val) {
                best_val = score;
                best = u;
            }
        }
        cluster[v] = count;
        if (best > -1) {
            cluster[best] = count;
            taken[(size_t)best] = 1;
        }
        ++count;
    }
    return TGCN_OK;
}

}  // namespace

extern "C" int tgcn_pair_one_level_f32(const int64_t* rr_host, const int64_t* cc_host, const float* vv_host,
                                       int64_t nnz, const int64_t* visit_host, const float* weights_host,
                                       int64_t n, int32_t* cluster_host) {
    TGCN_REQUIRE(rr_host && cc_host && vv_host && visit_host && weights_host && cluster_host,
                 "tgcn_pair_one_level_f32: null pointer");
    return pair_one_level<float>(rr_host, cc_host, vv_host, nnz, visit_host, weights_host, n, cluster_host);
}

extern "C" int tgcn_pair_one_level_f64(const int64_t* rr_host, const int64_t* cc_host, const double* vv_host,
                                       int64_t nnz, const int64_t* visit_host, const double* weights_host,
                                       int64_t n, int32_t* cluster_host) {
    TGCN_REQUIRE(rr_host && cc_host && vv_host && visit_host && weights_host && cluster_host,
                 "tgcn_pair_one_level_f64: null pointer");
    return pair_one_level<double>(rr_host, cc_host, vv_host, nnz, visit_host, weights_host, n, cluster_host);
}
