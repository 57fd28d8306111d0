"""Host-side row-tile plan of the register-tiled SpMM (tgcn_rowtile_plan_host): a pure CPU function of the C-ABI.
The plan is replayed in numpy exactly as spmm_step_rtile_kernel walks it (per tile: distinct source rows in
ascending order, a dense R-vector of coefficients each) and compared with the sparse product it must equal
(gcn_matmul.py:154)."""
import numpy as np
import pytest
import scipy.sparse as sp

from tgcn_b200 import _lib


def _plan(m, R, pad=1):
    lib = _lib.load()
    m = m.tocsr(); m.sort_indices()
    rp = m.indptr.astype(np.int32); c = m.indices.astype(np.int32); v = m.data.astype(np.float32)
    n = m.shape[0]
    nt = (n + R - 1) // R
    tile_ptr = np.full(nt + 1, -1, np.int32)
    total = lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, R, pad, tile_ptr.ctypes.data, None, None)
    src = np.full(max(total, 1), -1, np.int32)
    w = np.full((max(total, 1), R), np.nan, np.float32)
    total2 = lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, R, pad, tile_ptr.ctypes.data,
                                        src.ctypes.data, w.ctypes.data)
    assert total == total2 == tile_ptr[nt]
    return rp, c, v, tile_ptr, src[:total], w[:total]


def _replay(tile_ptr, src, w, x, n, R):
    out = np.zeros((n, x.shape[1]), np.float64)
    for t in range(len(tile_ptr) - 1):
        for s in range(tile_ptr[t], tile_ptr[t + 1]):
            for q in range(R):
                row = t * R + q
                if row < n:
                    out[row] += np.float64(w[s, q]) * x[src[s]]
    return out


@pytest.mark.parametrize("R", [4, 8])
@pytest.mark.parametrize("n", [1, 7, 64, 203])
def test_plan_replay_equals_sparse_product(R, n):
    rng = np.random.default_rng(n * 10 + R)
    m = sp.random(n, n, density=min(1.0, 6.0 / n), random_state=n + R, format="lil", dtype=np.float32)
    if n > 10:
        m[3, :] = 0                      # an empty row
        m[5, ::2] = 0.5                  # a long row
    m = m.tocsr()
    rp, c, v, tile_ptr, src, w = _plan(m, R)
    x = rng.standard_normal((n, 5))
    ref = sp.csr_matrix((v.astype(np.float64), c, rp), shape=(n, n)) @ x
    np.testing.assert_allclose(_replay(tile_ptr, src, w, x, n, R), ref, rtol=1e-12, atol=1e-12)
    # structure: sources strictly ascending inside a tile, every stored coefficient is a CSR value or 0,
    # one (tile, source) pair per distinct column of the tile, padded rows of the last tile carry zeros
    for t in range(len(tile_ptr) - 1):
        s0, s1 = tile_ptr[t], tile_ptr[t + 1]
        assert np.all(np.diff(src[s0:s1]) > 0)
        r0, r1 = t * R, min(n, t * R + R)
        assert set(src[s0:s1].tolist()) == set(c[rp[r0]:rp[r1]].tolist())
        assert np.all(w[s0:s1, r1 - r0:] == 0)
    assert np.count_nonzero(w) == np.count_nonzero(v)


@pytest.mark.parametrize("R", [4, 8])
@pytest.mark.parametrize("pad", [2, 4, 8])
def test_padded_plan_is_the_unpadded_plan_plus_zero_coefficient_repeats(R, pad):
    n = 203
    rng = np.random.default_rng(pad)
    m = sp.random(n, n, density=0.03, random_state=7, format="lil", dtype=np.float32)
    m[8:16, :] = 0                                   # an empty tile (R = 8) / two empty tiles (R = 4): stays empty
    m = m.tocsr()
    rp, c, v, tp0, src0, w0 = _plan(m, R, 1)
    _, _, _, tp, src, w = _plan(m, R, pad)
    for t in range(len(tp) - 1):
        k = tp0[t + 1] - tp0[t]
        s0, s1 = tp[t], tp[t + 1]
        assert s1 - s0 == (k + pad - 1) // pad * pad
        assert np.array_equal(src[s0:s0 + k], src0[tp0[t]:tp0[t + 1]]) and np.array_equal(w[s0:s0 + k], w0[tp0[t]:tp0[t + 1]])
        if k:
            assert np.all(src[s0 + k:s1] == src[s0 + k - 1]) and np.all(w[s0 + k:s1] == 0)
    x = rng.standard_normal((n, 3))
    ref = sp.csr_matrix((v.astype(np.float64), c, rp), shape=(n, n)) @ x
    np.testing.assert_allclose(_replay(tp, src, w, x, n, R), ref, rtol=1e-12, atol=1e-12)


def test_plan_counts_shared_sources_once_and_rejects_bad_arguments():
    lib = _lib.load()
    # 8 rows that all read sources {0, 1}: one tile of 8 has 2 sources, two tiles of 4 have 2 each
    m = sp.csr_matrix(np.tile(np.array([[1.0, 2.0] + [0.0] * 6], np.float32), (8, 1)))
    for R, want in ((8, 2), (4, 4)):
        rp, c, v, tile_ptr, src, w = _plan(m, R)
        assert len(src) == want and np.all(w[:, :] == np.where(src[:, None] == 0, 1.0, 2.0))
    rp = m.indptr.astype(np.int32); c = m.indices.astype(np.int32); v = m.data.astype(np.float32)
    tp = np.zeros(9, np.int32)
    assert lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, 8, 3, 1, tp.ctypes.data, None, None) == -1
    assert lib.tgcn_rowtile_plan_host(None, c.ctypes.data, v.ctypes.data, 8, 8, 1, tp.ctypes.data, None, None) == -1
    assert lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, 8, 8, 0, tp.ctypes.data, None, None) == -1
    assert lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, 8, 8, 17, tp.ctypes.data, None, None) == -1
    src = np.zeros(16, np.int32)
    assert lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, 8, 8, 1, tp.ctypes.data, src.ctypes.data, None) == -1
    assert lib.tgcn_rowtile_plan_destroy(12345) != 0


def test_make_rowtile_plan_registers_and_filters_by_locality():
    """csr.make_rowtile_plan: host plan + registration keyed by the `col` array (a CPU tensor stands in for the device
    array here -- nothing is launched); operands without locality are refused by min_gain."""
    import torch
    from tgcn_b200.csr import make_rowtile_plan
    lib = _lib.load()
    # a banded matrix in natural order: neighbouring rows share sources
    n = 64
    band = sp.diags([1.0, 2.0, 3.0, 2.0, 1.0], [-2, -1, 0, 1, 2], shape=(n, n), format="csr", dtype=np.float32)
    col_t = torch.from_numpy(band.indices.astype(np.int32))
    made = make_rowtile_plan(band.indptr, band.indices, band.data, n, 8, col_t, min_gain=1.5)
    assert made is not None
    h, arrays, stats = made
    assert stats["rows_per_tile"] == 8 and stats["sources"] == stats["entries"] == int(arrays[0][-1]) and stats["gain"] > 2.5
    padded = make_rowtile_plan(band.indptr, band.indices, band.data, n, 8, col_t, min_gain=1.5, pad=8)
    assert padded[2]["sources"] == stats["sources"] and padded[2]["entries"] % 8 == 0 and padded[2]["entries"] == int(padded[1][0][-1])
    assert padded[2]["gain"] == stats["gain"] and lib.tgcn_rowtile_plan_destroy(padded[0]) == 0
    assert arrays[2].shape == (stats["sources"], 8) and int(arrays[1].max()) < n
    assert lib.tgcn_rowtile_plan_destroy(h) == 0
    assert lib.tgcn_rowtile_plan_destroy(h) != 0                      # already dropped
    # a random permutation destroys the locality: one source per entry, gain ~1 -> no plan
    rng = np.random.default_rng(0)
    p = rng.permutation(n)
    scattered = sp.csr_matrix(sp.random(n, n, density=0.02, random_state=1, dtype=np.float32))[p][:, p].tocsr()
    col_s = torch.from_numpy(scattered.indices.astype(np.int32))
    assert make_rowtile_plan(scattered.indptr, scattered.indices, scattered.data, n, 4, col_s, min_gain=1.5) is None
    # empty operand
    assert make_rowtile_plan(np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), 4, 4, col_s) is None


@pytest.mark.parametrize("world", [2, 3])
def test_plan_of_a_row_partition_replays_to_the_global_product(world):
    """Config 4: every rank builds the row-tile plan of ITS rows over the extended (own + halo) column space
    (tgcn_b200.parallel.RowPartition); replayed on the extended slab it gives the global product's rows."""
    from tgcn_b200.parallel import RowPartition
    from tgcn_b200 import workloads as wl
    L, _ = wl.random_geometric(n=3001, mean_degree=10.0, seed=3)
    L = sp.csr_matrix(L)
    n = L.shape[0]
    rng = np.random.default_rng(world)
    x = rng.standard_normal((n, 4))
    ref = sp.csr_matrix(L, dtype=np.float64) @ x
    lib = _lib.load()
    for rank in range(world):
        part = RowPartition(L, rank, world)
        x_ext = np.concatenate([x[part.lo:part.hi], x[part.halo_ids]])          # owned rows, then halo rows by global id
        R = 4
        nt = (part.n_own + R - 1) // R
        tp = np.zeros(nt + 1, np.int32)
        rp, c, v = part.rowptr, part.col, part.val
        total = lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, part.n_own, R, 1, tp.ctypes.data, None, None)
        src = np.zeros(total, np.int32); w = np.zeros((total, R), np.float32)
        lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, part.n_own, R, 1, tp.ctypes.data, src.ctypes.data,
                                   w.ctypes.data)
        assert src.max() < part.n_own + part.n_halo and (world == 1 or src.max() >= part.n_own)   # halo rows are gathered
        got = _replay(tp, src, w, x_ext, part.n_own, R)
        np.testing.assert_allclose(got, ref[part.lo:part.hi], rtol=1e-6, atol=1e-7)    # fp32 coefficients vs fp64 product of fp32 values
        assert total < c.size / 1.3                                                      # the strip+Morton order shares sources
