"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there).

* cyclic-shift operand: L = row-shift by one => L^j x is x rolled by j: closed form for the whole layer
* linearity in x and in W
* K=1 reduces to a per-vertex dense contraction
* data-parallel invariance: batch halves give the same outputs / summed weight gradients
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def shift_operator(N, device):
    idx = torch.arange(N, device=device)
    i = torch.stack([idx, (idx + 1) % N])
    return torch.sparse_coo_tensor(i, torch.ones(N, device=device), (N, N))


def reference_closed_form(x, W, b, K):
    """x[Q,N,H], L = shift: (L x)[n] = x[n+1].  Reference stack: Xt_k = 2 roll_k(x) - Xt_{k-2}."""
    P = [torch.roll(x, -j, dims=1) for j in range(K)]
    Xt = []
    for k in range(K):
        Xt.append(P[k] if k < 2 else 2 * P[k] - Xt[k - 2])
    out = sum(torch.einsum("qnh,hg->qng", Xt[k].double(), W[k, :, 0, :].double()) for k in range(K))
    return out + b.double()


@pytest.mark.parametrize("N,Q,H,G,K", [(41552, 8, 30, 32, 10), (992, 100, 12, 15, 10), (448, 64, 15, 32, 10)])
def test_full_size_layer_against_closed_form(N, Q, H, G, K):
    from tgcn_b200.nn.gcn import TGCNCheb_H
    torch.manual_seed(1)
    L = shift_operator(N, "cuda")
    lay = TGCNCheb_H(L, 1, G, K, H).cuda()
    x = torch.randn(Q, N, H, device="cuda")
    out = lay(x)
    ref = reference_closed_form(x, lay.weight.detach(), lay.bias.detach(), K)
    err = (out.double() - ref).abs().max() / ref.abs().max()
    assert err < 1e-4, float(err)
    # gradient of sum(out * dout) w.r.t. W against the closed form
    dout = torch.randn_like(out)
    out.backward(dout)
    P = [torch.roll(x, -j, dims=1) for j in range(K)]
    Xt = []
    for k in range(K):
        Xt.append(P[k] if k < 2 else 2 * P[k] - Xt[k - 2])
    dW = torch.stack([torch.einsum("qnh,qng->hg", Xt[k].double(), dout.double()) for k in range(K)])
    got = lay.weight.grad[:, :, 0, :].double()
    assert ((got - dW).abs().max() / dW.abs().max()) < 1e-4
    db = dout.double().sum(0, keepdim=True)
    assert ((lay.bias.grad.double() - db).abs().max() / db.abs().max()) < 1e-4


def test_full_size_second_layer_dx_is_adjoint():
    """<layer(x) - bias, y> == <x, dx(y)> (the backward really is the adjoint of the forward)."""
    from tgcn_b200.nn.gcn import GCNCheb
    torch.manual_seed(2)
    N, Q, F, G, K = 10388, 8, 32, 64, 10
    idx = torch.arange(N)
    i = torch.cat([torch.stack([idx, (idx + 1) % N]), torch.stack([idx, (idx + 7) % N]), torch.stack([idx, (idx * 3 + 1) % N])], 1)
    v = torch.rand(i.shape[1]) * 0.3
    L = torch.sparse_coo_tensor(i, v, (N, N)).coalesce().cuda()
    lay = GCNCheb(L, F, G, K, bias=False).cuda()
    x = torch.randn(Q, N, F, device="cuda", requires_grad=True)
    y = torch.randn(Q, N, G, device="cuda")
    out = lay(x)
    lhs = (out.double() * y.double()).sum()
    out.backward(y)
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-4
    # linearity in x
    x2 = torch.randn_like(x)
    with torch.no_grad():
        a = lay(2.0 * x.detach() - 3.0 * x2)
        b = 2.0 * lay(x.detach()) - 3.0 * lay(x2)
    assert float((a - b).abs().max() / b.abs().max()) < 1e-4


def test_k1_is_a_plain_contraction():
    from tgcn_b200.nn.gcn import GCNCheb
    torch.manual_seed(3)
    N, Q, F, G = 5000, 4, 16, 24
    lay = GCNCheb(shift_operator(N, "cuda"), F, G, 1).cuda()
    x = torch.randn(Q, N, F, device="cuda")
    ref = torch.einsum("qnf,fg->qng", x.double(), lay.weight[0].double()) + lay.bias.double()
    assert float((lay(x).double() - ref).abs().max() / ref.abs().max()) < 1e-5


def test_batch_sharding_invariance():
    """Data-parallel contract: per-sample outputs do not depend on the batch they ride in,
    and weight gradients of shards add up to the full-batch gradient."""
    from tgcn_b200.nn.gcn import TGCNCheb_H
    torch.manual_seed(4)
    N, Q, H, G, K = 2000, 16, 15, 32, 6
    lay = TGCNCheb_H(shift_operator(N, "cuda"), 1, G, K, H).cuda()
    x = torch.randn(Q, N, H, device="cuda")
    dout = torch.randn(Q, N, G, device="cuda")
    full = lay(x); full.backward(dout)
    gW, gb = lay.weight.grad.clone(), lay.bias.grad.clone()
    lay.zero_grad()
    parts = []
    for s in (slice(0, 8), slice(8, 16)):
        o = lay(x[s]); o.backward(dout[s]); parts.append(o.detach())
    assert torch.equal(torch.cat(parts), full.detach())            # forward is bit-identical per sample
    assert float((lay.weight.grad - gW).abs().max() / gW.abs().max()) < 1e-5
    assert float((lay.bias.grad - gb).abs().max() / gb.abs().max()) < 1e-5


@pytest.mark.parametrize("variant", [("SPMM_PIPE", 4), ("SPMM_PIPE", 1), ("SPMM_CSM", 4), ("SPMM_CSM", 6), ("SPMM_CSM", 8)])
@pytest.mark.parametrize("has_prev", [False, True])
def test_spmm_variants_are_bit_identical(variant, has_prev):
    """The persistent pipelined and the staged-CSR SpMM kernels keep the per-row summation order of the plain
    kernel: bit-identical outputs on a ragged graph (empty rows, rows longer than one batch)."""
    from tgcn_b200 import _lib
    from tgcn_b200.csr import build_csr
    import scipy.sparse as sp
    lib = _lib.load()
    rng = np.random.default_rng(3)
    n, C = 1237, 72
    A = sp.random(n, n, density=0.006, random_state=5, format="lil", dtype=np.float32)
    A[7, :] = 0
    A[11, ::3] = 1.5            # a long row
    plan = build_csr(A.tocsr(), torch.device("cuda"))
    x = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    prev = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda") if has_prev else None
    st = torch.cuda.current_stream().cuda_stream

    def run():
        out = torch.full((n, C), float("nan"), device="cuda")
        rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), n, x.data_ptr(),
                                None if prev is None else prev.data_ptr(), out.data_ptr(), C, 2.0, -1.0, st)
        assert rc == 0, _lib.last_error()
        return out
    try:
        assert lib.tgcn_set_tuning(b"SPMM_PIPE", 0) == 0 and lib.tgcn_set_tuning(b"SPMM_CSM", 0) == 0
        base = run()
        assert lib.tgcn_set_tuning(variant[0].encode(), variant[1]) == 0
        got = run()
    finally:
        for key in (b"SPMM_PIPE", b"SPMM_CSM"):
            lib.tgcn_set_tuning(key, -1)
    assert torch.equal(base, got)
    assert lib.tgcn_set_tuning(b"NOPE", 1) == -1


@pytest.mark.parametrize("u", [4, 5, 8])
@pytest.mark.parametrize("C", [8, 72, 192])
def test_spmm_staged_csr_dense_rows_fall_back_to_global_entries(u, C):
    """Staged-CSR kernel on a graph whose row blocks hold more entries than the staging buffer (1536): those blocks
    read (col, val) from global memory; rows of 0, 1, 2, 3, 5 and ~150 entries; bit-identical to the plain kernel."""
    from tgcn_b200 import _lib
    from tgcn_b200.csr import build_csr
    import scipy.sparse as sp
    lib = _lib.load()
    rng = np.random.default_rng(9)
    n = 301
    A = sp.random(n, n, density=0.5, random_state=2, format="lil", dtype=np.float32)
    for r, k in ((0, 0), (1, 1), (2, 2), (3, 3), (4, 5), (300, 0)):
        A[r, :] = 0
        A[r, :k] = 0.5
    for r in range(200, 264):                # a stretch of short rows: this block IS staged
        A[r, :] = 0
        A[r, r % 7:r % 7 + 6] = 0.25
    plan = build_csr(A.tocsr(), torch.device("cuda"))
    x = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    prev = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        out = torch.full((n, C), float("nan"), device="cuda")
        rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), n, x.data_ptr(),
                                prev.data_ptr(), out.data_ptr(), C, 2.0, -1.0, st)
        assert rc == 0, _lib.last_error()
        return out
    try:
        for key in (b"SPMM_PIPE", b"SPMM_CSM"):
            assert lib.tgcn_set_tuning(key, 0) == 0
        base = run()
        assert lib.tgcn_set_tuning(b"SPMM_CSM", u) == 0
        got = run()
    finally:
        for key in (b"SPMM_PIPE", b"SPMM_CSM"):
            lib.tgcn_set_tuning(key, -1)
    assert torch.equal(base, got)


@pytest.mark.parametrize("R", [4, 8])
@pytest.mark.parametrize("C", [8, 72, 192, 240, 1024])      # V = 2, 18, 48, 60 (warps straddle tiles), 256 (one tile per block)
def test_register_tiled_spmm_matches_plain_kernel(R, C):
    """tgcn_rowtile_plan_create + the register-tiled SpMM kernel (every distinct source row of a tile of R rows
    loaded once and applied to all R rows) against the plain kernel and an fp64 product on the same operands, with
    and without `prev`, `prev` aliasing `out`, N not a multiple of R, empty rows, a non-symmetric operand.  The
    summation order differs from the plain kernel's (one ascending chain instead of two), hence a tolerance: 1e-5 of
    the largest output against the plain kernel, 1e-4 (north_star) against fp64."""
    from tgcn_b200 import _lib
    from tgcn_b200.csr import build_csr
    from conftest import csr_from, load_golden
    import scipy.sparse as sp
    lib = _lib.load()
    L = csr_from(load_golden("graph_grid28_k8_seed0.npz"), "L_0").tolil()     # coarsening order: locality + empty rows
    L[5, 17] = 0.3                                                           # break the symmetry: L^T gets its own plan
    L = L.tocsr()[:-3, :-3]                                                  # N % 8 != 0
    n = L.shape[0]
    assert n % 8 != 0
    rng = np.random.default_rng(R + C)
    x = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    prev = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(plan, with_prev, transpose=False, alias=False):
        out = prev.clone() if alias else torch.full((n, C), float("nan"), device="cuda")
        rp, c, v = (plan.rowptr_t, plan.col_t, plan.val_t) if transpose else (plan.rowptr, plan.col, plan.val)
        pv = out if alias else (prev if with_prev else None)
        before = lib.tgcn_launch_count()
        rc = lib.tgcn_spmm_step(rp.data_ptr(), c.data_ptr(), v.data_ptr(), n, x.data_ptr(),
                                None if pv is None else pv.data_ptr(), out.data_ptr(), C, 2.0, -1.0, st)
        assert rc == 0, _lib.last_error()
        assert lib.tgcn_launch_count() == before + 1
        return out
    plain = build_csr(L, torch.device("cuda"))
    tiled = build_csr(L, torch.device("cuda"))
    info = tiled.ensure_rowtile_plans(rows_per_tile=R, min_gain=0.0)
    assert len(info) == 2 and info[0][2]["gain"] > 1.2 and not tiled.symmetric
    Ld = sp.csr_matrix(L, dtype=np.float64)
    x64, p64 = x.double().cpu().numpy(), prev.double().cpu().numpy()
    for transpose in (False, True):
        M = Ld.T if transpose else Ld
        for with_prev, alias in ((False, False), (True, False), (True, True)):
            base = run(plain, with_prev, transpose, alias)
            got = run(tiled, with_prev, transpose, alias)
            ref = 2.0 * (M @ x64) - (p64 if with_prev else 0.0)
            scale = float(np.abs(ref).max())
            assert not torch.equal(got, torch.full_like(got, float("nan")))
            assert float((got - base).abs().max()) <= 1e-5 * scale
            assert float(np.abs(got.double().cpu().numpy() - ref).max()) <= 1e-4 * scale
    # the persistent, plan-prefetching builds keep the summation order of the one-shot row-tile kernel: bit-identical
    try:
        lib.tgcn_set_tuning(b"SPMM_RTILE", 1)
        one_shot = [run(tiled, True, tr) for tr in (False, True)]
        for mode in (2, 3):
            lib.tgcn_set_tuning(b"SPMM_RTILE", mode)
            for tr in (False, True):
                assert torch.equal(run(tiled, True, tr), one_shot[tr])
    finally:
        lib.tgcn_set_tuning(b"SPMM_RTILE", -1)
    # the tuning key switches the register-tiled path off without touching the plan: bit-identical to plain again
    try:
        lib.tgcn_set_tuning(b"SPMM_RTILE", 0)
        assert torch.equal(run(plain, True), run(tiled, True))
    finally:
        lib.tgcn_set_tuning(b"SPMM_RTILE", -1)


@pytest.mark.parametrize("R", [4, 8])
@pytest.mark.parametrize("C", [72, 256])
def test_persistent_rowtile_spmm_many_groups_per_block(R, C):
    """The persistent row-tile kernel with more tile groups than resident blocks (every block walks several groups,
    prefetching the next plan while it gathers): bit-identical to the one-shot row-tile kernel, 1e-4 of fp64."""
    from tgcn_b200 import _lib, synth
    from tgcn_b200.csr import build_csr
    import scipy.sparse as sp
    lib = _lib.load()
    n = 60001
    A = synth.rgg_adjacency(n, seed=3)[0]
    L = sp.csr_matrix(A, dtype=np.float32)
    plan = build_csr(L, torch.device("cuda"))
    info = plan.ensure_rowtile_plans(rows_per_tile=R, min_gain=0.0)
    assert info
    rng = np.random.default_rng(R * C)
    x = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    prev = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(with_prev):
        out = torch.full((n, C), float("nan"), device="cuda")
        rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), n, x.data_ptr(),
                                prev.data_ptr() if with_prev else None, out.data_ptr(), C, 2.0, -1.0, st)
        assert rc == 0, _lib.last_error()
        return out
    try:
        lib.tgcn_set_tuning(b"SPMM_RTILE", 1)
        base = [run(False), run(True)]
        for mode in (2, 3):
            lib.tgcn_set_tuning(b"SPMM_RTILE", mode)
            assert torch.equal(run(False), base[0]) and torch.equal(run(True), base[1])
    finally:
        lib.tgcn_set_tuning(b"SPMM_RTILE", -1)
    ref = 2.0 * (sp.csr_matrix(L, dtype=np.float64) @ x.double().cpu().numpy()) - prev.double().cpu().numpy()
    assert float(np.abs(base[1].double().cpu().numpy() - ref).max()) <= 1e-4 * float(np.abs(ref).max())


def test_register_tiled_spmm_through_the_layer():
    """A streaming-engine layer whose operand has a row-tile plan: forward, dW, db and dx within 1e-4 of the oracle."""
    from oracle import layers_np
    from tgcn_b200.nn.gcn import TGCNCheb_H
    from conftest import csr_from, load_golden
    L = csr_from(load_golden("graph_grid28_k8_seed0.npz"), "L_0")
    n = L.shape[0]
    Ld = np.asarray(L.todense(), dtype=np.float32)
    Q, H, G, K = 3, 4, 8, 5
    torch.manual_seed(0)
    lay = TGCNCheb_H(torch.tensor(Ld), 1, G, K, H, engine="ffma").cuda()
    x = torch.randn(Q, n, H, device="cuda", requires_grad=True)
    info = lay._plan(x.device).ensure_rowtile_plans(rows_per_tile=8, min_gain=0.0)
    assert info and info[0][2]["gain"] > 1.2
    out = lay(x)
    dout = torch.randn_like(out)
    out.backward(dout)
    W, b = lay.weight.detach().cpu().numpy(), lay.bias.detach().cpu().numpy()
    ref = layers_np.layer_forward(Ld, x.detach().cpu().numpy(), W, b, kind="tgcn_h")
    dW, db, dx = layers_np.layer_backward(Ld, x.detach().cpu().numpy(), W, dout.cpu().numpy(), b.shape, kind="tgcn_h")
    rel = lambda a, r: float(np.abs(a - r).max() / np.abs(r).max())
    assert rel(out.detach().cpu().numpy(), ref) < 1e-4
    assert rel(lay.weight.grad.cpu().numpy(), dW) < 1e-4
    assert rel(lay.bias.grad.cpu().numpy(), db) < 1e-4
    assert rel(x.grad.cpu().numpy(), dx) < 1e-4


# SPMM_RTILE values that select the one-shot row-tile kernel: 1 (default occupancy), 4 (8-row tiles built for 4 blocks per SM)
_RT_MODES = [1, 4]


@pytest.mark.parametrize("mode", _RT_MODES)
@pytest.mark.parametrize("R", [4, 8])
def test_register_tiled_spmm_long_plan_runs_walk_global_memory(R, mode):
    """Blocks whose tiles hold more (tile, source) pairs than the shared-memory stage (768) walk the plan in global
    memory; a stretch of short rows IS staged.  Rows of 0, 1, 2, 3, 5 and ~150 entries, both occupancy builds."""
    from tgcn_b200 import _lib
    from tgcn_b200.csr import build_csr
    import scipy.sparse as sp
    lib = _lib.load()
    rng = np.random.default_rng(R)
    n, C = 301, 72
    A = sp.random(n, n, density=0.5, random_state=2, format="lil", dtype=np.float32)
    for r, k in ((0, 0), (1, 1), (2, 2), (3, 3), (4, 5), (300, 0)):
        A[r, :] = 0
        A[r, :k] = 0.5
    for r in range(112, 300):                # V = 18: blocks of 14 tiles = 56 (R=4) / 112 (R=8) rows; these are staged
        A[r, :] = 0
        A[r, r % 7:r % 7 + 6] = 0.25
    A = A.tocsr()
    x = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    prev = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    tiled = build_csr(A, torch.device("cuda"))
    info = tiled.ensure_rowtile_plans(rows_per_tile=R, min_gain=0.0)
    assert info
    out = torch.full((n, C), float("nan"), device="cuda")
    try:
        assert lib.tgcn_set_tuning(b"SPMM_RTILE", mode) == 0
        rc = lib.tgcn_spmm_step(tiled.rowptr.data_ptr(), tiled.col.data_ptr(), tiled.val.data_ptr(), n, x.data_ptr(),
                                prev.data_ptr(), out.data_ptr(), C, 2.0, -1.0, st)
        assert rc == 0, _lib.last_error()
    finally:
        lib.tgcn_set_tuning(b"SPMM_RTILE", -1)
    ref = 2.0 * (sp.csr_matrix(A, dtype=np.float64) @ x.double().cpu().numpy()) - prev.double().cpu().numpy()
    assert float(np.abs(out.double().cpu().numpy() - ref).max()) <= 1e-5 * float(np.abs(ref).max())
