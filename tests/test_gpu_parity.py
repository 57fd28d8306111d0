"""GPU parity: the CUDA path (through the C-ABI) against (a) golden outputs of the unmodified
reference and (b) the numpy oracle on the same seeded inputs.

Tolerance (north_star): 1e-4 relative in fp32 for activations and gradients, measured as
max|a-ref| / max|ref| (conftest.rel_err); bit-exact for pool indices.
"""
import numpy as np
import pytest
import torch

from conftest import LAYER_CASES, csr_from, load_golden, rel_err
from oracle import layers_np

pytestmark = pytest.mark.gpu
TOL = 1e-4
KIND = {"TGCNCheb_H": "tgcn_h", "GCNCheb": "gcn", "TGCNCheb": "tgcn"}


def make_layer(r, L, **kw):
    from tgcn_b200.nn import gcn as G
    cls = getattr(G, str(r["cls"]))
    K, fin, fout, H = int(r["K"]), int(r["in_ch"]), int(r["out_ch"]), int(r["H"])
    bias = "b" in r
    lay = cls(L, fin, fout, K, H, bias=bias, **kw) if str(r["cls"]) == "TGCNCheb_H" else cls(L, fin, fout, K, bias=bias, **kw)
    with torch.no_grad():
        lay.weight.copy_(torch.tensor(r["W"]))
        if bias:
            lay.bias.copy_(torch.tensor(r["b"]))
    return lay.cuda()


@pytest.mark.parametrize("case", LAYER_CASES)
@pytest.mark.parametrize("ltype", ["dense", "sparse_coo", "scipy"])
def test_layer_forward_backward_vs_reference_golden(case, ltype):
    r = load_golden(case)
    Lsp = csr_from(r, "L")
    dense = torch.tensor(np.asarray(Lsp.todense()), dtype=torch.float)
    L = {"dense": dense.cuda(), "sparse_coo": dense.to_sparse().cuda(), "scipy": Lsp}[ltype]
    if ltype == "scipy" and str(r["cls"]) != "GCNCheb":
        # bias shape is taken from L[0].shape[0] exactly like the reference (gcn.py:96): needs tensor-like L
        L = dense
    lay = make_layer(r, L)
    x = torch.tensor(r["x"], device="cuda", requires_grad=True)
    out = lay(x)
    assert out.shape == r["out"].shape and out.is_contiguous()
    assert rel_err(out.detach().cpu().numpy(), r["out"]) < TOL
    out.backward(torch.tensor(r["dout"], device="cuda"))
    assert rel_err(lay.weight.grad.cpu().numpy(), r["dW"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), r["dx"]) < TOL
    if "b" in r:
        assert lay.bias.grad.shape == r["db"].shape
        assert rel_err(lay.bias.grad.cpu().numpy(), r["db"]) < TOL


@pytest.mark.parametrize("case", LAYER_CASES)
def test_layer_vs_numpy_oracle_fresh_inputs(case):
    """Same graphs, fresh seeded inputs/weights, checker = float64 oracle."""
    r = load_golden(case)
    Lsp = csr_from(r, "L")
    kind = KIND[str(r["cls"])]
    rng = np.random.default_rng(7)
    x = rng.standard_normal(r["x"].shape).astype(np.float32)
    W = rng.standard_normal(r["W"].shape).astype(np.float32) * 0.3
    b = rng.standard_normal(r["b"].shape).astype(np.float32) if "b" in r else None
    dout = rng.standard_normal(r["out"].shape).astype(np.float32)
    r2 = dict(r); r2["W"] = W
    if b is not None:
        r2["b"] = b
    for recursion in ("reference", "chebyshev"):
        lay = make_layer(r2, torch.tensor(np.asarray(Lsp.todense()), dtype=torch.float), recursion=recursion)
        xt = torch.tensor(x, device="cuda", requires_grad=True)
        out = lay(xt)
        out.backward(torch.tensor(dout, device="cuda"))
        ref_out = layers_np.layer_forward(Lsp, x, W, b, kind=kind, recursion=recursion)
        dW, db, dx = layers_np.layer_backward(Lsp, x, W, dout, None if b is None else b.shape, kind=kind,
                                              recursion=recursion)
        assert rel_err(out.detach().cpu().numpy(), ref_out) < TOL
        assert rel_err(lay.weight.grad.cpu().numpy(), dW) < TOL
        assert rel_err(xt.grad.cpu().numpy(), dx) < TOL
        if b is not None:
            assert rel_err(lay.bias.grad.cpu().numpy(), db) < TOL


@pytest.mark.parametrize("case", ["layer_tgcnh_c1_q3.npz", "layer_tgcnh_f3_matmul.npz", "layer_gcn_f32_g64.npz",
                                  "layer_tgcn_f4.npz", "layer_tgcnh_k1.npz"])
def test_basis_method_matches_reference_stack(case):
    r = load_golden(case)
    lay = make_layer(r, torch.tensor(np.asarray(csr_from(r, "L").todense()), dtype=torch.float))
    x = torch.tensor(r["x"], device="cuda")
    fn = lay._chebyshev if str(r["cls"]) == "GCNCheb" else lay._time_chebyshev
    Xt = fn(x).cpu().numpy()
    assert rel_err(Xt.reshape(r["basis"].shape), r["basis"]) < TOL


def test_no_grad_for_input_when_not_required():
    r = load_golden("layer_tgcnh_k3.npz")
    lay = make_layer(r, torch.tensor(np.asarray(csr_from(r, "L").todense()), dtype=torch.float))
    x = torch.tensor(r["x"], device="cuda")          # layer-1 case: data does not require grad
    out = lay(x)
    out.backward(torch.tensor(r["dout"], device="cuda"))
    assert rel_err(lay.weight.grad.cpu().numpy(), r["dW"]) < TOL
    with torch.no_grad():
        assert rel_err(lay(x).cpu().numpy(), r["out"]) < TOL


def test_nonsymmetric_operand_uses_true_transpose_in_backward():
    rng = np.random.default_rng(3)
    N, Q, H, G, K = 40, 3, 4, 5, 5
    A = (rng.random((N, N)) < 0.15) * rng.standard_normal((N, N))
    A = (A / (np.abs(A).sum(1, keepdims=True) + 1)).astype(np.float32)     # keep powers bounded
    from tgcn_b200.nn.gcn import TGCNCheb_H
    lay = TGCNCheb_H(torch.tensor(A), 1, G, K, H).cuda()
    x = rng.standard_normal((Q, N, H)).astype(np.float32)
    dout = rng.standard_normal((Q, N, G)).astype(np.float32)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    lay(xt).backward(torch.tensor(dout, device="cuda"))
    W = lay.weight.detach().cpu().numpy(); b = lay.bias.detach().cpu().numpy()
    dW, db, dx = layers_np.layer_backward(A, x, W, dout, b.shape, kind="tgcn_h")
    assert rel_err(xt.grad.cpu().numpy(), dx) < TOL
    assert rel_err(lay.weight.grad.cpu().numpy(), dW) < TOL


# ------------------------------------------------------------------------------------------------
# pooling: bit-exact values and indices
# ------------------------------------------------------------------------------------------------
def test_pool_bit_exact_vs_reference_golden():
    from tgcn_b200.nn.gcn import gcn_pool, gcn_pool_4, gcn_pool_with_indices
    r = load_golden("pool.npz")
    for p, fn in ((2, gcn_pool), (4, gcn_pool_4)):
        x = torch.tensor(r["x"], device="cuda", requires_grad=True)
        y = fn(x)
        assert np.array_equal(y.detach().cpu().numpy(), r["y%d" % p], equal_nan=True)
        y.backward(torch.tensor(r["dy%d" % p], device="cuda"))
        assert np.array_equal(x.grad.cpu().numpy(), r["dx%d" % p])
        _, idx = gcn_pool_with_indices(x.detach(), p)
        assert idx.dtype == torch.int64
        assert np.array_equal(idx.cpu().numpy(), r["idx%d" % p])


@pytest.mark.parametrize("G", [1, 7, 32, 64])
@pytest.mark.parametrize("p", [2, 4])
def test_pool_and_fused_relu_pool_vs_torch(G, p):
    from tgcn_b200.nn.gcn import gcn_pool, gcn_pool_4, relu_pool
    torch.manual_seed(G * 10 + p)
    x = torch.randn(5, 24 * p, G, device="cuda")
    x[x.abs() < 0.3] = 0.0
    fn = gcn_pool if p == 2 else gcn_pool_4
    for fused in (False, True):
        a = x.clone().requires_grad_(True)
        b = x.clone().requires_grad_(True)
        ya = relu_pool(a, p) if fused else fn(a)
        src = torch.relu(b) if fused else b
        yb = torch.max(src.reshape(5, 24, p, G), dim=2)[0]
        assert torch.equal(ya, yb)
        dy = torch.randn_like(ya)
        ya.backward(dy); yb.backward(dy)
        assert torch.equal(a.grad, b.grad)


def test_pool_rejects_ragged():
    from tgcn_b200.nn.gcn import gcn_pool_4
    with pytest.raises(RuntimeError):
        gcn_pool_4(torch.zeros(2, 62, 3, device="cuda"))


# ------------------------------------------------------------------------------------------------
# raw C-ABI: SpMM step and layout kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [1, 3, 4, 60, 240, 250])
def test_spmm_step_cabi(C):
    import scipy.sparse as sp
    from tgcn_b200 import _lib
    from tgcn_b200.csr import build_csr
    lib = _lib.load()
    rng = np.random.default_rng(C)
    N = 203
    A = sp.random(N, N, density=0.05, random_state=5, dtype=np.float32).tocsr()
    A[17, :] = 0; A[100:110, :] = 0            # empty rows (fake vertices)
    A.eliminate_zeros()
    plan = build_csr(A, "cuda")
    xin = rng.standard_normal((N, C)).astype(np.float32)
    prev = rng.standard_normal((N, C)).astype(np.float32)
    tin, tprev = torch.tensor(xin, device="cuda"), torch.tensor(prev, device="cuda")
    out = torch.empty_like(tin)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, tin.data_ptr(),
                            tprev.data_ptr(), out.data_ptr(), C, 2.0, -1.0, st)
    assert rc == 0, _lib.last_error()
    ref = 2.0 * (A.astype(np.float64) @ xin.astype(np.float64)) - prev
    assert rel_err(out.cpu().numpy(), ref) < 1e-5
    rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, tin.data_ptr(),
                            None, out.data_ptr(), C, 1.0, 0.0, st)
    assert rc == 0
    assert rel_err(out.cpu().numpy(), A.astype(np.float64) @ xin.astype(np.float64)) < 1e-5
    # in-place accumulate (prev aliases out), as the adjoint recursion uses it
    acc = tprev.clone()
    rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, tin.data_ptr(),
                            acc.data_ptr(), acc.data_ptr(), C, 1.0, 1.0, st)
    assert rc == 0
    assert rel_err(acc.cpu().numpy(), A.astype(np.float64) @ xin.astype(np.float64) + prev) < 1e-5
    # aliasing in/out is rejected with a message, not silently wrong
    rc = lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, tin.data_ptr(),
                            None, tin.data_ptr(), C, 1.0, 0.0, st)
    assert rc == -1 and "alias" in _lib.last_error()


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 7), (64, 448, 15), (8, 100, 30), (2, 33, 192), (9, 17, 200)])
def test_slab_layout_round_trip(shape):
    from tgcn_b200 import _lib
    lib = _lib.load()
    Q, N, D = shape
    x = torch.randn(Q, N, D, device="cuda")
    slab = torch.empty(N, Q, D, device="cuda")
    back = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.tgcn_to_slab(x.data_ptr(), slab.data_ptr(), Q, N, D, st) == 0
    assert torch.equal(slab, x.permute(1, 0, 2).contiguous())
    assert lib.tgcn_from_slab(slab.data_ptr(), back.data_ptr(), Q, N, D, st) == 0
    assert torch.equal(back, x)


def test_empty_batch_and_error_reporting():
    from tgcn_b200 import _lib
    from tgcn_b200.nn.gcn import GCNCheb
    lib = _lib.load()
    L = torch.eye(8)
    lay = GCNCheb(L, 2, 3, 4).cuda()
    out = lay(torch.zeros(0, 8, 2, device="cuda"))
    assert out.shape == (0, 8, 3)
    with pytest.raises(RuntimeError, match="vertices"):
        lay(torch.zeros(2, 9, 2, device="cuda"))
    with pytest.raises(RuntimeError):
        lay(torch.zeros(2, 8, 5, device="cuda"))
    with pytest.raises(RuntimeError, match="fp32"):
        lay(torch.zeros(2, 8, 2, device="cuda", dtype=torch.float64))
    rc = lib.tgcn_pool_max_fwd(None, None, None, 1, 6, 2, 3, 0, None, None)
    assert rc == -2 and "pool size" in _lib.last_error()
