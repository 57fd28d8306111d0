"""Host-side behaviour of the drop-in layers that needs no GPU: operand size detection, the `L` attribute,
copy / pickle of modules, the COO -> CSR builder of the edge-index operators (pure torch, runs on the CPU device)."""
import copy
import io
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tgcn_b200.csr import build_csr_from_coo
from tgcn_b200.nn import gcn as G


def _ring(n):
    A = sp.diags([np.ones(n - 1), np.ones(n - 1)], [1, -1]).tolil()
    A[0, n - 1] = A[n - 1, 0] = 1.0
    d = np.asarray(A.sum(1)).ravel()
    return sp.csr_matrix(-(sp.diags(d ** -0.5) @ A.tocsr() @ sp.diags(d ** -0.5))).astype(np.float32)


@pytest.mark.parametrize("form", ["scipy", "dense", "torch_sparse", "numpy"])
def test_bias_is_sized_from_the_operand_shape(form):
    """Reference: bias [1, L[0].shape[0], G] (gcn.py:22,96).  For a scipy CSR `L[0]` is a 1 x N row -- the bias must
    still have N rows (an [1,1,G] bias with the per-vertex kernels would be an out-of-bounds access)."""
    n = 12
    Ls = _ring(n)
    L = {"scipy": Ls, "dense": torch.tensor(Ls.toarray()), "numpy": Ls.toarray(),
         "torch_sparse": torch.tensor(Ls.toarray()).to_sparse()}[form]
    assert tuple(G.TGCNCheb_H(L, 2, 5, 3, 4).bias.shape) == (1, n, 5)
    assert tuple(G.TGCNCheb(L, 2, 5, 3).bias.shape) == (1, n, 5)
    assert tuple(G.GCNCheb(L, 2, 5, 3).bias.shape) == (1, 1, 5)


def test_assigning_L_drops_the_cached_operand():
    lay = G.GCNCheb(_ring(8), 1, 2, 3)
    lay._csr._plans[("cpu", None)] = "stale"
    newL = _ring(8) * 0.5
    lay.L = newL
    assert lay.L is newL
    assert lay._csr._plans == {}
    assert "L" not in lay.state_dict() and sorted(lay.state_dict()) == ["bias", "weight"]


def test_modules_can_be_deep_copied_pickled_and_saved():
    """The reference modules are plain nn.Modules: copy.deepcopy / pickle / torch.save(model) all work on them."""
    lay = G.TGCNCheb_H(torch.tensor(_ring(8).toarray()), 1, 2, 3, 4)
    lay._csr._plans[("cpu", None)] = object()           # cached device plans never travel
    for clone in (copy.deepcopy(lay), pickle.loads(pickle.dumps(lay))):
        assert torch.equal(clone.weight, lay.weight) and torch.equal(clone.L, lay.L)
        assert clone._csr._plans == {}
    buf = io.BytesIO()
    torch.save(lay, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert torch.equal(back.bias, lay.bias)


def test_coo_to_csr_builder_matches_scipy():
    rng = np.random.default_rng(0)
    n, e = 50, 400
    row, col = rng.integers(0, n, e), rng.integers(0, n, e)
    val = rng.standard_normal(e).astype(np.float32)
    plan = build_csr_from_coo(torch.tensor(row), torch.tensor(col), torch.tensor(val), n, "cpu")
    ref = sp.coo_matrix((val.astype(np.float64), (row, col)), shape=(n, n)).tocsr()
    ref.sort_indices()
    assert np.array_equal(plan.rowptr.numpy(), ref.indptr)
    assert np.array_equal(plan.col.numpy(), ref.indices)
    np.testing.assert_allclose(plan.val.numpy(), ref.data, rtol=1e-6, atol=1e-6)
    refT = ref.T.tocsr()
    refT.sort_indices()
    assert np.array_equal(plan.rowptr_t.numpy(), refT.indptr) and np.array_equal(plan.col_t.numpy(), refT.indices)
    np.testing.assert_allclose(plan.val_t.numpy(), refT.data, rtol=1e-6, atol=1e-6)
    assert not plan.symmetric
    # symmetric input is recognised (one CSR shared by the forward and the adjoint recursion)
    up = sp.triu(ref, k=1).tocoo()                                          # duplicate-free upper triangle
    r2, c2 = np.concatenate([up.row, up.col]), np.concatenate([up.col, up.row])
    v2 = np.concatenate([up.data, up.data]).astype(np.float32)
    plan2 = build_csr_from_coo(torch.tensor(r2), torch.tensor(c2), torch.tensor(v2), n, "cpu")
    assert plan2.symmetric and plan2.rowptr_t is plan2.rowptr


def test_edge_operator_refuses_to_drop_the_edge_weight_gradient():
    lay = G.ChebConv(1, 2, 3)
    ei = torch.tensor([[0, 1], [1, 0]])
    ew = torch.ones(2, requires_grad=True)
    with pytest.raises(NotImplementedError, match="edge_weight"):
        lay._edge_plan(ei, ew, 2, torch.device("cpu"))
    a = lay._edge_plan(ei, None, 2, torch.device("cpu"))
    assert lay._edge_plan(ei, None, 2, torch.device("cpu")) is a            # same tensor object, same version: hit
    assert lay._edge_plan(ei.clone(), None, 2, torch.device("cpu")) is not a   # equal contents, other object: rebuilt
    ei[0, 0] = 0                                                             # in-place edit bumps the version
    assert lay._edge_plan(ei, None, 2, torch.device("cpu")) is not a


def test_port_model_matches_the_product_model_parameter_names():
    """state_dict keys of the CPU port and of the product model agree, so a state_dict copies across."""
    from oracle import model_torch
    from tgcn_b200 import workloads as wl
    graphs, perm, Ls, n_real = wl.hcp_parcellation(n_real=40, knn=6)
    Lt = wl.as_torch_operands(Ls, dense=True)
    a = wl.NetTGCN_HCP(Lt, horizon=5, K=3, g1=4, g2=4, hidden=8)
    b = model_torch.PortNetTGCN_HCP(Lt, horizon=5, K=3, g1=4, g2=4, hidden=8)
    assert sorted(a.state_dict()) == sorted(b.state_dict())
    b.load_state_dict(a.state_dict())
    # the port runs on the CPU (dropout off in eval mode): finite log-probabilities
    b.eval()
    out = b(torch.randn(3, Ls[0].shape[0], 5))
    assert out.shape == (3, 6) and torch.isfinite(out).all()
