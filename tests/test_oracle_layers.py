"""Pin the numpy oracle (oracle/layers_np.py) and the torch-CPU port (oracle/layers_torch.py)
against golden outputs of the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import LAYER_CASES, csr_from, load_golden, rel_err
from oracle import layers_np, layers_torch

ORACLE_TOL = 2e-5     # float64 oracle vs the reference's own fp32 arithmetic
KIND = {"TGCNCheb_H": "tgcn_h", "GCNCheb": "gcn", "TGCNCheb": "tgcn"}


@pytest.mark.parametrize("case", LAYER_CASES)
def test_numpy_oracle_matches_reference(case):
    r = load_golden(case)
    L = csr_from(r, "L")
    kind = KIND[str(r["cls"])]
    K = int(r["K"])
    basis = layers_np.cheb_basis(L, layers_np._canon_x(r["x"], kind), K)
    assert rel_err(basis.reshape(r["basis"].shape), r["basis"]) < ORACLE_TOL
    b = r.get("b")
    out = layers_np.layer_forward(L, r["x"], r["W"], b, kind=kind)
    assert out.shape == r["out"].shape
    assert rel_err(out, r["out"]) < ORACLE_TOL
    dW, db, dx = layers_np.layer_backward(L, r["x"], r["W"], r["dout"], None if b is None else b.shape, kind=kind)
    assert rel_err(dW, r["dW"]) < ORACLE_TOL
    assert rel_err(dx, r["dx"]) < ORACLE_TOL
    if b is not None:
        assert db.shape == r["db"].shape
        assert rel_err(db, r["db"]) < ORACLE_TOL


@pytest.mark.parametrize("case", LAYER_CASES)
def test_reference_recursion_is_not_textbook_chebyshev(case):
    """Documents the semantic finding: for K >= 4 the reference stack differs from T_k."""
    r = load_golden(case)
    K = int(r["K"])
    if K < 4:
        pytest.skip("recursions coincide for K < 4")
    L = csr_from(r, "L")
    kind = KIND[str(r["cls"])]
    x = layers_np._canon_x(r["x"], kind)
    true_cheb = layers_np.cheb_basis(L, x, K, recursion="chebyshev")
    assert rel_err(true_cheb[:3].reshape(r["basis"][:3].shape), r["basis"][:3]) < ORACLE_TOL
    assert rel_err(true_cheb[3].reshape(r["basis"][3].shape), r["basis"][3]) > 1e-2


@pytest.mark.parametrize("case", LAYER_CASES)
def test_mix_matrix_reproduces_stack(case):
    r = load_golden(case)
    K = int(r["K"])
    L = csr_from(r, "L")
    kind = KIND[str(r["cls"])]
    x = layers_np._canon_x(r["x"], kind)
    P = [x]
    for _ in range(1, K):
        P.append(layers_np._apply_vertex_op(layers_np._as_op(L)[0], P[-1]))
    M = layers_np.mix_matrix(K)
    stack = np.einsum("kj,j...->k...", M, np.stack(P))
    assert rel_err(stack.reshape(r["basis"].shape), r["basis"]) < ORACLE_TOL


@pytest.mark.parametrize("case", LAYER_CASES)
def test_torch_port_matches_reference(case):
    r = load_golden(case)
    L = torch.tensor(np.asarray(csr_from(r, "L").todense()), dtype=torch.float32)
    x = torch.tensor(r["x"], requires_grad=True)
    W = torch.tensor(r["W"], requires_grad=True)
    b = torch.tensor(r["b"], requires_grad=True) if "b" in r else None
    out = layers_torch.cheb_layer(L, x, W, b)
    out.backward(torch.tensor(r["dout"]))
    # same ATen calls as the reference on the same machine type: expect (near) bit equality
    assert rel_err(out.detach().numpy(), r["out"]) < 1e-6
    assert rel_err(W.grad.numpy(), r["dW"]) < 1e-6
    assert rel_err(x.grad.numpy(), r["dx"]) < 1e-6
    if b is not None:
        assert rel_err(b.grad.numpy(), r["db"]) < 1e-6


def test_torch_port_sparse_slab_path():
    r = load_golden("layer_tgcnh_rand.npz")
    Lsp = csr_from(r, "L")
    L = torch.sparse_csr_tensor(torch.tensor(Lsp.indptr), torch.tensor(Lsp.indices), torch.tensor(Lsp.data),
                                size=Lsp.shape)
    out = layers_torch.cheb_layer(L, torch.tensor(r["x"]), torch.tensor(r["W"]), torch.tensor(r["b"]))
    assert rel_err(out.numpy(), r["out"]) < 1e-5


def test_pool_oracle_matches_reference_values_and_indices():
    r = load_golden("pool.npz")
    for p in (2, 4):
        y, idx = layers_np.pool_forward(r["x"], p)
        assert np.array_equal(idx, r["idx%d" % p])                       # bit-exact indices
        assert np.array_equal(y, r["y%d" % p], equal_nan=True)           # bit-exact values
        dx = layers_np.pool_backward(r["dy%d" % p], idx, p)
        assert np.array_equal(dx, r["dx%d" % p])


def test_pool_rejects_ragged_vertex_count():
    with pytest.raises(ValueError):
        layers_np.pool_forward(np.zeros((1, 6, 2), np.float32), 4)
