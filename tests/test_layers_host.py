"""Host-side behaviour of the drop-in modules that needs no GPU: constructor surface, parameter
shapes, RNG draw order (bit-exact vs the reference), repr, state_dict, loud failure off-GPU."""
import numpy as np
import pytest
import torch

from conftest import csr_from, load_golden
from tgcn_b200.csr import build_csr
from tgcn_b200.nn.gcn import GCNCheb, TGCNCheb, TGCNCheb_H, gcn_pool, gcn_pool_4


def L62():
    r = load_golden("layer_tgcnh_k1.npz")
    return torch.tensor(np.asarray(csr_from(r, "L").todense()), dtype=torch.float)


def test_init_matches_reference_rng_order_bit_exact():
    g = load_golden("init.npz")
    Lt = L62()
    torch.manual_seed(42)
    a = TGCNCheb_H(Lt, 2, 5, 3, 4)
    b = GCNCheb(Lt, 5, 6, 4)
    c = TGCNCheb(Lt, 3, 2, 5)
    d = TGCNCheb_H(Lt, 1, 3, 2, 4, bias=False)
    e = GCNCheb(Lt, 2, 2, 2)
    for nm, lay in (("a", a), ("b", b), ("c", c), ("d", d), ("e", e)):
        assert np.array_equal(lay.weight.detach().numpy(), g[nm + "_weight"])
        if lay.bias is not None:
            assert np.array_equal(lay.bias.detach().numpy(), g[nm + "_bias"])
        assert repr(lay) == str(g[nm + "_repr"])
    assert d.bias is None
    assert list(a.state_dict().keys()) == ["weight", "bias"]
    assert list(d.state_dict().keys()) == ["weight"]
    assert a.L is Lt and a.filter_order == 3 and a.in_channels == 2 and a.out_channels == 5


def test_shapes():
    Lt = L62()
    assert tuple(TGCNCheb_H(Lt, 2, 5, 3, 4).weight.shape) == (3, 4, 2, 5)
    assert tuple(TGCNCheb_H(Lt, 2, 5, 3, 4).bias.shape) == (1, 62, 5)
    assert tuple(GCNCheb(Lt, 2, 5, 3).bias.shape) == (1, 1, 5)
    assert tuple(TGCNCheb(Lt, 2, 5, 3).bias.shape) == (1, 62, 5)


def test_no_cpu_fallback():
    Lt = L62()
    with pytest.raises(RuntimeError, match="no CPU path"):
        TGCNCheb_H(Lt, 1, 3, 2, 4)(torch.zeros(2, 62, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gcn_pool(torch.zeros(2, 62, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gcn_pool_4(torch.zeros(2, 64, 4))


def test_csr_plan_from_every_accepted_input_type():
    r = load_golden("layer_tgcnh_rand.npz")
    Lsp = csr_from(r, "L")
    dense = torch.tensor(np.asarray(Lsp.todense()), dtype=torch.float)
    coo = dense.to_sparse()
    csr = dense.to_sparse_csr()
    import scipy.sparse as sp
    plans = [build_csr(v, "cpu") for v in (dense, coo, csr, Lsp, np.asarray(Lsp.todense()))]
    ref = Lsp.tocsr(); ref.sort_indices()
    for p in plans:
        assert p.n == ref.shape[0] and p.nnz == ref.nnz
        assert np.array_equal(p.rowptr.numpy(), ref.indptr) and np.array_equal(p.col.numpy(), ref.indices)
        assert np.array_equal(p.val.numpy(), ref.data)
        assert p.rowptr.dtype == torch.int32 and p.col.dtype == torch.int32
        # D^-1/2 W D^-1/2 in fp32 is symmetric only up to rounding: the adjoint must use the true L^T
        T = sp.csr_matrix((p.val_t.numpy(), p.col_t.numpy(), p.rowptr_t.numpy()), shape=ref.shape)
        assert (T != ref.T.tocsr()).nnz == 0
    # non-symmetric operand: transposed triplet really is the transpose
    A = torch.tensor([[0., 1., 0.], [0., 0., 2.], [3., 0., 0.]])
    p = build_csr(A, "cpu")
    assert not p.symmetric
    T = sp.csr_matrix((p.val_t.numpy(), p.col_t.numpy(), p.rowptr_t.numpy()), shape=(3, 3)).toarray()
    assert np.array_equal(T, A.numpy().T)
