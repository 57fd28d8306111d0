"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there).

* cyclic-shift operand: L = row-shift by one => L^j x is x rolled by j: closed form for the whole layer
* linearity in x and in W
* K=1 reduces to a per-vertex dense contraction
* data-parallel invariance: batch halves give the same outputs / summed weight gradients
"""
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
