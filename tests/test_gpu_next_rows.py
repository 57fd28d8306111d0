"""SURVEY 8f "next" rows on the GPU: the edge-index operators against goldens of the unmodified reference
(tests/golden/make_golden_edge.py), and the folded time-DFT prologue.  Tolerance 1e-4 (conftest.rel_err)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
EDGE_CASES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("edge_") and f.endswith(".npz"))


@pytest.mark.parametrize("case", EDGE_CASES)
@pytest.mark.parametrize("engine", ["auto", "ffma"])
def test_edge_operators_vs_reference_golden(case, engine):
    from tgcn_b200.nn import gcn as G
    r = load_golden(case)
    K, F, Gc, H = int(r["K"]), int(r["F"]), int(r["G"]), int(r["H"])
    bias = "b" in r
    lay = G.ChebConv(F, Gc, K, bias=bias, engine=engine) if str(r["cls"]) == "ChebConv" else \
        G.ChebTimeConv(F, Gc, K, H, bias=bias, engine=engine)
    with torch.no_grad():
        lay.weight.copy_(torch.tensor(r["W"]))
        if bias:
            lay.bias.copy_(torch.tensor(r["b"]))
    lay = lay.cuda()
    ei = torch.tensor(r["edge_index"], device="cuda")
    ew = torch.tensor(r["edge_weight"], device="cuda") if "edge_weight" in r else None
    x = torch.tensor(r["x"], device="cuda", requires_grad=True)
    out = lay(x, ei, ew)
    assert tuple(out.shape) == r["out"].shape
    assert rel_err(out.detach().cpu().numpy(), r["out"]) < TOL
    out.backward(torch.tensor(r["dout"], device="cuda"))
    assert rel_err(lay.weight.grad.cpu().numpy(), r["dW"]) < TOL
    assert rel_err(x.grad.cpu().numpy(), r["dx"]) < TOL
    if bias:
        assert lay.bias.grad.shape == r["db"].shape
        assert rel_err(lay.bias.grad.cpu().numpy(), r["db"]) < TOL
    # second call with the same edge tensors hits the cached CSR operand
    n_cached = len(lay._edge_cache_list)
    lay(x.detach(), ei, ew)
    assert len(lay._edge_cache_list) == n_cached


def test_time_dft_folded_into_the_weights():
    """layer(time_dft=True)(x) == layer(real(fft(x, axis=2))) (pytorch_mnist_tgcn.py:87), values and gradients."""
    from tgcn_b200.nn import gcn as G
    rng = np.random.default_rng(0)
    N, Q, H, Gc, K = 48, 5, 12, 7, 4
    A = (rng.random((N, N)) < 0.1) * rng.random((N, N)); A = np.maximum(A, A.T); np.fill_diagonal(A, 0)
    d = A.sum(0) + 1e-30
    L = torch.tensor((-(A / np.sqrt(d)[:, None]) / np.sqrt(d)[None, :]).astype(np.float32))
    torch.manual_seed(0)
    a = G.TGCNCheb_H(L, 1, Gc, K, H, time_dft=True).cuda()
    b = G.TGCNCheb_H(L, 1, Gc, K, H).cuda()
    b.load_state_dict(a.state_dict())
    x = torch.tensor(rng.standard_normal((Q, N, H)).astype(np.float32), device="cuda")
    xf = torch.fft.fft(x.double(), dim=2).real.float()
    oa, ob = a(x), b(xf)
    assert rel_err(oa.detach().cpu().numpy(), ob.detach().cpu().numpy()) < 1e-5
    g = torch.randn(oa.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    oa.backward(g); ob.backward(g)
    assert rel_err(a.weight.grad.cpu().numpy(), b.weight.grad.cpu().numpy()) < 1e-5


def test_perm_data_time_on_device():
    from tgcn_b200.nn import gcn as G
    x = torch.arange(2 * 5 * 3, dtype=torch.float32, device="cuda").reshape(2, 5, 3)
    out = G.perm_data_time(x, [4, 7, 0, 1, 5, 2, 3, 6])
    assert out.is_cuda and tuple(out.shape) == (2, 8, 3)
    assert torch.equal(out[:, 0], x[:, 4]) and torch.equal(out[:, 2], x[:, 0])
    assert float(out[:, [1, 4, 7]].abs().sum()) == 0.0


def test_row_partitioned_layer_single_rank_matches_module():
    """RowPartitionedLayer (the config-4 driver, straight C-ABI calls) with world 1 == TGCNCheb_H on the same graph."""
    from tgcn_b200 import workloads as wl
    from tgcn_b200.nn import gcn as G
    from tgcn_b200.parallel import RowPartitionedLayer
    L, _ = wl.random_geometric(n=6000, mean_degree=10.0, seed=3)
    n = L.shape[0]
    Q, H, F, Gc, K = 2, 3, 8, 16, 5
    torch.manual_seed(0)
    lay = G.TGCNCheb_H(L, F, Gc, K, H, engine="ffma")
    lay.bias = torch.nn.Parameter(torch.randn(1, n, Gc))
    lay = lay.cuda()
    x = torch.randn(Q, n, H, F, device="cuda")
    out = lay(x)
    dout = torch.randn_like(out)
    out.backward(dout)
    part = RowPartitionedLayer(L, K, H * F, Gc, device="cuda")
    o2 = part.forward(x.reshape(Q, n, H * F), lay.weight.detach().reshape(K, H * F, Gc).contiguous(), lay.bias.detach()[0].contiguous())
    assert rel_err(o2.cpu().numpy(), out.detach().cpu().numpy()) < 1e-5
    dW, db = part.backward(dout)
    assert rel_err(dW.cpu().numpy(), lay.weight.grad.reshape(K, H * F, Gc).cpu().numpy()) < 1e-4
    assert rel_err(db.cpu().numpy(), lay.bias.grad[0].cpu().numpy()) < 1e-5


def test_peer_allreduce_sgd_single_rank_matches_torch_sgd():
    """csrc/peer.cu with world 1 (own region only): same trajectory as torch.optim.SGD(lr, momentum), eager and
    replayed from a CUDA graph."""
    from tgcn_b200.parallel import PeerAllreduceSGD
    torch.manual_seed(0)
    shapes = [(10, 15, 32), (1, 384, 32), (200,), (6, 200), (7,)]
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    opt_a = PeerAllreduceSGD(pa, lr=0.05, momentum=0.5)
    opt_b = torch.optim.SGD(pb, lr=0.05, momentum=0.5)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(4):
        grads = [torch.randn(p.shape, device="cuda", generator=g) for p in pa]
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone(); q.grad = gr.clone()
        if it == 2:
            pa[3].grad = None; pb[3].grad = torch.zeros_like(pb[3])      # a missing gradient counts as zero
        opt_a.step(); opt_b.step()
        for p, q in zip(pa, pb):
            assert rel_err(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 1e-6
    assert int(opt_a.state[0]) == 4
    # CUDA-graph replay: the step number lives in device memory
    static = [torch.randn(p.shape, device="cuda", generator=g) for p in pa]
    for p, q, gr in zip(pa, pb, static):
        p.grad = gr; q.grad = gr.clone()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt_a.step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt_a.step()
    graph.replay(); graph.replay()
    for _ in range(3):                      # one eager warm-up step + two replays (capture itself does not execute)
        opt_b.step()
    torch.cuda.synchronize()
    assert int(opt_a.state[0]) == 7
    for p, q in zip(pa, pb):
        assert rel_err(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 1e-5
