"""Host-side pieces of the SURVEY 8f "next" rows (no GPU): edge list -> rescaled Laplacian, the folded
time-DFT matrix and the device-capable perm_data_time, each against the reference's own arithmetic."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from tgcn_b200 import coarsening
from tgcn_b200.nn import gcn as G

EDGE_CASES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("edge_") and f.endswith(".npz"))


@pytest.mark.parametrize("case", EDGE_CASES)
def test_edge_laplacian_reproduces_reference_operator(case):
    """Dense CPU evaluation of the layer with the Laplacian built by laplacian_from_edges (textbook recursion,
    gcn.py:400-417) must reproduce the golden output of the unmodified reference operator."""
    r = load_golden(case)
    ei = torch.tensor(r["edge_index"])
    ew = torch.tensor(r["edge_weight"]) if "edge_weight" in r else None
    x = torch.tensor(r["x"], dtype=torch.float64)
    W = torch.tensor(r["W"], dtype=torch.float64)
    N = x.shape[1]
    row, col, lap = G.laplacian_from_edges(ei, ew, N, dtype=torch.float64)
    L = torch.zeros(N, N, dtype=torch.float64).index_put_((row, col), lap, accumulate=True)
    K = W.shape[0]
    if str(r["cls"]) == "ChebConv":
        x4 = (x if x.dim() == 3 else x.unsqueeze(-1)).unsqueeze(2)     # [Q,N,1,F]
        W4 = W.unsqueeze(1)
    else:
        x4 = x if x.dim() == 4 else x.unsqueeze(-1)
        W4 = W
    T = [x4]
    if K > 1:
        T.append(torch.einsum("nm,qmhf->qnhf", L, x4))
    for k in range(2, K):
        T.append(2 * torch.einsum("nm,qmhf->qnhf", L, T[-1]) - T[-2])
    out = sum(torch.einsum("qnhf,hfg->qng", T[k], W4[k]) for k in range(K))
    if "b" in r:
        out = out + torch.tensor(r["b"], dtype=torch.float64)
    assert np.abs(out.numpy() - r["out"]).max() / np.abs(r["out"]).max() < 1e-5


def test_dft_real_matrix_is_the_real_part_of_numpy_fft():
    for H in (1, 5, 12, 15):
        x = np.random.default_rng(H).standard_normal((3, 4, H))
        ref = np.real(np.fft.fft(x, axis=2))                        # pytorch_mnist_tgcn.py:87
        got = x @ G.dft_real_matrix(H).double().numpy()
        assert np.abs(got - ref).max() < 1e-5


def test_device_perm_data_time_matches_host_version():
    r = load_golden("perm_data.npz")
    perm = r["perm"].tolist()
    M = int(max(p for p in perm if p < len(perm)) + 1)
    M = min(M, 40)
    perm_small = [p for p in perm if p < M] + [M + 3, M + 1]          # two fake vertices at the end
    x = np.random.default_rng(0).standard_normal((3, M, 4)).astype(np.float32)
    ref = coarsening.perm_data_time(x, perm_small)
    got = G.perm_data_time(torch.tensor(x), perm_small)
    assert got.dtype == torch.float32 and np.array_equal(got.numpy().astype(np.float64), ref)
    assert G.perm_data_time(torch.tensor(x), None) is not None
    with pytest.raises(AssertionError):
        G.perm_data_time(torch.tensor(x), perm_small[:M - 1])


def test_edge_layers_mirror_reference_surface():
    torch.manual_seed(0)
    c = G.ChebConv(3, 5, 4)
    t = G.ChebTimeConv(2, 5, 4, 6, bias=False)
    assert repr(c) == "ChebConv(3, 5, K=4)" and repr(t) == "ChebTimeConv(2, 5, K=4)"
    assert tuple(c.weight.shape) == (4, 3, 5) and tuple(c.bias.shape) == (5,)
    assert tuple(t.weight.shape) == (4, 6, 2, 5) and t.bias is None
    assert sorted(c.state_dict()) == ["bias", "weight"]
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        c(torch.zeros(1, 4, 3), torch.tensor([[0, 1], [1, 0]]))
