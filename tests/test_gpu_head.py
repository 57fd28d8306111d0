"""Fused classifier head (csrc/head.cu) against the plain torch modules it replaces (fp32 reference of the same
ops: Linear -> BatchNorm1d -> ReLU -> Linear -> log_softmax), forward, every gradient and the running statistics."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _mods(I, Hd, C, seed):
    torch.manual_seed(seed)
    fc1, bn, fc2 = nn.Linear(I, Hd), nn.BatchNorm1d(Hd), nn.Linear(Hd, C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    return fc1.cuda(), bn.cuda(), fc2.cuda()


@pytest.mark.parametrize("shape", [(64, 1536, 200, 6), (7, 50, 13, 3), (100, 300, 30, 10), (5, 8, 4, 32)])
def test_fused_head_matches_torch(shape):
    from tgcn_b200.nn.head import fused_head
    Q, I, Hd, C = shape
    a = _mods(I, Hd, C, 0)
    b = _mods(I, Hd, C, 0)
    for ma, mb in zip(a, b):
        mb.load_state_dict(ma.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(Q, I, device="cuda", generator=g)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y = torch.randint(0, C, (Q,), device="cuda", generator=g)
    for m in a + b:
        m.train()
    la = fused_head(xa, *a)
    lb = F.log_softmax(b[2](F.relu(b[1](b[0](xb)))), dim=1)
    assert rel_err(la.detach().cpu().numpy(), lb.detach().cpu().numpy()) < 1e-5
    F.nll_loss(la, y).backward()
    F.nll_loss(lb, y).backward()
    assert rel_err(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) < 1e-4
    for ma, mb in zip(a, b):
        for (na, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            if ma is a[0] and na == "bias":
                # BatchNorm removes the batch mean, so d loss / d fc1.bias is exactly zero: both sides are rounding noise
                assert float(pa.grad.abs().max()) < 1e-6 and float(pb.grad.abs().max()) < 1e-6
                continue
            assert rel_err(pa.grad.cpu().numpy(), pb.grad.cpu().numpy()) < 1e-4, na
    assert rel_err(a[1].running_mean.cpu().numpy(), b[1].running_mean.cpu().numpy()) < 1e-5
    assert rel_err(a[1].running_var.cpu().numpy(), b[1].running_var.cpu().numpy()) < 1e-5
    assert int(a[1].num_batches_tracked) == int(b[1].num_batches_tracked) == 1
    # evaluation mode uses the running statistics
    for m in a + b:
        m.eval()
    with torch.no_grad():
        ea = fused_head(x, *a)
        eb = F.log_softmax(b[2](F.relu(b[1](b[0](x)))), dim=1)
    assert rel_err(ea.cpu().numpy(), eb.cpu().numpy()) < 1e-5


def test_model_with_fused_head_matches_unfused():
    from tgcn_b200 import workloads as wl
    graphs, perm, Ls, n_real = wl.hcp_parcellation()
    Lt = wl.as_torch_operands(Ls, device="cuda")
    torch.manual_seed(0)
    ma = wl.NetTGCN_HCP(Lt, horizon=15, fused_head=True, drop1=0.0, drop2=0.0).cuda()
    mb = wl.NetTGCN_HCP(Lt, horizon=15, fused_head=False, drop1=0.0, drop2=0.0).cuda()
    mb.load_state_dict(ma.state_dict())
    x = wl.synthetic_signals(16, Ls[0].shape[0], 15, n_real, perm, seed=2).cuda()
    y = torch.randint(0, 6, (16,), device="cuda")
    ma.train(); mb.train()
    F.nll_loss(ma(x), y).backward()
    F.nll_loss(mb(x), y).backward()
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        if n == "fc1.bias":
            assert float(pa.grad.abs().max()) < 1e-6       # exactly zero in exact arithmetic (BatchNorm follows)
            continue
        assert rel_err(pa.grad.cpu().numpy(), pb.grad.cpu().numpy()) < 1e-4, n
