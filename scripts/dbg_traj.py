"""Loss trajectory of the mesh model under different optimizer paths vs the float64 port."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from oracle import model_torch
from tgcn_b200 import workloads as wl
from tgcn_b200.nn.head import Fc1FusedSGD
from tgcn_b200.parallel import PeerAllreduceSGD
graphs, perm, Ls, n_real = wl.cortical_mesh(n_real=9000)
H, Q, steps = 30, 8, 5
Lt = wl.as_torch_operands(Ls, device="cuda")
Lcpu = [t.to_sparse_csr() for t in wl.as_torch_operands(Ls)]
xs = [wl.synthetic_signals(Q, Ls[0].shape[0], H, n_real, perm, seed=20 + i) for i in range(steps)]
gy = torch.Generator().manual_seed(4)
ys = [torch.randint(0, 6, (Q,), generator=gy) for _ in range(steps)]
def fresh():
    torch.manual_seed(1)
    return wl.NetTGCN_HCP(Lt, horizon=H, drop1=0.0, drop2=0.0).cuda()
base = fresh()
port64 = model_torch.PortNetTGCN_HCP([t.to(torch.float64) for t in Lcpu], horizon=H, drop1=0.0, drop2=0.0).double()
port64.load_state_dict({k: v.cpu().double() for k, v in base.state_dict().items()})
o64 = torch.optim.SGD(port64.parameters(), lr=0.01, momentum=0.5); port64.train()
ref = []
for i in range(steps):
    o64.zero_grad(); l = F.nll_loss(port64(xs[i].double()), ys[i]); l.backward(); o64.step(); ref.append(float(l.detach()))
print("f64        ", ["%.7f" % v for v in ref])
def run(tag, fused, optk, fused_head=True, engine="auto"):
    torch.manual_seed(1)
    m = wl.NetTGCN_HCP(Lt, horizon=H, drop1=0.0, drop2=0.0, fused_head=fused_head, engine=engine).cuda(); m.train()
    params = list(m.parameters())
    if fused:
        m.fc1_update = Fc1FusedSGD(m.fc1.weight, lr=0.01, momentum=0.5)
        params = [p for p in params if p is not m.fc1.weight]
    opt = PeerAllreduceSGD(params, lr=0.01, momentum=0.5) if optk == "peer" else torch.optim.SGD(params, lr=0.01, momentum=0.5)
    out = []
    for i in range(steps):
        opt.zero_grad(set_to_none=True)
        l = F.nll_loss(m(xs[i].cuda()), ys[i].cuda()); l.backward(); opt.step(); out.append(float(l.detach()))
    print("%-11s" % tag, ["%.7f" % v for v in out], "max dev %.2e" % max(abs(a - b) for a, b in zip(out, ref)))
run("torch-sgd", False, "torch")
run("peer-sgd", False, "peer")
run("fused-fc1", True, "peer")
run("plain-head", False, "torch", fused_head=False)
run("ffma", False, "torch", engine="ffma")
