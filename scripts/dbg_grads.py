"""One training step of the mesh model: per-parameter gradient error of the GPU path and of the fp32 CPU port against
the float64 port."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from oracle import model_torch
from tgcn_b200 import workloads as wl
n_real = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
graphs, perm, Ls, n_real = wl.cortical_mesh(n_real=n_real)
H, Q = 30, 8
torch.manual_seed(1)
model = wl.NetTGCN_HCP(wl.as_torch_operands(Ls, device="cuda"), horizon=H, drop1=0.0, drop2=0.0).cuda()
Lcpu = [t.to_sparse_csr() for t in wl.as_torch_operands(Ls)]
port = model_torch.PortNetTGCN_HCP(Lcpu, horizon=H, drop1=0.0, drop2=0.0)
port.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
port64 = model_torch.PortNetTGCN_HCP([t.to(torch.float64) for t in Lcpu], horizon=H, drop1=0.0, drop2=0.0).double()
port64.load_state_dict({k: v.cpu().double() for k, v in model.state_dict().items()})
x = wl.synthetic_signals(Q, Ls[0].shape[0], H, n_real, perm, seed=20)
y = torch.randint(0, 6, (Q,), generator=torch.Generator().manual_seed(4))
for m in (model, port, port64):
    m.train()
lg = F.nll_loss(model(x.cuda()), y.cuda()); lg.backward()
l32 = F.nll_loss(port(x), y); l32.backward()
l64 = F.nll_loss(port64(x.double()), y); l64.backward()
print("loss gpu %.9f cpu32 %.9f f64 %.9f" % (float(lg), float(l32), float(l64)))
def rel(a, b):
    s = float(np.abs(b).max()) or 1.0
    return float(np.abs(a - b).max()) / s
for (n, p), q, r in zip(model.named_parameters(), port.parameters(), port64.parameters()):
    ref = r.grad.numpy()
    print("%-18s |g|max %.3e   gpu err %.2e   cpu32 err %.2e" % (n, np.abs(ref).max(), rel(p.grad.cpu().numpy().astype(np.float64), ref), rel(q.grad.numpy().astype(np.float64), ref)))
