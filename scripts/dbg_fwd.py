"""Teacher-forced forward comparison (mesh model): activations of the GPU model vs the float64 port at each step."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from oracle import model_torch, layers_torch as ref
from tgcn_b200 import workloads as wl
graphs, perm, Ls, n_real = wl.cortical_mesh(n_real=9000)
H, Q, steps = 30, 8, 5
Lt = wl.as_torch_operands(Ls, device="cuda")
Lcpu = [t.to_sparse_csr() for t in wl.as_torch_operands(Ls)]
xs = [wl.synthetic_signals(Q, Ls[0].shape[0], H, n_real, perm, seed=20 + i) for i in range(steps)]
gy = torch.Generator().manual_seed(4)
ys = [torch.randint(0, 6, (Q,), generator=gy) for _ in range(steps)]
torch.manual_seed(1)
m = wl.NetTGCN_HCP(Lt, horizon=H, drop1=0.0, drop2=0.0).cuda(); m.train()
p64 = model_torch.PortNetTGCN_HCP([t.to(torch.float64) for t in Lcpu], horizon=H, drop1=0.0, drop2=0.0).double(); p64.train()
p64.load_state_dict({k: v.cpu().double() for k, v in m.state_dict().items()})
o64 = torch.optim.SGD(p64.parameters(), lr=0.01, momentum=0.5)
def rel(a, b):
    a = a.detach().cpu().double().numpy(); b = b.detach().double().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for i in range(steps):
    with torch.no_grad():
        for n, t in m.state_dict().items():
            t.copy_(p64.state_dict()[n].to(t.dtype))
    x = xs[i]
    with torch.no_grad():
        # GPU
        a1 = m.tgcn1.forward_relu_pool(x.cuda(), 4)
        a2 = m.gcn2.forward_relu_pool(a1, 4)
        h = F.linear(a2.reshape(Q, -1), m.fc1.weight, m.fc1.bias)
        # port64
        xf = torch.fft.fft(x.double(), dim=2).real
        b1 = ref.pool(F.relu(p64.tgcn1(xf)), 4)
        b2 = ref.pool(F.relu(p64.gcn2(b1)), 4)
        h64 = p64.fc1(b2.reshape(Q, -1))
        # head statistics
        sd = h64.std(dim=0, unbiased=False); mu = h64.mean(dim=0).abs()
    lg = F.nll_loss(m(x.cuda()), ys[i].cuda())
    o64.zero_grad(); l64 = F.nll_loss(p64(x.double()), ys[i]); l64.backward(); o64.step()
    print("step %d  a1 %.2e  a2 %.2e  fc1 %.2e  | loss gpu %.8f f64 %.8f  | |W1| %.3f  min std/|mean| %.2e  max|h| %.2f"
          % (i, rel(a1, b1), rel(a2, b2), rel(h, h64), float(lg), float(l64), float(m.fc1.weight.abs().max()), float((sd / (mu + 1e-12)).min()), float(h64.abs().max())), flush=True)
