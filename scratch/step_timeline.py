"""In-graph (warm) kernel durations of one hcp360 training step via the torch profiler (CUPTI)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from tgcn_b200 import workloads as wl
graphs, perm, Ls, n_real = wl.hcp_parcellation()
dev = torch.device("cuda")
Lt = wl.as_torch_operands(Ls, device=dev)
torch.manual_seed(0)
model = wl.NetTGCN_HCP(Lt, horizon=15).to(dev)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.5, fused=True)
x = wl.synthetic_signals(64, Ls[0].shape[0], 15, n_real, perm, seed=1).to(dev)
y = torch.randint(0, 6, (64,), device=dev)
loss_dev = torch.zeros((), device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = F.nll_loss(model(x), y); loss.backward(); loss_dev.copy_(loss.detach()); opt.step()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(3): g.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(10):
        flush.fill_(float(i)); g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
    k = e.name[:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = 0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if "Fill" in k and n <= 10: continue
    print("%8.1f us/step  n/step=%.1f  %s" % (t / 10, n / 10, k)); tot += t / 10
print("sum of kernel time per step: %.1f us" % tot)
ts = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs if "Fill" not in e.name)
# span of one step: first to last kernel of replay 5
per = len(ts) // 10
seg = ts[5 * per:6 * per]
print("step span (first kernel start -> last kernel end): %.1f us, %d kernels" % ((seg[-1][1] - seg[0][0]), len(seg)))
