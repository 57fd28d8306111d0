import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from tgcn_b200.parallel import init_distributed
rank, world, local = init_distributed("nccl")
torch.cuda.set_device(local)
def log(*a):
    print("[r%d %.1f]" % (rank, time.time() % 1000), *a, flush=True)
    open("/root/repo/gpurun_out/probe_r%d.log" % rank, "a").write(" ".join(str(v) for v in a) + "\n")
x = torch.ones(1 << 18, device="cuda") * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize(); log("eager allreduce ok", float(x[0]))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        x.mul_(0.5); dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); log("side-stream warmup ok")
gs = []
for b in range(2):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        x.mul_(0.5); dist.all_reduce(x, op=dist.ReduceOp.AVG); x.add_(1.0)
    gs.append(g); log("captured", b)
torch.cuda.synchronize(); dist.barrier(); log("barrier ok")
for i in range(20):
    gs[i % 2].replay()
torch.cuda.synchronize(); log("replays ok", float(x[0]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(50): gs[i % 2].replay()
b.record(); torch.cuda.synchronize(); log("50 replays: %.1f us each" % (a.elapsed_time(b) * 20))
gl = dist.new_group(backend="gloo")
log("gloo group made")
dist.barrier(group=gl); log("gloo barrier ok")
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize(); log("eager nccl allreduce after replays ok", float(t[0]))
dist.barrier(); log("nccl barrier ok")
import threading
threading.Timer(15.0, lambda: (log("destroy hung -> _exit"), os._exit(0))).start()
dist.destroy_process_group(); log("done")
os._exit(0)
