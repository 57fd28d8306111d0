"""Pure-read, pure-write and copy bandwidth of this GPU (context for the read-dominated / write-dominated kernels)."""
import torch
n = 1 << 28   # 1 GiB of fp32
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
w = t(lambda: a.fill_(1.0)); r = t(lambda: a.sum()); c = t(lambda: b.copy_(a))
print("write-only %.0f GB/s   read-only (sum) %.0f GB/s   copy (r+w) %.0f GB/s" % (4 * n / w / 1e6, 4 * n / r / 1e6, 8 * n / c / 1e6))
