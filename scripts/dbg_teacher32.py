#!/usr/bin/env python
"""How far does the REFERENCE's own fp32 arithmetic land from its float64 run in ONE teacher-forced SGD step?

Same models, data and seeds as tests/test_gpu_model.py; CPU only (oracle port of pytorch_hcp_tgcn.py:93-169).  Before
every step the fp32 port receives the float64 port's parameters and momentum; after the step its parameters are compared
with the float64 ones (relative to each tensor's scale, like the GPU test).  This is the noise floor any fp32
implementation of the model has: the GPU test's tolerance cannot be tighter than this.
    python scripts/dbg_teacher32.py [hcp|mesh]"""
import copy
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_torch  # noqa: E402
from tgcn_b200 import workloads as wl  # noqa: E402


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def clone_port(port, dtype):
    ops = {}
    for name, m in port.named_modules():
        if hasattr(m, "L") and isinstance(m.L, torch.Tensor):
            ops[name] = m.L
            m.L = None
    c = copy.deepcopy(port).to(dtype)
    for name, m in port.named_modules():
        if name in ops:
            m.L = ops[name]
    for name, m in c.named_modules():
        if name in ops:
            m.L = ops[name].to(dtype)
    return c


def main(which, threads_list=(16, 1)):
    if which == "mesh":
        graphs, perm, Ls, n_real = wl.cortical_mesh(n_real=9000)
        H, Q, steps, seed0, yseed, mseed = 30, 8, 5, 20, 4, 1
        ops = [t.to_sparse_csr() for t in wl.as_torch_operands(Ls)]
    else:
        graphs, perm, Ls, n_real = wl.hcp_parcellation()
        H, Q, steps, seed0, yseed, mseed = 15, 64, 5, 10, 3, 0
        ops = wl.as_torch_operands(Ls, dense=True)
    torch.manual_seed(mseed)
    port = model_torch.PortNetTGCN_HCP(ops, horizon=H, drop1=0.0, drop2=0.0)
    xs = [wl.synthetic_signals(Q, Ls[0].shape[0], H, n_real, perm, seed=seed0 + i) for i in range(steps)]
    gy = torch.Generator().manual_seed(yseed)
    ys = [torch.randint(0, 6, (Q,), generator=gy) for _ in range(steps)]
    for nthreads in threads_list:                      # a different thread count = a different fp32 summation order
        torch.set_num_threads(nthreads)
        p64, p32 = clone_port(port, torch.float64), clone_port(port, torch.float32)
        o64 = torch.optim.SGD(p64.parameters(), lr=0.01, momentum=0.5)
        o32 = torch.optim.SGD(p32.parameters(), lr=0.01, momentum=0.5)
        p64.train(); p32.train()
        worst = {}
        for i in range(steps):
            with torch.no_grad():
                sd64 = p64.state_dict()
                for name, t in p32.state_dict().items():
                    t.copy_(sd64[name].to(t.dtype))
                for (n64, a), (n32, b) in zip(p64.named_parameters(), p32.named_parameters()):
                    buf = o64.state.get(a, {}).get("momentum_buffer")
                    if buf is not None:
                        o32.state[b]["momentum_buffer"] = buf.to(torch.float32).clone()
            for p_, o_, x_ in ((p64, o64, xs[i].double()), (p32, o32, xs[i].float())):
                o_.zero_grad()
                F.nll_loss(p_(x_), ys[i]).backward()
                o_.step()
            for (name, a), (_, b) in zip(p64.named_parameters(), p32.named_parameters()):
                e = rel_err(b.detach().double().numpy(), a.detach().numpy())
                worst[name] = max(worst.get(name, 0.0), e)
        print("%s  threads=%d  fp32 port vs float64 port, one teacher-forced step, worst over %d steps:" % (which, nthreads, steps))
        for k, v in worst.items():
            print("    %-22s %.3e" % (k, v))
        print("    max %.3e" % max(worst.values()), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "mesh")
