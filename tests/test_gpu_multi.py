"""Multi-GPU checks (need >= 2 visible GPUs: `gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_and_halo_exchange_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_OK world=2" in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_shot_peer_exchange_two_gpus():
    """The reduce-scatter + pushed all-gather variant of the peer exchange (default from 4 ranks and 2 MB up), forced on
    at world 2: same result as NCCL average + torch.optim.SGD, replicas bit-identical, also under the early/late overlap."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29733", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ, TGCN_PEER_TWOSHOT="1", MGPU_ONLY="peer")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_PEER_OK world=2" in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_layers_under_torch_dataparallel():
    """the layers must also work under torch.nn.DataParallel (the reference's own multi-GPU mode,
    pytorch_hcp_tgcn.py:271-272): replicas run in threads, one per device, each with its own CSR operand."""
    import numpy as np
    import torch.nn as nn
    from conftest import rel_err
    from tgcn_b200 import workloads as wl
    graphs, perm, Ls, n_real = wl.hcp_parcellation()
    Lt = wl.as_torch_operands(Ls, device="cuda:0")
    torch.manual_seed(0)
    model = wl.NetTGCN_HCP(Lt, horizon=15, fused_head=False, drop1=0.0, drop2=0.0).to("cuda:0")
    model.eval()
    x = wl.synthetic_signals(8, Ls[0].shape[0], 15, n_real, perm, seed=1).to("cuda:0")
    ref = model(x)
    dp = nn.DataParallel(model, device_ids=[0, 1])
    out = dp(x)
    assert out.device.index == 0
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    out.sum().backward()
    g_dp = [p.grad.clone() for p in model.parameters()]
    model.zero_grad()
    model(x).sum().backward()
    for a, b in zip(g_dp, model.parameters()):
        assert rel_err(a.cpu().numpy(), b.grad.cpu().numpy()) < 1e-4
