"""Worker for the multi-GPU checks; launched by tests/test_gpu_multi.py under torchrun (NCCL).

A. data parallel: sharded batch + one flat allreduce == single-GPU full-batch gradients (rtol 1e-4)
B. row-partitioned Chebyshev recursion with NVLink halo exchange == unpartitioned, bit for bit
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tgcn_b200 import _lib, workloads as wl  # noqa: E402
from tgcn_b200.csr import build_csr  # noqa: E402
from tgcn_b200.parallel import (FlatGradients, PeerAllreduceSGD, RowPartition, RowPartitionedLayer,  # noqa: E402
                                broadcast_parameters, halo_exchange, init_distributed, shard_range)


def check_dp(rank, world, dev):
    graphs, perm, Ls, n_real = wl.hcp_parcellation()
    Lt = wl.as_torch_operands(Ls, device=dev)
    torch.manual_seed(10 + rank)
    model = wl.NetTGCN_HCP(Lt, horizon=15, drop1=0.0, drop2=0.0).to(dev)
    broadcast_parameters(model)
    model.eval()
    Q = 16 * world
    x = wl.synthetic_signals(Q, Ls[0].shape[0], 15, n_real, perm, seed=3).to(dev)
    y = torch.randint(0, 6, (Q,), generator=torch.Generator().manual_seed(4)).to(dev)
    ref = wl.NetTGCN_HCP(Lt, horizon=15, drop1=0.0, drop2=0.0).to(dev)
    ref.load_state_dict(model.state_dict()); ref.eval()
    F.nll_loss(ref(x), y).backward()
    ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    grads = FlatGradients(model.parameters())
    lo, hi = shard_range(Q, rank, world)
    grads.zero_()
    F.nll_loss(model(x[lo:hi]), y[lo:hi]).backward()
    flat = grads.allreduce_mean()
    err = float((flat - ref_flat).abs().max() / ref_flat.abs().max())
    assert err < 1e-4, err
    return err


def check_halo(rank, world, dev):
    lib = _lib.load()
    L, _ = wl.random_geometric(n=40000, mean_degree=12.0, seed=1)
    n = L.shape[0]
    C, K = 48, 6
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.standard_normal((n, C)).astype(np.float32), device=dev)
    # unpartitioned reference on this GPU
    full = build_csr(L, dev)
    st = torch.cuda.current_stream().cuda_stream
    ref = [x]
    for k in range(1, K):
        out = torch.empty_like(x)
        rc = lib.tgcn_spmm_step(full.rowptr.data_ptr(), full.col.data_ptr(), full.val.data_ptr(), n, ref[-1].data_ptr(),
                                None, out.data_ptr(), C, 1.0, 0.0, st)
        assert rc == 0, _lib.last_error()
        ref.append(out)
    plan = RowPartition(L, rank, world).build_send_lists()
    rowptr = torch.tensor(plan.rowptr, device=dev); col = torch.tensor(plan.col, device=dev)
    val = torch.tensor(plan.val, device=dev)
    ext = torch.empty(plan.n_own + plan.n_halo, C, device=dev)
    ext[:plan.n_own] = x[plan.lo:plan.hi]
    for k in range(1, K):
        halo_exchange(ext[:plan.n_own], plan, halo_out=ext[plan.n_own:])
        nxt = torch.empty_like(ext)
        rc = lib.tgcn_spmm_step(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), plan.n_own, ext.data_ptr(), None,
                                nxt.data_ptr(), C, 1.0, 0.0, st)
        assert rc == 0, _lib.last_error()
        assert torch.equal(nxt[:plan.n_own], ref[k][plan.lo:plan.hi]), "step %d differs" % k
        ext = nxt
    return plan.n_halo


def check_partitioned_layer(rank, world, dev):
    """C. row-partitioned layer (forward, dW, db) == the same layer on one GPU holding the whole graph."""
    L, _ = wl.random_geometric(n=30000, mean_degree=10.0, seed=2)
    n = L.shape[0]
    Q, D, G, K = 2, 24, 16, 5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(Q, n, D, generator=g).to(dev)
    W = (torch.randn(K, D, G, generator=g) * 0.2).to(dev)
    bias = torch.randn(n, G, generator=g).to(dev)
    dout = torch.randn(Q, n, G, generator=g).to(dev)
    whole = RowPartitionedLayer(L, K, D, G, rank=0, world=1, device=dev)
    ref_out = whole.forward(x, W, bias).clone()
    ref_dW, ref_db = whole.backward(dout)
    part = RowPartitionedLayer(L, K, D, G, rank=rank, world=world, device=dev)
    lo, hi = part.plan.lo, part.plan.hi
    out = part.forward(x[:, lo:hi].contiguous(), W, bias[lo:hi].contiguous())
    assert float((out - ref_out[:, lo:hi]).abs().max() / ref_out.abs().max()) < 1e-5
    dW, db = part.backward(dout[:, lo:hi].contiguous())
    assert float((dW - ref_dW).abs().max() / ref_dW.abs().max()) < 1e-4
    assert float((db - ref_db[lo:hi]).abs().max() / ref_db.abs().max()) < 1e-5
    return part.plan.n_halo


def check_halo_modes(rank, world, dev):
    """E. halo rows pulled over NVLink peer memory (tgcn_halo_signal / tgcn_halo_pull) == halo rows exchanged by NCCL
    send/recv: forward, dW and db bit for bit."""
    L, _ = wl.random_geometric(n=20000, mean_degree=10.0, seed=4)
    n = L.shape[0]
    Q, D, G, K = 2, 32, 32, 6
    g = torch.Generator().manual_seed(1)
    x = torch.randn(Q, n, D, generator=g).to(dev)
    W = (torch.randn(K, D, G, generator=g) * 0.2).to(dev)
    bias = torch.randn(n, G, generator=g).to(dev)
    dout = torch.randn(Q, n, G, generator=g).to(dev)
    res = []
    for mode in ("peer", "nccl"):
        part = RowPartitionedLayer(L, K, D, G, rank=rank, world=world, device=dev, halo=mode)
        lo, hi = part.plan.lo, part.plan.hi
        for _ in range(2):                                         # twice: the step flags / buffer parity advance
            out = part.forward(x[:, lo:hi].contiguous(), W, bias[lo:hi].contiguous()).clone()
            dW, db = part.backward(dout[:, lo:hi].contiguous())
        res.append((out, dW.clone(), db.clone()))
        torch.cuda.synchronize()
        dist.barrier()
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b), "peer-memory halo differs from the NCCL halo"
    return part.plan.n_halo


def check_fc1_exchange(rank, world, dev):
    """F. large head with the fused fc1 update, data parallel by ACTIVATION exchange (csrc/bighead.cu) == plain autograd
    on this rank's shard + allreduce(AVG) of the fc1 gradient + torch.optim.SGD; replicas bit-identical."""
    import torch.nn as nn
    from tgcn_b200.nn.head import Fc1FusedSGD, fused_head
    Q, I, Hd, C = 4, 16384, 200, 6
    torch.manual_seed(5)
    mods = [nn.Linear(I, Hd).to(dev), nn.BatchNorm1d(Hd).to(dev), nn.Linear(Hd, C).to(dev)]
    ref = [nn.Linear(I, Hd).to(dev), nn.BatchNorm1d(Hd).to(dev), nn.Linear(Hd, C).to(dev)]
    for m, r in zip(mods, ref):
        r.load_state_dict(m.state_dict())
        m.train(); r.train()
    upd = Fc1FusedSGD(mods[0].weight, lr=0.05, momentum=0.5, batch=Q)
    opt = torch.optim.SGD([ref[0].weight], lr=0.05, momentum=0.5)
    g = torch.Generator(device=dev).manual_seed(200 + rank)        # different samples per rank
    worst = 0.0
    for it in range(4):
        x = torch.randn(Q, I, device=dev, generator=g)
        y = torch.randint(0, C, (Q,), device=dev, generator=g)
        F.nll_loss(fused_head(x, *mods, fc1_update=upd), y).backward()
        assert mods[0].weight.grad is None
        h = F.relu(ref[1](ref[0](x)))
        F.nll_loss(F.log_softmax(ref[2](h), dim=1), y).backward()
        dist.all_reduce(ref[0].weight.grad, op=dist.ReduceOp.AVG)
        opt.step()
        # the other parameters are not stepped here: copy them so both sides stay comparable, and clear the gradients
        for m, r in zip(mods, ref):
            for pm, pr in zip(m.parameters(), r.parameters()):
                if pm is not mods[0].weight:
                    pm.grad = None
                pr.grad = None
        torch.cuda.synchronize()
        upd_mag = float((ref[0].weight - mods[0].weight).abs().max() / ref[0].weight.abs().max())
        worst = max(worst, upd_mag)
    assert worst < 1e-5, worst
    mine = mods[0].weight.detach().reshape(-1)
    others = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    assert all(torch.equal(o, mine) for o in others), "fc1 replicas diverged"
    return worst


def check_peer_sgd(rank, world, dev):
    """D. fused peer-memory allreduce + SGD == NCCL allreduce (mean) + torch SGD, and the replicas stay identical."""
    torch.manual_seed(3)                                           # same initial parameters on every rank
    shapes = [(10, 15, 32), (384, 32), (200, 1536), (6,)]
    pa = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    opt_a = PeerAllreduceSGD(pa, lr=0.05, momentum=0.5)
    opt_b = torch.optim.SGD(pb, lr=0.05, momentum=0.5)
    g = torch.Generator(device=dev).manual_seed(100 + rank)        # different gradients per rank
    for it in range(5):
        for p, q in zip(pa, pb):
            gr = torch.randn(p.shape, device=dev, generator=g)
            p.grad = gr.clone()
            avg = gr.clone()
            dist.all_reduce(avg, op=dist.ReduceOp.AVG)
            q.grad = avg
        opt_a.step(); opt_b.step()
    torch.cuda.synchronize()
    err = max(float((p - q).abs().max() / q.abs().max()) for p, q in zip(pa, pb))
    assert err < 1e-5, err
    mine = torch.cat([p.detach().reshape(-1) for p in pa])
    others = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    assert all(torch.equal(o, mine) for o in others), "replicas diverged"
    return err


def check_peer_sgd_overlap(rank, world, dev):
    """E. `late=`: the early group's exchange + update is launched from autograd hooks on a side stream during the
    backward, the late group at step(); same result as NCCL-averaged gradients + torch SGD, eager and graph-replayed."""
    def make():
        torch.manual_seed(5)
        return torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.ReLU(), torch.nn.Linear(64, 200),
                                   torch.nn.ReLU(), torch.nn.Linear(200, 6)).to(dev)
    ma, mb = make(), make()
    opt_a = PeerAllreduceSGD(ma.parameters(), lr=0.05, momentum=0.5, late=list(ma[0].parameters()))
    assert opt_a._early is not None and len(opt_a._main.params) == 2
    opt_b = torch.optim.SGD(mb.parameters(), lr=0.05, momentum=0.5)
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    xs = [torch.randn(32, 40, device=dev, generator=g) for _ in range(6)]
    x_static = xs[0].clone()

    def step_a():
        opt_a.zero_grad(set_to_none=True)
        ma(x_static).square().mean().backward()
        opt_a.step()

    def step_b(x):
        opt_b.zero_grad(set_to_none=True)
        mb(x).square().mean().backward()
        for q in mb.parameters():
            dist.all_reduce(q.grad, op=dist.ReduceOp.AVG)
        opt_b.step()

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for it in range(3):                                        # eager, hooks fire inside backward
            x_static.copy_(xs[it]); step_a()
    torch.cuda.current_stream().wait_stream(s)
    for it in range(3):
        step_b(xs[it])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):                                  # the whole step (both groups, fork/join) in one graph
        step_a()
    x_static.copy_(xs[3]); torch.cuda.synchronize()
    # the capture itself does not execute: replay for batches 3..5
    for it in range(3, 6):
        x_static.copy_(xs[it]); graph.replay(); step_b(xs[it])
    torch.cuda.synchronize()
    err = max(float((p - q).abs().max() / q.abs().max()) for p, q in zip(ma.parameters(), mb.parameters()))
    assert err < 1e-5, err
    mine = torch.cat([p.detach().reshape(-1) for p in ma.parameters()])
    others = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    assert all(torch.equal(o, mine) for o in others), "replicas diverged"
    return err


def main():
    rank, world, local = init_distributed("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if os.environ.get("MGPU_ONLY") == "peer":                     # second run of the test: TGCN_PEER_TWOSHOT=1 in the environment
        pe = check_peer_sgd(rank, world, dev)
        po = check_peer_sgd_overlap(rank, world, dev)
        dist.barrier()
        if rank == 0:
            print("MGPU_PEER_OK world=%d peer_sgd_err=%.2e peer_overlap_err=%.2e" % (world, pe, po))
        dist.destroy_process_group()
        return
    e = check_dp(rank, world, dev)
    h = check_halo(rank, world, dev)
    h2 = check_partitioned_layer(rank, world, dev)
    h3 = check_halo_modes(rank, world, dev)
    fe = check_fc1_exchange(rank, world, dev)
    pe = check_peer_sgd(rank, world, dev)
    po = check_peer_sgd_overlap(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d dp_err=%.2e halo_rows=%d layer_halo_rows=%d halo_modes_rows=%d fc1_exchange_err=%.2e peer_sgd_err=%.2e peer_overlap_err=%.2e"
              % (world, e, h, h2, h3, fe, pe, po))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
