"""Multi-process (world_size 2, gloo, CPU) tests of the data-parallel and row-partition plumbing.

The host logic under test is tgcn_b200.parallel; the per-rank compute is done by the oracle
(CPU) because the CUDA kernels need a GPU -- the oracle is only the stand-in compute / checker here.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import csr_from, load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from tgcn_b200.parallel import init_distributed
    init_distributed("gloo")
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _dp_worker(rank, world):
    import torch.nn.functional as F
    from oracle import model_torch
    from tgcn_b200.parallel import FlatGradients, broadcast_parameters, shard_range
    r = load_golden("graph_grid28_k8_seed0.npz")        # 992 / 496 / 248 vertices: divisible by the two pool4
    Ls = [torch.tensor(np.asarray(csr_from(r, "L_%d" % i).todense()), dtype=torch.float) for i in range(3)]
    torch.manual_seed(100 + rank)                       # deliberately different init per rank
    model = model_torch.PortNetTGCN_HCP(Ls, horizon=5, K=4, g1=6, g2=8, hidden=10, n_classes=3, drop1=0.0, drop2=0.0)
    broadcast_parameters(model, src=0)
    gen = torch.Generator().manual_seed(7)
    Q = 8
    x = torch.randn(Q, Ls[0].shape[0], 5, generator=gen)
    y = torch.randint(0, 3, (Q,), generator=gen)
    model.eval()                                        # BN in eval mode: batch statistics are not sharded
    # single-process result on the full batch (every rank computes it for reference)
    ref_model = model_torch.PortNetTGCN_HCP(Ls, horizon=5, K=4, g1=6, g2=8, hidden=10, n_classes=3, drop1=0.0, drop2=0.0)
    ref_model.load_state_dict(model.state_dict()); ref_model.eval()
    F.nll_loss(ref_model(x), y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
    # sharded: per-rank mean loss, one flat allreduce, divide by world
    grads = FlatGradients(model.parameters())
    lo, hi = shard_range(Q, rank, world)
    grads.zero_()
    F.nll_loss(model(x[lo:hi]), y[lo:hi]).backward()
    flat = grads.allreduce_mean()
    assert flat.data_ptr() == grads.flat.data_ptr()
    for p in model.parameters():                        # grads are views of the flat buffer
        assert p.grad.data_ptr() >= grads.flat.data_ptr()
    err = float((flat - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err


def test_data_parallel_flat_allreduce_matches_single_process():
    _run(_dp_worker, 2)


def _bucket_worker(rank, world):
    """GradientBucket: autograd-allocated gradients (set_to_none), one packed collective, p.grad re-bound to
    the averaged flat views; an SGD step on every rank then gives identical parameters."""
    import torch.nn.functional as F
    from oracle import model_torch
    from tgcn_b200.parallel import GradientBucket, broadcast_parameters, shard_range
    r = load_golden("graph_grid28_k8_seed0.npz")
    Ls = [torch.tensor(np.asarray(csr_from(r, "L_%d" % i).todense()), dtype=torch.float) for i in range(3)]
    torch.manual_seed(5 + rank)
    model = model_torch.PortNetTGCN_HCP(Ls, horizon=5, K=4, g1=6, g2=8, hidden=10, n_classes=3, drop1=0.0, drop2=0.0)
    broadcast_parameters(model, src=0)
    model.eval()
    gen = torch.Generator().manual_seed(3)
    Q = 6
    x = torch.randn(Q, Ls[0].shape[0], 5, generator=gen)
    y = torch.randint(0, 3, (Q,), generator=gen)
    ref_model = model_torch.PortNetTGCN_HCP(Ls, horizon=5, K=4, g1=6, g2=8, hidden=10, n_classes=3, drop1=0.0, drop2=0.0)
    ref_model.load_state_dict(model.state_dict()); ref_model.eval()
    F.nll_loss(ref_model(x), y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
    bucket = GradientBucket(model.parameters(), late=list(model.tgcn1.parameters()))    # two buckets, early one overlapped
    assert bucket.late and bucket.early and len(bucket._hooks) == len(bucket.early)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    opt.zero_grad(set_to_none=True)
    lo, hi = shard_range(Q, rank, world)
    F.nll_loss(model(x[lo:hi]), y[lo:hi]).backward()
    bucket.sync()
    for p in model.parameters():
        assert p.grad.data_ptr() >= bucket.flat.data_ptr()
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])     # model order (the flat buffer is early|late)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-5
    opt.step()
    mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    assert all(torch.equal(o, mine) for o in other)       # replicas stay bit-identical


def test_gradient_bucket_keeps_replicas_identical():
    _run(_bucket_worker, 2)


def _halo_worker(rank, world):
    import scipy.sparse as sp
    from tgcn_b200.parallel import RowPartition, halo_exchange
    r = load_golden("graph_grid28_k8_seed0.npz")
    L = csr_from(r, "L_0")
    n = L.shape[0]
    rng = np.random.default_rng(0)
    C, K = 12, 5
    x = rng.standard_normal((n, C)).astype(np.float32)
    plan = RowPartition(L, rank, world).build_send_lists()
    local = sp.csr_matrix((plan.val, plan.col, plan.rowptr), shape=(plan.n_own, plan.n_own + plan.n_halo))
    cur = torch.tensor(x[plan.lo:plan.hi])
    full = x.copy()
    for _ in range(1, K):
        halo = halo_exchange(cur, plan)
        ext = np.concatenate([cur.numpy(), halo.numpy()], axis=0)
        cur = torch.tensor(np.asarray(local @ ext, dtype=np.float32))
        full = np.asarray(L @ full, dtype=np.float32)
        # same per-row summation order => bit-for-bit equal to the unpartitioned product
        assert np.array_equal(cur.numpy(), full[plan.lo:plan.hi])


def test_row_partition_halo_exchange_matches_unpartitioned():
    _run(_halo_worker, 2)


def test_row_partition_plan_single_process():
    import scipy.sparse as sp
    from tgcn_b200.parallel import RowPartition, shard_range
    r = load_golden("graph_rand300_seed2.npz")
    L = csr_from(r, "L_0")
    n = L.shape[0]
    for world in (1, 2, 3, 4):
        plans = RowPartition.exchange_plans([RowPartition(L, p, world) for p in range(world)])
        x = np.random.default_rng(1).standard_normal((n, 3)).astype(np.float32)
        want = np.asarray(L @ x)
        covered = 0
        for p in plans:
            halo = x[p.halo_ids]
            # what the owners would send really is what this rank expects
            off = 0
            for q in range(world):
                cnt = p.recv_cnt[q]
                if cnt:
                    sent = x[plans[q].lo:plans[q].hi][plans[q].send_idx[p.rank]]
                    assert np.array_equal(sent, halo[off:off + cnt])
                off += cnt
            local = sp.csr_matrix((p.val, p.col, p.rowptr), shape=(p.n_own, p.n_own + p.n_halo))
            got = np.asarray(local @ np.concatenate([x[p.lo:p.hi], halo]))
            assert np.array_equal(got, want[p.lo:p.hi])
            covered += p.n_own
            assert (p.lo, p.hi) == shard_range(n, p.rank, world)
        assert covered == n


def test_shard_range_partitions_everything():
    from tgcn_b200.parallel import shard_range
    for total in (0, 1, 7, 64, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
