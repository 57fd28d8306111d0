#!/usr/bin/env python
"""bench.py -- TGCN train samples/s on B200 (+ Chebyshev SpMM roofline, + CPU reference arm).

Contract (see DESIGN.md section "Measurement"):
    python bench.py --gpus N --steps K --warmup W [--workload hcp360|mesh32k|mnist] [--impl reference]
One "step" = forward + loss + backward (+ gradient allreduce for N > 1) + SGD update of the
reference's model for the workload on one batch of synthetic data.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "tgcn_train_samples_per_s"
ROWTILE_DEFAULT = 4      # rows per tile of the register-tiled SpMM on the streaming workloads (0: per-entry kernels)
UNIT = "samples/s"

WORKLOADS = {
    # name: (description, per-GPU batch, horizon, graph builder kwargs)
    "hcp360": dict(desc="BASELINE.json configs[1]: HCP-shaped parcellation graph (360 regions, top-16 sparse "
                        "connectome, 4 coarsening levels), T=15, batch 64 per GPU, NetTGCN_HCP "
                        "(TGCNCheb_H(1,32,K10,H15)->relu->pool4->GCNCheb(32,64,K10)->relu->pool4->fc200->bn->fc6)",
                   batch=64, H=15, model="hcp", classes=6),
    "hcp360dense": dict(desc="configs[1] with the fully dense connectome variant", batch=64, H=15, model="hcp",
                        classes=6),
    "mesh32k": dict(desc="BASELINE.json configs[2]: cortical-surface mesh (32492 vertices -> 41856 padded), T=30, "
                         "batch 8 per GPU, two TGCN layers + pooling", batch=8, H=30, model="hcp", classes=6),
    "rgg1m": dict(desc="BASELINE.json configs[3]: synthetic random geometric graph, 1M vertices (mean degree 12, x-sorted), "
                       "one TGCNCheb_H layer F=64 -> G=64, K=8, Kt=H=3, batch 1; forward + dW/db backward + SGD; graph rows "
                       "partitioned over the GPUs with a halo exchange before every recursion step (strong scaling)",
                  batch=1, H=3, model="rgg", classes=0),
    "mnist": dict(desc="BASELINE.json configs[0]: 28x28 8-NN grid, 4 coarsening levels, batch 100, H=12, "
                       "TGCNCheb_H(1,15,K10)->relu->fc10", batch=100, H=12, model="mnist", classes=10),
}


def build_graph(name):
    from tgcn_b200 import workloads as wl
    if name == "hcp360":
        return wl.hcp_parcellation(dense=False)
    if name == "hcp360dense":
        return wl.hcp_parcellation(dense=True)
    if name == "mesh32k":
        return wl.cortical_mesh()
    if name == "mnist":
        return wl.mnist_grid()
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region of the small workloads is a
    few milliseconds, far shorter than one `nvidia-smi` start-up, so the samples come from NVML in a thread of this
    process (one query is ~0.1 ms and releases the GIL); `nvidia-smi -lms` is only the fall-back when NVML cannot be
    loaded.  `mark()`/`unmark()` bracket the timed loops: `sm_mhz` is the median of the samples taken inside."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index, period_s=0.001):
        self.index = index
        self.period = period_s
        self.proc = None
        self.nvml = None
        self.handle = None
        self.lines = []
        self.samples = []          # (inside timed region, sm MHz, reason bit mask)
        self.inside = False
        self.running = False
        self.sm_max = None
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:                       # CUDA_VISIBLE_DEVICES may renumber: go through the PCI address of the torch device
            pr = torch.cuda.get_device_properties(self.index)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        self.inside = True

    def unmark(self):
        self.inside = False

    def _poll(self):
        n = self.nvml
        while self.running:
            try:
                inside = self.inside
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((inside and self.inside, mhz, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((self.inside, line.strip()))

    def stop(self):
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=2)
            n = self.nvml
            bits = ((n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                    (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                    (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                    (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
            inside = [s for s in self.samples if s[0]]
            use = inside if inside else self.samples
            reasons = sorted({nm for _, _, m in use for bit, nm in bits if m & bit})
            sm = [s[1] for s in use]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(inside), "samples_total": len(self.samples), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, n_in = [], [], set(), 0
        inside_any = any(i for i, _ in self.lines)
        for ins, ln in self.lines:
            if inside_any and not ins:
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            n_in += int(ins)
            for nm, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": n_in, "samples_total": len(self.lines), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_step_bytes(N, C, nnz, has_prev):
    S = 4 * N * C
    E = 8 * nnz + 4 * (N + 1)
    return (3 if has_prev else 2) * S + E


def run_b200(args):
    import torch.distributed as dist
    from tgcn_b200 import _lib, workloads as wl
    from tgcn_b200.parallel import GradientBucket, PeerAllreduceSGD, broadcast_parameters, init_distributed

    os.environ.pop("NCCL_DEBUG", None)      # its version banner goes to stdout; the contract is ONE JSON line
    rank, world, local = init_distributed("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    if lib.tgcn_device_supported() != 1:
        raise SystemExit("libtgcn_b200 needs an sm_100 device")
    cfg = WORKLOADS[args.workload]
    Q = args.batch or cfg["batch"]
    H = cfg["H"]

    graphs, perm, Ls, n_real = build_graph(args.workload)
    Lt = wl.as_torch_operands(Ls, device=dev)
    torch.manual_seed(0)
    if cfg["model"] == "hcp":
        model = wl.NetTGCN_HCP(Lt, horizon=H, n_classes=cfg["classes"], engine=args.engine).to(dev)
    else:
        model = wl.NetTGCN_MNIST(Lt, horizon=H, n_classes=cfg["classes"], engine=args.engine).to(dev)
    broadcast_parameters(model)
    N0 = Ls[0].shape[0]
    rowtile_info = []
    if args.rowtile and args.workload == "mesh32k":
        # streaming layers only (the resident kernels keep the operand in shared memory and never call the SpMM)
        for name in ("tgcn1", "gcn2"):
            lay = getattr(model, name)
            K_, G_ = lay.weight.shape[0], lay.weight.shape[-1]
            plan = lay._plan(dev)
            if not lay._use_resident(plan, lay.weight.numel() // (K_ * G_), G_, K_):
                rowtile_info += [dict(i[2], layer=name) for i in plan.ensure_rowtile_plans(rows_per_tile=args.rowtile, pad=args.rowtile_pad)]
    # the first layer's gradients are produced last: their (small) bucket is reduced after the others,
    # whose allreduce runs under the layer-1 backward
    # N > 1 (default): gradient allreduce fused with the SGD update over NVLink peer memory (csrc/peer.cu);
    # --dp nccl: two-bucket NCCL allreduce overlapped with the layer-1 backward + torch's fused SGD
    use_peer = args.dp in ("peer", "peer-overlap", "peer-serial")
    # the conv layers' gradients come last in the backward: with peer-overlap the head's (large) exchange + update is
    # launched from autograd hooks on a side stream under the conv backward, only the conv group remains at the end.
    # Measured (profiles/r01/scaling_hcp360.txt): +2.6 % at 8 GPUs, -2 % at 2 GPUs (two more launches) => auto by world
    overlap = args.dp == "peer-overlap" or (args.dp == "peer" and world >= 4)
    late = [] if not overlap else [p for n_, m in model.named_children() if n_ in ("tgcn1", "gcn2")
                                                for p in m.parameters()]
    if use_peer and world == 1:
        opt = PeerAllreduceSGD(model.parameters(), lr=0.01, momentum=0.5)     # world 1: one fused update launch
    elif use_peer:
        # collective decision: if CUDA IPC is unavailable on any rank, every rank falls back to the NCCL path
        try:
            opt = PeerAllreduceSGD(model.parameters(), lr=0.01, momentum=0.5, late=late)
            ok = torch.ones(1, device=dev)
        except Exception as exc:                       # noqa: BLE001
            sys.stderr.write("[bench] peer-memory allreduce unavailable on rank %d: %s\n" % (rank, exc))
            opt, ok = None, torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        use_peer = bool(ok.item() > 0)
    if use_peer:
        grads = None
    else:
        grads = GradientBucket(model.parameters(), late=list(model.tgcn1.parameters()))
        # pytorch_hcp_tgcn.py defaults (lr 0.01, momentum 0.5); torch's fused multi-tensor implementation: one launch
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.5, fused=True)

    # synthetic data: a few distinct pinned host batches per rank, cycled
    n_host = 4
    hx = [wl.synthetic_signals(Q, N0, H, n_real, perm, seed=1000 * rank + i).pin_memory() for i in range(n_host)]
    g = torch.Generator().manual_seed(77 + rank)
    hy = [torch.randint(0, cfg["classes"], (Q,), generator=g).pin_memory() for _ in range(n_host)]
    # two device input buffers: while step i runs, the copy stream uploads batch i+1 into the other one
    xs = [hx[0].to(dev), hx[1].to(dev)]
    ys = [hy[0].to(dev), hy[1].to(dev)]
    loss_dev = torch.zeros((), device=dev)
    loss_host = torch.zeros((), pin_memory=True)

    def fwd_bwd(b):
        opt.zero_grad(set_to_none=True)      # autograd allocates the gradients: no zero fill, no accumulate kernels
        out = model(xs[b])
        loss = F.nll_loss(out, ys[b])
        loss.backward()
        loss_dev.copy_(loss.detach())

    def finish_step():
        if grads is not None:
            grads.sync()                     # world > 1, --dp nccl: bucketed NCCL allreduce (average)
        opt.step()

    # warm-up (eager, side stream) then capture the whole training step (forward, loss, backward, gradient
    # allreduce, SGD update) as ONE CUDA graph per input buffer
    model.train()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            fwd_bwd(i % 2); finish_step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    launches_per_step = None
    if use_graph:
        graphs_ = []
        for b in range(2):
            g = torch.cuda.CUDAGraph()
            c0 = lib.tgcn_launch_count()
            with torch.cuda.graph(g):
                fwd_bwd(b)
                finish_step()
            launches_per_step = lib.tgcn_launch_count() - c0
            graphs_.append(g)

        def step(b=0):
            graphs_[b].replay()
    else:
        def step(b=0):
            fwd_bwd(b); finish_step()
        c0 = lib.tgcn_launch_count(); step(); launches_per_step = lib.tgcn_launch_count() - c0

    # L2 hygiene: flush between timed iterations when the step's working set fits in L2
    D1 = H
    work_bytes = 4 * (10 * N0 * Q * D1) * 2
    flush = work_bytes < 256e6
    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if flush else None   # 256 MB

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream()

    def timed(nsteps, e2e):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        copied = [torch.cuda.Event() for _ in range(nsteps + 1)]
        done = [torch.cuda.Event() for _ in range(nsteps)]
        main = torch.cuda.current_stream()

        def upload(i):                       # batch i -> device buffer i % 2, on the copy stream
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(done[i - 2])      # the buffer's previous consumer has finished
                xs[i % 2].copy_(hx[i % n_host], non_blocking=True)
                ys[i % 2].copy_(hy[i % n_host], non_blocking=True)
                copied[i].record(copy_stream)
        barrier()
        sampler.mark()
        t0 = time.perf_counter()
        for i in range(nsteps):
            if flush:
                flush_buf.fill_(float(i))
            ev[i][0].record()
            if e2e:
                if i == 0:
                    upload(0)
                if i + 1 < nsteps:
                    upload(i + 1)            # overlaps with this step's compute
                main.wait_event(copied[i])
            step(i % 2 if e2e else 0)
            if e2e:
                done[i].record(main)
                loss_host.copy_(loss_dev, non_blocking=True)
            ev[i][1].record()
            if e2e:
                ev[i][1].synchronize()         # the user reads the loss every step
                _ = float(loss_host)
        barrier()
        sampler.unmark()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    ms_total, wall = timed(args.steps, e2e=False)
    ms_e2e, wall_e2e = timed(args.steps, e2e=True)
    # back-to-back (no flush, no per-step events) for reference
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms_warm = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None

    roof = measure_spmm_roofline(lib, model, Ls, Q, H, dev, flush_buf) if (rank == 0 and not args.no_roofline) else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_throughput(args.workload, Ls, perm, n_real, cfg, budget_s=args.cpu_budget)

    if rank == 0:
        ms_step = ms_total / args.steps
        line = {
            "metric": METRIC, "value": world * Q / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": cfg["desc"], "per_gpu_batch": Q, "global_batch": world * Q,
                       "N_padded": [int(L.shape[0]) for L in Ls], "nnz_L0": int(Ls[0].nnz), "K": 10, "H": H,
                       "parallelism": "dp%d" % world, "engine": args.engine, "spmm_rowtile": rowtile_info or None,
                       "gradient_exchange": ("none" if world == 1 else
                                             ("peer-memory allreduce fused with SGD (NVLink P2P loads)" +
                                              ("; head group exchanged under the conv backward" if late else "")) if use_peer else
                                             "NCCL allreduce, 2 buckets"),
                       "l2": "flushed between timed steps (256 MB write)" if flush else "working set exceeds L2 (K-slab stack > 126 MB)",
                       "cuda_graph": use_graph,
                       "e2e_pipeline": "batch i+1 is uploaded (pinned host -> device, copy stream) while step i runs; the loss is read back and synchronised every step"},
            "e2e": {"value": world * Q / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(hx[0].numel() * 4 + hy[0].numel() * 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "ms_per_step_back_to_back": ms_warm / args.steps,
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "final_loss": float(loss_host),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # ProcessGroupNCCL teardown blocks while CUDA graphs that captured collectives are alive (observed on
        # torch 2.11 / NCCL 2.28: destroy_process_group never returns): synchronise, flush and leave.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def run_rgg(args):
    """Config 4: one large-graph layer, rows partitioned over the ranks (tgcn_b200.parallel.RowPartitionedLayer)."""
    import torch.distributed as dist
    from tgcn_b200 import _lib, workloads as wl
    from tgcn_b200.parallel import RowPartitionedLayer, init_distributed
    os.environ.pop("NCCL_DEBUG", None)
    rank, world, local = init_distributed("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    cfg = WORKLOADS["rgg1m"]
    n = args.rgg_n
    K, H, Fin, G, Q = 8, 3, 64, 64, args.batch or 1
    D = H * Fin
    L, _ = wl.random_geometric(n=n, mean_degree=12.0, seed=0)
    layer = RowPartitionedLayer(L, K, D, G, rank=rank, world=world, device=dev, rows_per_tile=args.rowtile,
                                rowtile_pad=args.rowtile_pad)
    n_own = layer.n_own
    gen = torch.Generator().manual_seed(0)
    bound = 1.0 / (Fin * K) ** 0.5
    W = ((torch.rand(K, D, G, generator=gen) * 2 - 1) * bound).to(dev)
    bias = ((torch.rand(n, G, generator=gen) * 2 - 1) * bound)[layer.plan.lo:layer.plan.hi].contiguous().to(dev)
    mW, mb = torch.zeros_like(W), torch.zeros_like(bias)
    hx = [torch.randn(Q, n_own, D, generator=torch.Generator().manual_seed(100 + rank + 7 * i)).pin_memory() for i in range(2)]
    x = hx[0].to(dev)
    loss_dev = torch.zeros((), device=dev)
    loss_host = torch.zeros((), pin_memory=True)
    lr, mom = 0.01, 0.5
    scale = 1.0 / (Q * n * G)

    def step():
        out = layer.forward(x, W, bias)
        loss_dev.copy_((out * out).sum() * (0.5 * scale))      # mean-square objective; rank-local part of the loss
        dW, db = layer.backward(out * scale)
        mW.mul_(mom).add_(dW); W.add_(mW, alpha=-lr)
        mb.mul_(mom).add_(db); bias.add_(mb, alpha=-lr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        barrier()
        sampler.mark()
        for i in range(nsteps):
            ev[i][0].record()
            if e2e:
                x.copy_(hx[i % 2], non_blocking=True)
            step()
            if e2e:
                loss_host.copy_(loss_dev, non_blocking=True)
            ev[i][1].record()
            if e2e:
                ev[i][1].synchronize()
                _ = float(loss_host)
        barrier()
        sampler.unmark()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = lib.tgcn_launch_count(); step(); launches = lib.tgcn_launch_count() - c0
    for _ in range(max(args.warmup, 3) - 1):
        step()
    ms = timed(args.steps, False)
    ms_e2e = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    roof = None
    if rank == 0:
        peak, src = _peak()
        C = Q * D
        b = layer._buffers(Q)

        def spmm_steps():
            st = torch.cuda.current_stream().cuda_stream
            for j in range(1, K):
                lib.tgcn_spmm_step(layer.rowptr.data_ptr(), layer.col.data_ptr(), layer.val.data_ptr(), n_own,
                                   b["stack"][j - 1].data_ptr(), None, b["stack"][j].data_ptr(), C, 1.0, 0.0, st)
        per = _time_graph(spmm_steps, 5, None) / (K - 1)
        nbytes = 2 * 4 * n_own * C + 8 * int(layer.col.numel()) + 4 * (n_own + 1)
        ach = nbytes / (per * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": ("spmm_step_rtile_kernel" if layer.rowtile else "spmm_step_csm_kernel") +
                                          " (recursion step on this rank's rows, 2S+E bytes; slabs exceed L2)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "bytes_per_launch": int(nbytes),
                "us_per_launch": per * 1e3, "peak_source": src}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_rgg_throughput(n, K, H, Fin, G, budget_s=args.cpu_budget)
    if rank == 0:
        ms_step = ms / args.steps
        line = {"metric": METRIC, "value": Q / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "rgg1m", "description": cfg["desc"], "vertices": n, "nnz": int(L.nnz), "K": K, "H": H,
                           "F": Fin, "G": G, "batch": Q, "parallelism": "rows/%d" % world, "halo_rows_rank0": int(layer.plan.n_halo),
                           "spmm_rowtile": layer.rowtile[2] if layer.rowtile else None,
                           "l2": "working set exceeds L2 (768 MB slabs at 1 GPU)", "cuda_graph": False},
                "e2e": {"value": Q / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(hx[0].numel() * 4),
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": int(launches), "clocks": clocks,
                "roofline": roof, "cpu_baseline": cpu, "final_loss": float(loss_host)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def cpu_rgg_throughput(n, K, H, Fin, G, budget_s=20.0, sample_n=50000):
    """The reference's CPU path (gcn_matmul.py slab form with a torch.sparse_csr L~, oracle port) on a bounded
    sample: the first `sample_n` vertices of the x-sorted graph (a strip); the rate is scaled by sample_n / n."""
    from oracle import model_torch
    from tgcn_b200 import workloads as wl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Ls, _ = wl.random_geometric(n=sample_n, mean_degree=12.0, seed=0)
    coo = Ls.tocoo()
    Lt = torch.sparse_coo_tensor(np.vstack([coo.row, coo.col]), coo.data.astype(np.float32), coo.shape).coalesce().to_sparse_csr()
    torch.manual_seed(0)
    lay = model_torch._Conv(Lt, (K, H, Fin, G), (1, sample_n, G), Fin * K)
    x = torch.randn(1, sample_n, H, Fin)
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        lay.zero_grad()
        out = lay(x)
        (out * out).mean().backward()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s or len(times) >= 10:
            break
    med = float(np.median(times[1:] or times))
    return {"value": (sample_n / n) / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d fwd+bwd steps of the torch-CPU port (sparse_csr slab path) on the first %d of %d vertices, batch 1, "
                      "%d threads, median %.2f s/step; value scaled by %d/%d" % (len(times), sample_n, n, cores, med, sample_n, n),
            "ms_per_step": med * 1e3 * (n / sample_n)}


def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def _time_graph(fn, reps, flush_buf):
    """Average device time of fn() (CUDA events on the launching stream, CUDA-graph replay so that host
    launch latency is not what gets timed on the small workloads, L2 flushed between replays)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    total = 0.0
    for r in range(reps):
        if flush_buf is not None:
            flush_buf.fill_(float(r))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record(); b.synchronize()
        total += a.elapsed_time(b)
    return total / reps


def measure_spmm_roofline(lib, model, Ls, Q, H, dev, flush_buf):
    """Roofline of the dominant kernel of layer 1 on the workload's own operand shapes.

    Streaming path (large graphs): the CSR SpMM recursion step, 2S+E algorithmic bytes per launch.
    Resident path (graphs whose per-sample slab fits in shared memory): the fused whole-layer forward
    kernel; its algorithmic bytes are SURVEY.md section 8(d)'s fused-layer figure
    B_fwd = (3K-4) S + (K-1) E + 4QNG + 4KHFG + 4NG (what a design that writes every T_k to HBM once would
    move); the kernel keeps the recursion in shared memory, so its real DRAM traffic is far lower
    (reported as `compulsory_bytes`: x + pooled output + indices + the saved basis)."""
    from tgcn_b200.nn import functional as Fn
    peak, src = _peak()
    lay = model.tgcn1
    plan = lay._plan(dev)
    N = plan.n
    K, Hh, Fin, G = lay.weight.shape
    D = Hh * Fin
    C = Q * D
    S = 4 * N * C
    E = 8 * plan.nnz + 4 * (N + 1)
    reps = 20
    if lay._use_resident(plan, D, G, K):
        x = torch.randn(Q, N, D, device=dev)
        w = lay.weight.detach().reshape(K, D, G).contiguous()
        bias = lay.bias.detach().contiguous()
        y = torch.empty(Q, N // 4, G, device=dev)
        idx = torch.empty(Q, N // 4, G, dtype=torch.uint8, device=dev)
        stack = torch.empty(int(lib.tgcn_resident_stack_bytes(Q, N, D, K)) // 4, device=dev)
        wimg = torch.empty(int(lib.tgcn_resident_weights_bytes(D, G, K)) // 4, device=dev)
        rowinfo, entries, Ep = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 0))

        def call():
            st = torch.cuda.current_stream().cuda_stream
            rc = lib.tgcn_resident_layer_fwd(rowinfo.data_ptr(), entries.data_ptr(), N, Ep, x.data_ptr(), w.data_ptr(),
                                             bias.data_ptr(), 1, None, y.data_ptr(), idx.data_ptr(), 4, 1, stack.data_ptr(),
                                             wimg.data_ptr(), Q, D, G, K, 0, st)
            assert rc == 0
        ms = _time_graph(call, reps, flush_buf)
        bytes_per_launch = (3 * K - 4) * S + (K - 1) * E + 4 * Q * N * G + 4 * K * D * G + 4 * N * G
        compulsory = 4 * Q * N * D + 5 * Q * (N // 4) * G + stack.numel() * 4 + 4 * K * D * G + 4 * N * G + E
        achieved = bytes_per_launch / (ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at the hcp360 layer-1 shape from the round's
        # `ncu --set full` capture (profiles/r01/resident_kernels_ncu.txt): the saved basis is still in L2 when the
        # kernel ends, so DRAM sees little more than the input read
        traffic = 1657088 + 49664 if (Q, N, D, G, K) == (64, 384, 15, 32, 10) else None
        return {"bound": "hbm", "kernel": "resident_fwd_kernel (layer 1: recursion + contraction + bias + ReLU + max-pool in one "
                                          "launch; timed together with its weight-image prologue kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "bytes_per_launch": int(bytes_per_launch), "bytes_formula": "SURVEY 8d B_fwd = (3K-4)S+(K-1)E+4QNG+4KHFG+4NG",
                "compulsory_bytes": int(compulsory), "us_per_launch": ms * 1e3, "peak_source": src}
    stack = torch.randn(K, N, C, device=dev)

    def steps():
        st = torch.cuda.current_stream().cuda_stream
        for k in range(1, K):
            lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N,
                               stack[k - 1].data_ptr(), None, stack[k].data_ptr(), C, 1.0, 0.0, st)
    per_launch_ms = _time_graph(steps, reps, flush_buf) / (K - 1)
    bytes_per_launch = algorithmic_step_bytes(N, C, plan.nnz, has_prev=False)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    kname = "spmm_step_rtile_kernel" if getattr(plan, "_rowtiles", None) else "spmm_step_pipe_kernel"
    return {"bound": "hbm", "kernel": kname + " (layer-1 recursion step, 2S+E bytes)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "bytes_per_launch": int(bytes_per_launch), "us_per_launch": per_launch_ms * 1e3, "peak_source": src}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port = the reference's ATen call sequence)
# ------------------------------------------------------------------------------------------------
def cpu_port_throughput(workload, Ls, perm, n_real, cfg, budget_s=20.0, steps=None, warmup=1):
    from oracle import model_torch
    from tgcn_b200 import workloads as wl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N0 = Ls[0].shape[0]
    dense_ok = N0 <= 4096
    Lt = wl.as_torch_operands(Ls, device=None, dense=dense_ok)
    if not dense_ok:
        Lt = [t.to_sparse_csr() for t in Lt]      # gcn_matmul.py slab path (dense L~ is infeasible at this N)
    torch.manual_seed(0)
    H = cfg["H"]
    Q = cfg["batch"]
    sample_q = Q if dense_ok else max(1, min(Q, 2))
    if cfg["model"] == "hcp":
        model = model_torch.PortNetTGCN_HCP(Lt, horizon=H, n_classes=cfg["classes"])
    else:
        model = model_torch.PortNetTGCN_MNIST(Lt, horizon=H, n_classes=cfg["classes"])
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.5)
    x = wl.synthetic_signals(sample_q, N0, H, n_real, perm, seed=5)
    y = torch.randint(0, cfg["classes"], (sample_q,))
    model.train()

    def one():
        opt.zero_grad()
        loss = F.nll_loss(model(x), y)
        loss.backward()
        opt.step()
        return float(loss.detach())
    for _ in range(warmup):
        one()
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter(); one(); times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and (time.perf_counter() - t_start > budget_s or len(times) >= 50):
            break
    med = float(np.median(times))
    return {"value": sample_q / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d timed train steps (fwd+loss+bwd+SGD) of the torch-CPU port of the reference model, batch %d of %d, "
                      "%s L~, %d threads, median %.1f ms/step" % (len(times), sample_q, Q,
                                                                   "dense" if dense_ok else "torch.sparse_csr (gcn_matmul slab path)",
                                                                   cores, med * 1e3),
            "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    if args.workload == "rgg1m":
        cpu = cpu_rgg_throughput(args.rgg_n, 8, 3, 64, 64, budget_s=max(args.cpu_budget, 20.0))
    else:
        graphs, perm, Ls, n_real = build_graph(args.workload)
        cpu = cpu_port_throughput(args.workload, Ls, perm, n_real, cfg, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.workload == "rgg1m" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": cfg["desc"], "per_gpu_batch": cfg["batch"]},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hcp360", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--engine", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the separate roofline timing loop (profiling runs)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--rgg-n", type=int, default=1_000_000, help="vertices of the rgg1m workload")
    ap.add_argument("--dp", default="peer", choices=["peer", "peer-overlap", "peer-serial", "nccl"],
                    help="gradient exchange for N > 1: peer = fused peer-memory allreduce+SGD (overlap variant from 4 GPUs); "
                         "peer-overlap = the head group's exchange runs under the conv backward; peer-serial = one exchange "
                         "after the backward; nccl = bucketed NCCL + torch SGD")
    ap.add_argument("--rowtile", type=int, default=ROWTILE_DEFAULT, choices=[0, 4, 8],
                    help="rows per tile of the register-tiled SpMM kernel for the streaming layers (0 = per-entry kernels)")
    ap.add_argument("--rowtile-pad", type=int, default=1,
                    help="pad every row tile's plan to a multiple of this many entries (1 = none, the measured configuration)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "rgg1m":
        run_rgg(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
