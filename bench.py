#!/usr/bin/env python
"""bench.py -- TGCN train samples/s on B200 (+ Chebyshev SpMM roofline, + CPU reference arm).

Contract (DESIGN.md section "Measurement"):
    python bench.py --gpus N --steps K --warmup W [--workload mesh32k|hcp360|hcp360dense|mnist|rgg1m] [--impl reference]
One "step" = forward + loss + backward (+ gradient exchange for N > 1) + SGD update of the reference's model for the
workload (pytorch_hcp_tgcn.py:93-169: real-FFT prologue, two conv layers with ReLU / dropout / pooling, dense head) on
one batch of synthetic data.  Prints ONE JSON line (rank 0).

Default workload: `mesh32k` = BASELINE.json configs[2], the configuration the metric's targets live on (data-parallel
over the GPUs; its dominant kernel is the streaming Chebyshev SpMM).  At N = 1 the parcellation workload `hcp360`
(configs[1]) is run as well and reported under "secondary" in the same line.
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
UNIT = "samples/s"
ROWTILE_DEFAULT = 4      # rows per tile of the register-tiled SpMM on the streaming workloads (0: per-entry kernels)
LR, MOMENTUM = 0.01, 0.5  # pytorch_hcp_tgcn.py defaults (--lr 0.01 --momentum 0.5)

WORKLOADS = {
    "hcp360": dict(desc="BASELINE.json configs[1]: HCP-shaped parcellation graph (360 regions, top-16 sparse connectome, 4 "
                        "coarsening levels), T=15, batch 64 per GPU, NetTGCN of pytorch_hcp_tgcn.py:93-155 (real-FFT prologue, "
                        "TGCNCheb_H(1,32,K10,H15)->relu->drop0.1->pool4->GCNCheb(32,64,K10)->relu->pool4->fc200->bn->relu->drop0.5->fc6)",
                   batch=64, H=15, K=10, model="hcp", classes=6),
    "hcp360dense": dict(desc="configs[1] with the fully dense connectome variant", batch=64, H=15, K=10, model="hcp", classes=6),
    "mesh32k": dict(desc="BASELINE.json configs[2]: cortical-surface mesh (32492 vertices -> 41856 padded), T=30, batch 8 per "
                         "GPU, NetTGCN of pytorch_hcp_tgcn.py:93-155 (real-FFT prologue, two Chebyshev layers K=10 with ReLU / "
                         "dropout / pool4, fc 167424->200 -> bn -> relu -> dropout -> fc6), data-parallel over the GPUs",
                    batch=8, H=30, K=10, model="hcp", classes=6),
    "rgg1m": dict(desc="BASELINE.json configs[3]: synthetic random geometric graph, 1M vertices (mean degree 12, strip+Morton "
                       "order), one TGCNCheb_H layer F=64 -> G=64, K=8, Kt=H=3, batch 1; forward + dW/db backward + SGD; graph "
                       "rows partitioned over the GPUs with a halo exchange before every recursion step (strong scaling)",
                  batch=1, H=3, K=8, model="rgg", classes=0),
    "mnist": dict(desc="BASELINE.json configs[0]: 28x28 8-NN grid, 4 coarsening levels, batch 100, H=12, "
                       "TGCNCheb_H(1,15,K10)->relu->fc10", batch=100, H=12, K=10, model="mnist", classes=10),
}


def build_graph(name):
    """Product-side graph build (tgcn_b200.graph / .coarsening, native pairing kernel)."""
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


def build_graph_oracle(name):
    """The SAME graphs (bit-identical adjacency, permutation and rescaled Laplacians -- tests/test_oracle_workloads.py)
    built WITHOUT the product: tgcn_b200.synth is pure numpy, graph + coarsening come from oracle/ (restatements of
    gcn/graph.py and gcn/coarsening.py).  Used by the CPU arms so that they never map libtgcn_b200.so."""
    from oracle import coarsening_np as C, graph_np as Gn
    from tgcn_b200 import synth
    if name in ("hcp360", "hcp360dense"):
        A, n_real = synth.hcp_adjacency(dense=(name == "hcp360dense")), 360
    elif name == "mesh32k":
        A, n_real = synth.mesh_adjacency(), 32492
    elif name == "mnist":
        dist, idx = Gn.knn_exact(Gn.grid_embedding(28), k=8, metric="euclidean")
        A, n_real = Gn.knn_adjacency(dist, idx), 784
    else:
        raise ValueError(name)
    np.random.seed(0)
    graphs, perm = C.coarsen(A, levels=4, self_connections=False)[:2]
    Ls = []
    for g in graphs:
        L = Gn.rescale_laplacian(Gn.laplacian(g, normalized=True), 2).tocsr()
        L.eliminate_zeros()
        Ls.append(L.astype(np.float32))
    return graphs, perm, Ls, n_real


def workload_config(workload, world, Q, Ls, H):
    """Keys that describe the WORKLOAD (identical in the b200 arm and the reference arm)."""
    cfg = WORKLOADS[workload]
    return {"workload": workload, "description": cfg["desc"], "per_gpu_batch": Q, "global_batch": world * Q,
            "N_padded": [int(L.shape[0]) for L in Ls], "nnz_L0": int(Ls[0].nnz), "K": cfg["K"], "H": H,
            "dropout": [0.1, 0.5] if cfg["model"] == "hcp" else None, "time_dft": cfg["model"] == "hcp",
            "optimizer": "SGD lr %g momentum %g" % (LR, MOMENTUM), "parallelism": "dp%d" % world,
            "l2": "working set exceeds L2 (126 MB) or L2 flushed between timed steps (256 MB write); see impl_detail"}


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


def _peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def _ncu_traffic(kernel, shape_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from a committed `ncu --set full` capture
    (profiles/r02/ncu_traffic.json: {kernel: {shape_key: bytes}}), or None when no capture of this kernel at this
    shape has been committed -- never a literal in the code."""
    path = os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")
    try:
        table = json.load(open(path))
        return int(table[kernel][shape_key])
    except Exception:
        return None


def _time_graph(fn, reps, flush_buf):
    """Average device time of fn() (CUDA events on the launching stream, CUDA-graph replay so that host launch latency
    is not what gets timed on the small workloads, L2 flushed between replays when flush_buf is given)."""
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


def init_dist(args):
    from tgcn_b200.parallel import init_distributed
    os.environ.pop("NCCL_DEBUG", None)      # its version banner goes to stdout; the contract is ONE JSON line
    rank, world, local = init_distributed("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    return rank, world, local


def run_model(args, workload, rank, world, local, primary=True):
    """One model workload (hcp360 / mesh32k / mnist) on this process group; returns the JSON line as a dict (rank 0)."""
    import torch.distributed as dist
    from tgcn_b200 import _lib, workloads as wl
    from tgcn_b200.nn.head import Fc1FusedSGD
    from tgcn_b200.parallel import GradientBucket, PeerAllreduceSGD, broadcast_parameters

    dev = torch.device("cuda", local)
    lib = _lib.load()
    if lib.tgcn_device_supported() != 1:
        raise SystemExit("libtgcn_b200 needs an sm_100 device")
    cfg = WORKLOADS[workload]
    Q = (args.batch if primary else 0) or cfg["batch"]
    H = cfg["H"]

    graphs, perm, Ls, n_real = build_graph(workload)
    Lt = wl.as_torch_operands(Ls, device=dev)
    torch.manual_seed(0)
    if cfg["model"] == "hcp":
        model = wl.NetTGCN_HCP(Lt, horizon=H, K=cfg["K"], n_classes=cfg["classes"], engine=args.engine).to(dev)
    else:
        model = wl.NetTGCN_MNIST(Lt, horizon=H, K=cfg["K"], n_classes=cfg["classes"], engine=args.engine).to(dev)
    broadcast_parameters(model)
    N0 = Ls[0].shape[0]
    rowtile_info = []
    if args.rowtile:
        for name in ("tgcn1", "gcn2"):
            lay = getattr(model, name, None)
            if lay is None:
                continue
            K_, G_ = lay.weight.shape[0], lay.weight.shape[-1]
            plan = lay._plan(dev)
            # streaming layers only (the resident kernels keep the operand in shared memory and never call the SpMM)
            if not lay._use_resident(plan, lay.weight.numel() // (K_ * G_), G_, K_):
                rowtile_info += [dict(i[2], layer=name) for i in plan.ensure_rowtile_plans(rows_per_tile=args.rowtile, pad=args.rowtile_pad)]

    # ---- optimizer: fc1.weight of a large head is updated inside the backward (csrc/bighead.cu; for N > 1 the ranks
    # exchange activations, not the 134 MB gradient); everything else goes through the fused peer-memory allreduce + SGD
    params = list(model.parameters())
    fc1_fused = False
    if cfg["model"] == "hcp" and not args.no_fused_fc1 and args.dp != "nccl":
        w1 = model.fc1.weight
        if lib.tgcn_head_fused_update_supported(Q, int(w1.shape[1]), int(w1.shape[0])):
            model.fc1_update = Fc1FusedSGD(w1, lr=LR, momentum=MOMENTUM, batch=Q)
            params = [p for p in params if p is not w1]
            fc1_fused = True
    use_peer = args.dp in ("peer", "peer-overlap", "peer-serial")
    overlap = args.dp == "peer-overlap" or (args.dp == "peer" and world >= 4 and not fc1_fused)
    late = [] if not overlap else [p for n_, m in model.named_children() if n_ in ("tgcn1", "gcn2") for p in m.parameters()]
    if use_peer and world == 1:
        opt = PeerAllreduceSGD(params, lr=LR, momentum=MOMENTUM)
    elif use_peer:
        try:
            opt = PeerAllreduceSGD(params, lr=LR, momentum=MOMENTUM, late=late)
            ok = torch.ones(1, device=dev)
        except Exception as exc:                       # noqa: BLE001
            sys.stderr.write("[bench] peer-memory allreduce unavailable on rank %d: %s\n" % (rank, exc))
            opt, ok = None, torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        use_peer = bool(ok.item() > 0)
    if use_peer:
        grads = None
        if hasattr(model, "set_dropout_step"):
            model.set_dropout_step(opt.state)          # element 0 = steps taken: a new dropout mask every step, no extra launch
    else:
        grads = GradientBucket(params, late=list(model.tgcn1.parameters()))
        opt = torch.optim.SGD(params, lr=LR, momentum=MOMENTUM, fused=True)

    # synthetic data: a few distinct pinned host batches per rank, cycled
    n_host = 4
    hx = [wl.synthetic_signals(Q, N0, H, n_real, perm, seed=1000 * rank + i).pin_memory() for i in range(n_host)]
    g = torch.Generator().manual_seed(77 + rank)
    hy = [torch.randint(0, cfg["classes"], (Q,), generator=g).pin_memory() for _ in range(n_host)]
    xs = [hx[0].to(dev), hx[1].to(dev)]     # two device input buffers: batch i+1 is uploaded while step i runs
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
            grads.sync()
        opt.step()

    model.train()
    # ---- data-parallel check (N > 1): ONE real N-rank step with the dropouts off must leave every rank with
    # bit-identical parameters, equal (to fp32 rounding) to what a single process computes from the N rank batches:
    # per-rank gradients averaged, then torch.optim.SGD.  This is the in-bench proof of csrc/peer.cu / bighead.cu.
    dp_check = None
    if world > 1 and not args.no_dp_check:
        dp_check = dp_verify(model, opt, fwd_bwd, finish_step, wl, cfg, Lt, hx, hy, Q, N0, H, n_real, perm, rank, world, dev, args)

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            fwd_bwd(i % 2); finish_step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    if use_graph:
        graphs_ = []
        for b in range(2):
            gph = torch.cuda.CUDAGraph()
            c0 = lib.tgcn_launch_count()
            with torch.cuda.graph(gph):
                fwd_bwd(b)
                finish_step()
            launches_per_step = lib.tgcn_launch_count() - c0
            graphs_.append(gph)

        def step(b=0):
            graphs_[b].replay()
    else:
        def step(b=0):
            fwd_bwd(b); finish_step()
        c0 = lib.tgcn_launch_count(); step(); launches_per_step = lib.tgcn_launch_count() - c0

    # L2 hygiene: flush between timed iterations when the step's working set fits in L2
    work_bytes = 4 * (10 * N0 * Q * H) * 2
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
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    ms_total = timed(args.steps, e2e=False)
    ms_e2e = timed(args.steps, e2e=True)
    barrier()                                 # back-to-back (no flush, no per-step events) for reference
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms_warm = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None

    roof = measure_roofline(lib, model, Q, H, dev, flush_buf) if (rank == 0 and not args.no_roofline) else None
    cpu = None
    if rank == 0 and world == 1 and primary and not args.no_cpu_baseline:
        cpu = cpu_port_throughput(workload, cfg, budget_s=args.cpu_budget)
    if rank != 0:
        return None
    ms_step = ms_total / args.steps
    line = {
        "metric": METRIC, "value": world * Q / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, world, Q, Ls, H),
        "impl_detail": {
            "engine": args.engine, "spmm_rowtile": rowtile_info or None, "cuda_graph": use_graph,
            "fc1_update": "fused into the head backward (csrc/bighead.cu)" + ("; ranks exchange x / dh over NVLink peer memory, "
                          "the fc1 gradient is never materialised or reduced" if world > 1 else "") if fc1_fused else "regular optimizer",
            "gradient_exchange": ("none" if world == 1 else
                                  ("peer-memory allreduce fused with SGD (NVLink P2P; one-shot, or reduce-scatter + pushed "
                                   "all-gather from 4 ranks and 2 MB)" + ("; head group exchanged under the conv backward" if late else ""))
                                  if use_peer else "NCCL allreduce, 2 buckets"),
            "l2": "flushed between timed steps (256 MB write)" if flush else "working set exceeds L2 (K-slab stack > 126 MB)",
            "e2e_pipeline": "batch i+1 is uploaded (pinned host -> device, copy stream) while step i runs; the loss is read back "
                            "and synchronised every step"},
        "e2e": {"value": world * Q / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(hx[0].numel() * 4 + hy[0].numel() * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "ms_per_step_back_to_back": ms_warm / args.steps,
        "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "dp_check": dp_check,
        "final_loss": float(loss_host),
    }
    return line


def dp_verify(model, opt, fwd_bwd, finish_step, wl, cfg, Lt, hx, hy, Q, N0, H, n_real, perm, rank, world, dev, args):
    """See run_model.  Returns {"replicas_bit_identical": bool, "max_rel_err_vs_single_process": float} on rank 0 (the
    run aborts when the replicas differ or the error exceeds 1e-4)."""
    import copy
    import torch.distributed as dist
    drops = [m for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    saved_p = [m.p for m in drops]
    for m in drops:
        m.p = 0.0
    ref_model = None
    if rank == 0:                                    # single-process emulation: same initial parameters
        upd = getattr(model, "fc1_update", None)
        model.fc1_update = None
        ref_model = copy.deepcopy(model)
        model.fc1_update = upd
        ref_model.fused_head = False                 # plain torch head: independent of csrc/head.cu / bighead.cu
    fwd_bwd(0)
    finish_step()
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    # bit-identical replicas: compare an exact integer checksum of the parameter bit patterns
    bits = flat.view(torch.int32).to(torch.int64)
    chk = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    identical = all(torch.equal(allc[0], c) for c in allc)
    out = None
    if rank == 0:
        ref_model.train()
        accum = [torch.zeros_like(p) for p in ref_model.parameters()]
        for r in range(world):
            xr = wl.synthetic_signals(Q, N0, H, n_real, perm, seed=1000 * r + 0).to(dev)
            yr = torch.randint(0, cfg["classes"], (Q,), generator=torch.Generator().manual_seed(77 + r)).to(dev)
            for p in ref_model.parameters():
                p.grad = None
            F.nll_loss(ref_model(xr), yr).backward()
            for a, p in zip(accum, ref_model.parameters()):
                if p.grad is not None:
                    a.add_(p.grad)
        ro = torch.optim.SGD(ref_model.parameters(), lr=LR, momentum=MOMENTUM)
        old = [p.detach().clone() for p in ref_model.parameters()]
        for a, p in zip(accum, ref_model.parameters()):
            p.grad = a / world
        ro.step()
        worst, worst_upd = 0.0, 0.0
        for p, q, q0 in zip(model.parameters(), ref_model.parameters(), old):
            scale = float(q.detach().abs().max()) or 1.0
            worst = max(worst, float((p.detach() - q.detach()).abs().max()) / scale)
            dref = q.detach() - q0
            dscale = float(dref.abs().max())
            if dscale > 64 * 1.2e-7 * scale:           # updates below fp32 resolution of the parameter carry no signal
                worst_upd = max(worst_upd, float(((p.detach() - q0) - dref).abs().max()) / dscale)
        out = {"replicas_bit_identical": bool(identical), "max_rel_err_params": worst, "max_rel_err_update": worst_upd,
               "ranks": world, "what": "one N-rank step (dropouts off) vs per-rank gradients averaged + torch.optim.SGD in one "
                                       "process; parameters, and the parameter UPDATE relative to its own size"}
        if not identical or worst > 1e-4 or worst_upd > 0.05:
            raise SystemExit("[bench] data-parallel check FAILED: %s" % json.dumps(out))
    for m, p in zip(drops, saved_p):
        m.p = p
    dist.barrier()
    return out


def _rtile_kernel_name(R):
    """Which row-tile SpMM kernel the library launches (csrc/spmm.cu dispatch on the SPMM_RTILE tuning value)."""
    mode = int(os.environ.get("TGCN_SPMM_RTILE", "-1") or -1)
    mode = 2 if mode < 0 else mode
    return "spmm_step_rtile_pipe_kernel" if (mode == 3 or (mode == 2 and R == 4)) else "spmm_step_rtile_kernel"


def measure_roofline(lib, model, Q, H, dev, flush_buf):
    """Roofline of the dominant kernel of layer 1 on the workload's own operand shapes, timed LIVE with CUDA events.

    Streaming path (large graphs): the CSR SpMM recursion step, 2S+E algorithmic bytes per launch, HBM-bound.
    Resident path (graphs whose per-sample slab fits in shared memory): the fused whole-layer forward kernel keeps the
    recursion in shared memory; it is bound by shared-memory operand delivery and latency, not by HBM, so its line is
    labelled "smem/latency" and `achieved` counts its COMPULSORY DRAM bytes (x + pooled output + indices + saved basis
    + operands); SURVEY's unfused-design figure B_fwd is kept as `unfused_bytes_formula` for context."""
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
                                             bias.data_ptr(), 1, None, y.data_ptr(), idx.data_ptr(), 4, 1, None, stack.data_ptr(),
                                             wimg.data_ptr(), Q, D, G, K, 0, st)
            assert rc == 0
        ms = _time_graph(call, reps, flush_buf)
        unfused = (3 * K - 4) * S + (K - 1) * E + 4 * Q * N * G + 4 * K * D * G + 4 * N * G
        compulsory = 4 * Q * N * D + 5 * Q * (N // 4) * G + stack.numel() * 4 + 4 * K * D * G + 4 * N * G + E
        achieved = compulsory / (ms * 1e-3) / 1e9
        return {"bound": "smem/latency", "kernel": "resident_fwd_kernel (layer 1: recursion + contraction + bias + ReLU + max-pool in "
                                                   "one launch; timed together with its weight-image prologue kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic("resident_fwd_kernel", "Q%d_N%d_D%d_G%d_K%d" % (Q, N, D, G, K)),
                "bytes_per_launch": int(compulsory), "bytes_formula": "compulsory DRAM bytes: x + pooled y + idx + saved basis + W + bias + CSR",
                "unfused_bytes_formula": "SURVEY 8d B_fwd = (3K-4)S+(K-1)E+4QNG+4KHFG+4NG = %d" % unfused,
                "us_per_launch": ms * 1e3, "peak_source": src}
    # the launch the layer really issues: slabs of Q * Dp columns (Dp = D padded to whole 128-byte blocks for the TMA-fed
    # contraction, mesh layer 1: 30 -> 32); the ALGORITHMIC bytes stay those of the workload's own D
    try:
        from tgcn_b200.nn.gcn import _ENGINE
        Dp = max(D, int(lib.tgcn_layer_slab_width(Q, N, D, G, K, _ENGINE[lay.engine])))
    except Exception:          # measurement code must not take the bench line down: fall back to the unpadded launch
        Dp = D
    Cp = Q * Dp
    stack = torch.randn(K, N, Cp, device=dev)

    def steps():
        st = torch.cuda.current_stream().cuda_stream
        for k in range(1, K):
            lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N,
                               stack[k - 1].data_ptr(), None, stack[k].data_ptr(), Cp, 1.0, 0.0, st)
    per_launch_ms = _time_graph(steps, reps, flush_buf) / (K - 1)
    bytes_per_launch = algorithmic_step_bytes(N, C, plan.nnz, has_prev=False)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    tiles = getattr(plan, "_rowtiles", None)
    kname = _rtile_kernel_name(tiles[0][2]["rows_per_tile"]) if tiles else "spmm_step_pipe_kernel"
    return {"bound": "hbm", "kernel": kname + " (layer-1 recursion step; K-1 launches per layer forward)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": _ncu_traffic(kname, "N%d_C%d" % (N, Cp)),
            "bytes_per_launch": int(bytes_per_launch), "bytes_formula": "SURVEY 8d per step: 2S + E (S = 4NC slab, E = 8 nnz + 4(N+1))",
            "traffic_note": "ncu dram bytes of one launch of this kernel at this shape (profiles/r02/ncu_traffic.json); slabs that fit "
                            "the 126 MB L2 are still there when the next launch reads them, so traffic can be below the algorithmic bytes",
            "launched_columns": Cp, "algorithmic_columns": C,
            "us_per_launch": per_launch_ms * 1e3, "peak_source": src}


def run_rgg(args, rank, world, local):
    """Config 4: one large-graph layer, rows partitioned over the ranks (tgcn_b200.parallel.RowPartitionedLayer)."""
    import torch.distributed as dist
    from tgcn_b200 import _lib, workloads as wl
    from tgcn_b200.parallel import RowPartitionedLayer
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
    xs = [hx[0].to(dev), hx[1].to(dev)]
    loss_dev = torch.zeros((), device=dev)
    loss_host = torch.zeros((), pin_memory=True)
    scale = 1.0 / (Q * n * G)

    def step(b=0):
        out = layer.forward(xs[b], W, bias)
        loss_dev.copy_((out * out).sum() * (0.5 * scale))      # mean-square objective; rank-local part of the loss
        dW, db = layer.backward(out * scale)
        mW.mul_(MOMENTUM).add_(dW); W.add_(mW, alpha=-LR)
        mb.mul_(MOMENTUM).add_(db); bias.add_(mb, alpha=-LR)

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

        def upload(i):                       # the 768 MB input of step i+1 is uploaded on the copy stream under step i
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(done[i - 2])
                xs[i % 2].copy_(hx[i % 2], non_blocking=True)
                copied[i].record(copy_stream)
        barrier()
        sampler.mark()
        for i in range(nsteps):
            ev[i][0].record()
            if e2e:
                if i == 0:
                    upload(0)
                if i + 1 < nsteps:
                    upload(i + 1)
                main.wait_event(copied[i])
            step(i % 2 if e2e else 0)
            if e2e:
                done[i].record(main)
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
        kname = _rtile_kernel_name(args.rowtile) if layer.rowtile else "spmm_step_csm_kernel"
        roof = {"bound": "hbm", "kernel": kname + " (recursion step on this rank's rows, 2S+E bytes; slabs exceed L2)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": _ncu_traffic(kname, "N%d_C%d" % (n_own, C)), "bytes_per_launch": int(nbytes),
                "us_per_launch": per * 1e3, "peak_source": src}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_rgg_throughput(n, K, H, Fin, G, budget_s=args.cpu_budget)
    if rank != 0:
        return None
    ms_step = ms / args.steps
    return {"metric": METRIC, "value": Q / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": rgg_config(n, Q, world),
            "impl_detail": {"nnz": int(L.nnz), "halo_rows_rank0": int(layer.plan.n_halo), "spmm_rowtile": layer.rowtile[2] if layer.rowtile else None,
                            "cuda_graph": False, "l2": "working set exceeds L2 (768 MB slabs at 1 GPU)"},
            "e2e": {"value": Q / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(hx[0].numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": int(launches), "clocks": clocks,
            "roofline": roof, "cpu_baseline": cpu, "final_loss": float(loss_host)}


def rgg_config(n, Q, world):
    return {"workload": "rgg1m", "description": WORKLOADS["rgg1m"]["desc"], "vertices": n, "K": 8, "H": 3,
            "F": 64, "G": 64, "batch": Q, "optimizer": "SGD lr %g momentum %g" % (LR, MOMENTUM), "parallelism": "rows/%d" % world}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch path on the host cores.  kind "reference": the UNMODIFIED reference layer
# classes (tgcn/nn/gcn.py, gcn_matmul.py) imported from $TGCN_REF or baseline/_ref when such a copy is present;
# kind "port": oracle/layers_torch.py, which performs the reference's ATen call sequence (pinned to its goldens).
# Neither imports the product's kernels (graphs come from build_graph_oracle).
# ------------------------------------------------------------------------------------------------
def _reference_root():
    for cand in (os.environ.get("TGCN_REF"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "tgcn", "nn", "gcn.py")):
            return cand
    return None


def _reference_model(root, Lt, H, K, classes, sparse):
    """pytorch_hcp_tgcn.py:93-155 composed from the reference's OWN layer classes (the example script itself cannot be
    imported: it pulls the HCP data loaders; its removed `torch.rfft` call is written `torch.fft.fft(...).real`)."""
    import importlib
    import types
    import warnings
    import torch.nn as nn
    for name, attrs in (("matplotlib", {}), ("matplotlib.pyplot", {}), ("torch_geometric", {}),
                        ("torch_geometric.utils", {"degree": None, "remove_self_loops": None}), ("torch_scatter", {"scatter_add": None})):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(mod, k, v)
            sys.modules[name] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["torch_geometric"].utils = sys.modules["torch_geometric.utils"]
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = importlib.import_module("tgcn.nn.gcn_matmul" if sparse else "tgcn.nn.gcn")
        pool4 = importlib.import_module("tgcn.nn.gcn").gcn_pool_4

    class NetTGCN(nn.Module):
        def __init__(self, L):
            super().__init__()
            self.tgcn1 = ref.TGCNCheb_H(L[0], 1, 32, K, H)
            self.drop1 = nn.Dropout(0.1)
            self.gcn2 = ref.GCNCheb(L[2], 32, 64, K)
            self.fc1 = nn.Linear(int(L[2].shape[0] * 64 / 4), 200)
            self.dense1_bn = nn.BatchNorm1d(200)
            self.drop2 = nn.Dropout(0.5)
            self.fc2 = nn.Linear(200, classes)

        def forward(self, x):
            x = torch.fft.fft(x, dim=2).real
            x = pool4(self.drop1(F.relu(self.tgcn1(x))))
            x = pool4(F.relu(self.gcn2(x)))
            x = x.view(x.shape[0], -1)
            x = self.drop2(F.relu(self.dense1_bn(self.fc1(x))))
            return F.log_softmax(self.fc2(x), dim=1)
    return NetTGCN(Lt)


def _cpu_operands(Ls, dense):
    out = []
    for L in Ls:
        coo = L.tocoo()
        t = torch.sparse_coo_tensor(np.vstack([coo.row, coo.col]), coo.data.astype(np.float32), coo.shape).coalesce()
        out.append(t.to_dense() if dense else t.to_sparse_csr())
    return out


def cpu_port_throughput(workload, cfg, budget_s=20.0, steps=None, warmup=1):
    from oracle import model_torch
    from tgcn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    graphs, perm, Ls, n_real = build_graph_oracle(workload)
    N0 = Ls[0].shape[0]
    dense_ok = N0 <= 4096
    Lt = _cpu_operands(Ls, dense_ok)       # dense L~ like the example scripts; torch.sparse_csr (gcn_matmul.py slab path) when infeasible
    torch.manual_seed(0)
    H = cfg["H"]
    Q = cfg["batch"]
    sample_q = Q if dense_ok else max(1, min(Q, 2))
    root = _reference_root() if cfg["model"] == "hcp" else None
    kind = "port"
    if root is not None:
        try:
            model = _reference_model(root, Lt, H, cfg["K"], cfg["classes"], sparse=not dense_ok)
            kind = "reference"
        except Exception as exc:                        # noqa: BLE001
            sys.stderr.write("[bench] reference tree at %s not usable (%s): using the oracle port\n" % (root, exc))
            root = None
    if root is None:
        if cfg["model"] == "hcp":
            model = model_torch.PortNetTGCN_HCP(Lt, horizon=H, K=cfg["K"], n_classes=cfg["classes"])
        else:
            model = model_torch.PortNetTGCN_MNIST(Lt, horizon=H, K=cfg["K"], n_classes=cfg["classes"])
    opt = torch.optim.SGD(model.parameters(), lr=LR, momentum=MOMENTUM)
    x = synth.synthetic_signals(sample_q, N0, H, n_real, perm, seed=5)
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
        if steps is not None and (len(times) >= steps or time.perf_counter() - t_start > 150.0):
            break
        if steps is None and (time.perf_counter() - t_start > budget_s or len(times) >= 50):
            break
    med = float(np.median(times))
    return {"value": sample_q / med, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d timed train steps (fwd+loss+bwd+SGD) of %s, batch %d of %d, %s L~, %d threads, median %.1f ms/step"
                      % (len(times), "the unmodified reference layers (%s)" % root if kind == "reference" else
                         "the torch-CPU port of the reference model (oracle/model_torch.py)", sample_q, Q,
                         "dense" if dense_ok else "torch.sparse_csr (gcn_matmul slab path)", cores, med * 1e3),
            "ms_per_step": med * 1e3}


def cpu_rgg_throughput(n, K, H, Fin, G, budget_s=20.0, sample_n=50000):
    """The reference's CPU path (gcn_matmul.py slab form with a torch.sparse_csr L~, oracle port) on a bounded
    sample: the first `sample_n` vertices of the graph; the rate is scaled by sample_n / n."""
    from oracle import graph_np as Gn, model_torch
    from tgcn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    A, _ = synth.rgg_adjacency(n=sample_n, mean_degree=12.0, seed=0)
    Ls = Gn.rescale_laplacian(Gn.laplacian(A, normalized=True), 2).tocsr()
    Lt = _cpu_operands([Ls], dense=False)[0]
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    world = args.gpus
    if args.workload == "rgg1m":
        cpu = cpu_rgg_throughput(args.rgg_n, 8, 3, 64, 64, budget_s=max(args.cpu_budget, 20.0))
        config = rgg_config(args.rgg_n, args.batch or 1, world)
    else:
        cpu = cpu_port_throughput(args.workload, cfg, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
        graphs, perm, Ls, n_real = build_graph_oracle(args.workload)
        config = workload_config(args.workload, world, args.batch or cfg["batch"], Ls, cfg["H"])
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.workload == "rgg1m" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mesh32k", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--engine", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the separate roofline timing loop (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the hcp360 block that rides along at N = 1")
    ap.add_argument("--no-dp-check", action="store_true", help="skip the N-rank vs single-process parameter check (N > 1)")
    ap.add_argument("--no-fused-fc1", action="store_true", help="keep fc1.weight in the regular optimizer (large heads)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--rgg-n", type=int, default=1_000_000, help="vertices of the rgg1m workload")
    ap.add_argument("--dp", default="peer", choices=["peer", "peer-overlap", "peer-serial", "nccl"],
                    help="gradient exchange for N > 1: peer = fused peer-memory allreduce+SGD; peer-overlap = the head group's "
                         "exchange runs under the conv backward; peer-serial = one exchange after the backward; nccl = bucketed "
                         "NCCL + torch SGD")
    ap.add_argument("--rowtile", type=int, default=ROWTILE_DEFAULT, choices=[0, 4, 8],
                    help="rows per tile of the register-tiled SpMM kernel for the streaming layers (0 = per-entry kernels)")
    ap.add_argument("--rowtile-pad", type=int, default=1,
                    help="pad every row tile's plan to a multiple of this many entries (1 = none)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch.distributed as dist
    rank, world, local = init_dist(args)
    if args.workload == "rgg1m":
        line = run_rgg(args, rank, world, local)
    else:
        line = run_model(args, args.workload, rank, world, local)
        if world == 1 and args.workload == "mesh32k" and not args.no_secondary:
            sec = run_model(args, "hcp360", rank, world, local, primary=False)
            line["secondary"] = {"hcp360": {k: sec[k] for k in ("value", "unit", "ms_per_step", "e2e", "config", "impl_detail",
                                                                "gpu_launches_per_step", "roofline", "final_loss")}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # ProcessGroupNCCL teardown blocks while CUDA graphs that captured collectives are alive (observed on
        # torch 2.11 / NCCL 2.28: destroy_process_group never returns): synchronise, flush and leave.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
